"""
ctypes loader for oracle/cpu_ref.c (TEST / BASELINE INFRASTRUCTURE ONLY).

``load()`` returns the shared library, building it with ``make -C oracle`` when missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcpu_ref.so")

BINARY = {"add": 0, "sub": 1, "mul": 2, "div": 3, "max": 4, "min": 5, "pow": 6}
UNARY = {n: i for i, n in enumerate(
    ["abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh",
     "acosh", "atanh", "exp", "log", "exp2", "log2", "sqrt", "invsqrt"])}
REDUCE = {"sum": 0, "prod": 1, "maximum": 2, "minimum": 3}

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "cpu_ref.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.ref_reduce_full.restype = C.c_float
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def ptr(a):
    return C.c_void_p(a.ctypes.data)
