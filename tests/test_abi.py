"""C-ABI surface: the shared library loads on a machine without a GPU, exports every symbol that
include/vulkpy_b200.h declares, resolves all 121 reference shader names, and the Python-side
parameter blocks have the layout of the reference's OpParams structs (_vkarray.cc:132-203).
No compute call is made here."""
import ctypes
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vulkpy_b200.h")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"^VKP_API\s+[\w\s\*]+?\b(vkp_\w+)\s*\(", text, flags=re.M)))


def test_header_declares_functions():
    syms = declared_symbols()
    assert len(syms) >= 40
    assert "vkp_submit" in syms and "vkp_rng_create" in syms and "vkp_gemm" in syms


def test_library_exports_every_declared_symbol():
    from vulkpy_b200 import _backend
    lib = ctypes.CDLL(_backend.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_covers_the_header():
    from vulkpy_b200 import _backend
    assert sorted(_backend.PROTOTYPES) == declared_symbols()


def test_abi_version_and_error_string():
    from vulkpy_b200 import _backend
    assert _backend.lib.vkp_abi_version() == 1
    assert isinstance(_backend.lib.vkp_last_error(), bytes)


def test_every_reference_shader_has_an_op():
    from vulkpy_b200 import _backend
    names = json.load(open(os.path.join(GOLDEN, "shader_names.json")))
    assert len(names) == 121
    for n in names:
        assert _backend.lib.vkp_op_id(n.encode()) >= 0, n
        # the reference passes the .spv path (util.py:58-72)
        assert _backend.lib.vkp_op_id(f"/x/vulkpy/shader/{n}.spv".encode()) == _backend.lib.vkp_op_id(n.encode())
    assert _backend.lib.vkp_op_count() == 121
    assert _backend.lib.vkp_op_id(b"no_such_shader") == -1


def test_param_block_layouts():
    from vulkpy_b200 import _backend as b
    assert ctypes.sizeof(b.VectorParams) == 4
    assert ctypes.sizeof(b.MultiVector2Params) == 8
    assert ctypes.sizeof(b.VectorScalarParams) == 8
    assert ctypes.sizeof(b.VectorScalar2Params) == 12
    assert ctypes.sizeof(b.MatMulParams) == 12
    assert ctypes.sizeof(b.AxisReductionParams) == 12
    assert ctypes.sizeof(b.BroadcastParams) == 12
    assert ctypes.sizeof(b.Multi3BroadcastParams) == 16
    assert ctypes.sizeof(b.BatchAffineParams) == 12
    assert ctypes.sizeof(b.VectorRangeParams) == 12
    assert ctypes.sizeof(b.AxisGatherParams) == 16
    p = b.VectorScalar2Params(7, 1.5, -2.0)
    assert (p.size, p.scalar0, p.scalar1) == (7, 1.5, -2.0)
    g = b.AxisGatherParams(2, 3, 4, 5)  # prev, post, axis, index (_vkarray.cc:197-202)
    assert (g.prev_prod, g.post_prod, g.axis_size, g.index_size) == (2, 3, 4, 5)


def test_public_names():
    import vulkpy_b200 as vk
    for name in ("GPU", "U32Array", "Shape", "Array", "zeros", "random", "nn", "util"):
        assert hasattr(vk, name)
    for name in ("abs sign sin cos tan asin acos atan sinh cosh tanh asinh acosh atanh exp log exp2 "
                 "log2 sqrt invsqrt max min clamp sum prod maximum minimum mean broadcast_to gather "
                 "reshape wait flush").split():
        assert callable(getattr(vk.Array, name)), name
    from vulkpy_b200 import nn
    for name in ("Optimizer OptimizerState Loss Regularizer Module Constant HeNormal SGD SGDState Adam "
                 "AdamState AdaGrad AdaGradState Dense ReLU Sigmoid Softmax CrossEntropyLoss "
                 "SoftmaxCrossEntropyLoss MSELoss HuberLoss Lasso Ridge Elastic Sequence").split():
        assert hasattr(nn, name), name
    from vulkpy_b200.nn.parameters import Parameter  # test/test_nn.py:483 imports it this way
    assert Parameter


def test_vulkpy_alias_package():
    import vulkpy_b200
    import vulkpy
    from vulkpy.util import enable_debug
    from vulkpy.nn.parameters import Parameter
    assert vulkpy.Array is vulkpy_b200.Array and callable(enable_debug) and Parameter


def test_no_device_fails_loudly():
    """Without a CUDA device the product raises; it never falls back to a CPU path."""
    from vulkpy_b200 import _backend
    n = ctypes.c_int(0)
    rc = _backend.lib.vkp_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    import vulkpy_b200 as vk
    with pytest.raises(RuntimeError):
        vk.GPU()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "vulkpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("vulkpy_oracle", "oracle") or f == "__none__", \
                    f"{f} mentions the oracle"


def test_compat_shim_exports_the_reference_extension_surface():
    """vulkpy_b200/compat/_vkarray.py defines every name the reference's Python layer takes from its
    pybind11 module (vkarray.py:26-44, random.py:26; _vkarray.cc:756-898) -- no device needed."""
    import importlib
    shim = importlib.import_module("vulkpy_b200.compat._vkarray")
    for name in ("createGPU", "DataShape", "VectorParams", "MultiVector2Params", "VectorScalarParams",
                 "VectorScalar2Params", "MatMulParams", "AxisReductionParams", "BroadcastParams",
                 "Multi3BroadcastParams", "BatchAffineParams", "AxisGatherParams", "VectorRangeParams", "Job",
                 "Buffer", "Shape", "GPU", "Xoshiro128pp"):
        assert hasattr(shim, name), name
    for meth in ("createBuffer", "createU32Buffer", "toBuffer", "toU32Buffer", "submit", "wait", "flush", "canSubgroupArithmetic"):
        assert callable(getattr(shim.GPU, meth))
    assert shim.VectorScalar2Params(7, 1.5, 2.5).scalar1 == 2.5 and shim.AxisGatherParams(1, 2, 3, 4).index_size == 4
    assert shim._first_shape_binding("x/add_broadcast.spv") == 3 and shim._first_shape_binding("iadd_broadcast") == 2
    assert shim._first_shape_binding("broadcast.spv") == 2 and shim._first_shape_binding("add.spv") is None
