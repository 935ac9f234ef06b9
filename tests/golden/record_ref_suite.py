"""Capture the device-level trace of the reference's own test-suite -- TEST INFRASTRUCTURE.

    python tests/golden/record_ref_suite.py          # needs /root/reference (development container)

Runs /root/reference/test (test_vulkpy.py, test_random.py, test_nn.py: the 233 known-answer tests
of the reference's CI, Dockerfile:27-33) unchanged and in place against this repository's Python
layer over the NumPy stand-in device (tests/fake_device.py, kernels answered by the CPU oracle),
and records EVERY call the suite makes below the Python layer: the method (``submit`` with its
shader name and parameter block, ``gemm``, ``ew_chain``, the nn kernels, the generator calls), the
contents of every buffer before the call and after it.  The run itself must pass (all reference
assertions hold for the recorded answers), so the trace is a set of known answers the reference's
tests accept.  ``tests/test_gpu_reference_trace.py`` replays the trace call by call through the
C ABI on the CUDA device and compares buffer contents -- that is how the reference's acceptance
suite travels to a box where /root/reference does not exist.  Nothing of the reference's sources is
stored: only call names, parameter bytes and array contents.

Output: tests/golden/ref_suite_trace.npz (arrays, de-duplicated) + ref_suite_trace.json (calls).
"""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/test"

_arrays = {}        # digest -> array
_calls = []
_current = {"test": ""}
_rng_ids = {}


def _store(a):
    a = np.ascontiguousarray(a)
    key = hashlib.sha1(a.tobytes() + str(a.dtype).encode() + str(a.shape).encode()).hexdigest()[:16]
    _arrays.setdefault(key, a.copy())
    return key


def _enc_pre(x, seen):
    """Encode one argument before the call; buffers get a slot number so aliasing inside a call survives."""
    import fake_device
    if isinstance(x, fake_device.FakeBuffer):
        slot = seen.setdefault(id(x), len(seen))
        return {"t": "buf", "slot": slot, "dtype": str(x.arr.dtype), "pre": _store(x.arr), "_obj": x}
    if isinstance(x, np.ndarray):
        return {"t": "host", "a": _store(x)}
    if isinstance(x, C.Structure):
        return {"t": "struct", "cls": type(x).__name__, "hex": bytes(x).hex()}
    if isinstance(x, (list, tuple)):
        return {"t": "list", "v": [_enc_pre(e, seen) for e in x]}
    if x is None or isinstance(x, (bool, int, float, str)):
        return {"t": "val", "v": x}
    if isinstance(x, (np.integer, np.floating, np.bool_)):
        return {"t": "val", "v": x.item()}
    raise TypeError(f"cannot record argument of type {type(x)!r}")


def _enc_post(e):
    if e["t"] == "buf":
        e["post"] = _store(e.pop("_obj").arr)
    elif e["t"] == "list":
        for s in e["v"]:
            _enc_post(s)


def _wrap(cls, name, kind):
    import fake_device
    orig = getattr(cls, name)

    def rec(self, *args, **kw):
        seen = {}
        enc = [_enc_pre(a, seen) for a in args]
        enck = {k: _enc_pre(v, seen) for k, v in kw.items()}
        out = orig(self, *args, **kw)
        for e in enc:
            _enc_post(e)
        for e in enck.values():
            _enc_post(e)
        call = {"test": _current["test"], "kind": kind, "method": name, "args": enc, "kwargs": enck}
        if kind == "dev" and name == "submit":
            spv = args[0]
            call["op"] = fake_device._NAMES[spv] if isinstance(spv, int) else os.path.basename(str(spv)).replace(".spv", "")
        if kind == "rng":
            call["rng"] = _rng_ids[id(self)]
        _calls.append(call)
        return out

    setattr(cls, name, rec)


def install():
    import fake_device
    for m in ("submit", "ew_chain", "fill", "fill_many", "gemm", "argreduce", "argsort_u32", "nn_adam",
              "nn_adam_apply_many", "nn_activation_backward", "nn_softmax_forward", "nn_softmax_ce_train"):
        _wrap(fake_device.FakeDevice, m, "dev")
    for m in ("random_uint32", "random_float", "normal", "advance"):
        _wrap(fake_device.FakeRng, m, "rng")
    init = fake_device.FakeRng.__init__

    def rng_init(self, gpu, spv_uint32="", spv_float="", size=64, seed=None):
        init(self, gpu, spv_uint32, spv_float, size, seed)
        _rng_ids[id(self)] = _current["n_rng"] = _current.get("n_rng", -1) + 1     # id() values get reused
        _calls.append({"test": _current["test"], "kind": "rng_new", "rng": _rng_ids[id(self)], "size": int(size),
                       "seed": None if seed is None else int(seed)})

    fake_device.FakeRng.__init__ = rng_init


# ---- pytest plugin hooks (this module is passed with -p) -------------------------------------------
def pytest_configure(config):
    import ref_suite_plugin  # noqa: F401  (routes vk.GPU() / the generator to the stand-ins)
    install()


def pytest_runtest_setup(item):
    _current["test"] = item.nodeid.split("/")[-1]


def pytest_sessionfinish(session, exitstatus):
    if exitstatus != 0:
        return
    np.savez_compressed(os.path.join(HERE, "ref_suite_trace.npz"), **{"a" + k: v for k, v in _arrays.items()})
    with open(os.path.join(HERE, "ref_suite_trace.json"), "w") as f:
        json.dump({"source": "reference test-suite run in place over tests/fake_device.py (oracle kernels)",
                   "calls": _calls}, f, separators=(",", ":"))
    print(f"\nrecorded {len(_calls)} calls, {len(_arrays)} distinct arrays")


if __name__ == "__main__":
    import subprocess
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "tests"), HERE, env.get("PYTHONPATH", "")])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", REF, "-q", "-p", "record_ref_suite", "-p", "no:cacheprovider",
                        "--rootdir", "/tmp"], env=env, cwd="/tmp")
    sys.exit(r.returncode)
