"""Replay of the recorded reference-suite trace (tests/golden/ref_suite_trace.*) -- TEST INFRASTRUCTURE.
See tests/golden/record_ref_suite.py (capture) and tests/test_gpu_reference_trace.py (the GPU test)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_EXACT_BASE = {"add", "sub", "mul", "div", "max", "min", "abs", "sign", "sqrt", "clamp", "gather", "gather_axis",
               "broadcast", "maximum", "minimum", "prng_randrange"}
_CR = {"exp", "log", "exp2", "log2", "pow", "invsqrt"}          # <= 0.5001 ulp kernels: at most 1 ulp from the oracle


def _base(op):
    for suf in ("_axis_rebroadcast", "_axis", "_broadcast", "_scalar", "_ss", "_sv", "_vs"):
        if op.endswith(suf):
            op = op[:-len(suf)]
            break
    if op[0] in "ir" and (op[1:] in _EXACT_BASE or op[1:] in _CR or op[1:] in
                          ("sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh")):
        op = op[1:]
    return op


def _tolerance(call):
    """(rtol, atol) for the buffers a call writes; (0, 0) means bit-exact."""
    m = call["method"]
    if call["kind"] == "rng":
        if m == "normal":                 # absolute error scales with stddev (args: n, buffer, mean, stddev)
            return 0.0, 5e-6 * max(1.0, abs(float(call["args"][3]["v"])))
        return 0.0, 0.0
    if m in ("fill", "fill_many", "argreduce", "argsort_u32"):
        return 0.0, 0.0
    if m == "submit":
        b = _base(call["op"])
        if b in _EXACT_BASE:
            return 0.0, 0.0
        if b in _CR:
            return 1.2e-7, 0.0
        if b in ("sum", "prod"):
            return 2e-6, 1e-7             # tree order differs from the oracle's exact sum
        if b in ("matmul", "batch_affine"):
            return 1e-5, 1e-5
        if b.startswith("prng_"):
            return 0.0, 2e-5              # Box-Muller shaders: 4e-6 x stddev (the suite uses stddev <= 3)
        if b.startswith("nn_cross_entropy"):
            return 2e-7, 1e-7
        return 5e-7, 1e-7                 # libdevice-grade transcendentals: <= 2 ulp
    if m == "gemm":
        return 1e-5, 1e-5
    return 1e-6, 1e-7                     # ew_chain (may contain exp / pow), nn kernels


def load_trace():
    calls = json.load(open(os.path.join(HERE, "golden", "ref_suite_trace.json")))["calls"]
    arrays = np.load(os.path.join(HERE, "golden", "ref_suite_trace.npz"))
    return calls, arrays


class _Replay:
    def __init__(self, dev, arrays):
        from vulkpy_b200 import _backend
        self.dev, self.arrays, self.backend = dev, arrays, _backend

    def make_buffer(self, pre, dtype):
        return (self.dev.toU32Buffer if dtype == "uint32" else self.dev.toBuffer)(pre)

    def arr(self, key):
        return self.arrays["a" + key]

    def decode(self, e, slots, outs):
        t = e["t"]
        if t == "buf":
            if e["slot"] not in slots:
                pre = self.arr(e["pre"])
                b = self.make_buffer(pre, e["dtype"])
                slots[e["slot"]] = b
                outs.append((b, e))
            return slots[e["slot"]]
        if t == "host":
            return self.arr(e["a"]).copy()
        if t == "struct":
            return getattr(self.backend, e["cls"]).from_buffer_copy(bytes.fromhex(e["hex"]))
        if t == "list":
            return [self.decode(s, slots, outs) for s in e["v"]]
        return e["v"]


def _compare(call, got, want, where):
    rtol, atol = _tolerance(call)
    if rtol == 0.0 and atol == 0.0:
        np.testing.assert_array_equal(got, want, err_msg=where)
    else:
        np.testing.assert_allclose(got, want, rtol=rtol, atol=atol, err_msg=where)


def replay(dev, make_rng, calls, arrays, exact=False, make_buffer=None):
    """Replay every recorded call on `dev`; compare each written buffer with the recorded contents.
    Returns (device calls, generator calls) replayed."""
    rp = _Replay(dev, arrays)
    if make_buffer is not None:
        rp.make_buffer = make_buffer
    rngs, seeded = {}, {}
    n_dev = n_rng = 0
    for i, call in enumerate(calls):
        where = f"call {i} of {call['test']}: {call.get('method')} {call.get('op', '')}"
        if call["kind"] == "rng_new":
            rngs[call["rng"]] = make_rng(dev, call["size"], call["seed"])
            seeded[call["rng"]] = call["seed"] is not None
            continue
        slots, outs = {}, []
        args = [rp.decode(e, slots, outs) for e in call["args"]]
        kw = {k: rp.decode(e, slots, outs) for k, e in call["kwargs"].items()}
        if call["kind"] == "dev":
            if call["method"] == "submit":
                args[0] = call["op"]            # by shader name: op ids are an implementation detail
            getattr(dev, call["method"])(*args, **kw)
            n_dev += 1
        else:
            getattr(rngs[call["rng"]], call["method"])(*args, **kw)
            n_rng += 1
            if not seeded[call["rng"]] and not exact:
                continue                        # entropy-seeded generator: values are not comparable
        dev.wait()
        for b, e in outs:
            b.host_acquire()                    # host-visible and synchronised (what Array.__array__ does)
            got, want = np.asarray(b.arr if hasattr(b, "arr") else b).copy(), rp.arr(e["post"])
            if exact:
                np.testing.assert_array_equal(got, want, err_msg=where)
            else:
                _compare(call, got, want, where)
    return n_dev, n_rng
