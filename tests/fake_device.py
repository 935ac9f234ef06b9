"""NumPy stand-in for the device object behind ``vk.GPU`` -- TEST INFRASTRUCTURE ONLY.

``tests/test_host_logic.py`` (CPU, no GPU) uses it to drive the Python layer of ``vulkpy_b200``
(shape rules, parameter blocks, result allocation, job / keep-alive bookkeeping, the nn
compositions and their fused variants) with every "kernel" answered by the CPU oracle
(``oracle/vulkpy_oracle.py``).  The product never imports this file; without the CUDA library and a
device ``vk.GPU()`` fails loudly.  What is checked here is everything ABOVE the C ABI: that the
Python layer asks for the right operation on the right buffers with the right parameters.
"""
from __future__ import annotations

import numpy as np

from oracle import vulkpy_oracle as orc
from vulkpy_b200 import _backend as _b

F = np.float32
BIN = ("add", "sub", "mul", "div", "max", "min", "pow")
UN = ("abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh",
      "exp", "log", "exp2", "log2", "sqrt", "invsqrt")
RED = ("sum", "prod", "maximum", "minimum")
_NAMES = {i: _b.lib.vkp_op_name(i).decode() for i in range(_b.lib.vkp_op_count())}


class FakeJob:
    def wait(self, timeout_ns=None):
        pass

    def done(self):
        return True


class FakeBuffer(_b._BufferBase):
    """Host array with the method surface of ``Buffer`` / ``Shape``."""

    def __init__(self, dev, n, dtype):      # no device allocation
        self._dev, self._n, self.ptr = dev, int(n), None
        self.arr = np.zeros(self._n, dtype=dtype)

    def host_view(self):
        pass

    def host_acquire(self, prefetch=False, write=False):
        pass

    @property
    def __array_interface__(self):
        return self.arr.__array_interface__

    def upload(self, host):
        self.arr[:] = np.asarray(host).reshape(-1)

    def __del__(self):
        pass


def _shape3(a, p):
    return a.reshape(p.prev_prod, p.axis_size, p.post_prod)


class FakeDevice:
    """The ``Device`` methods the Python layer calls, answered on the host by the oracle."""
    index = 0

    def __init__(self):
        self.launches = 0
        self.log = []            # (kernel name, number of buffers): tests assert WHAT was launched

    # -- buffers -------------------------------------------------------------------------------
    def createBuffer(self, n):
        return FakeBuffer(self, n, F)

    def createU32Buffer(self, n):
        return FakeBuffer(self, n, np.uint32)

    def wait(self):
        pass

    def flush(self, ranges):
        pass

    def canSubgroupArithmetic(self):
        return True

    def sm_count(self):
        return 148

    def launch_count(self):
        return self.launches

    def _done(self, name, nbuf):
        self.launches += 1
        self.log.append((name, nbuf))
        return FakeJob()

    # -- the 121 shader names --------------------------------------------------------------------
    def submit(self, spv, x, y, z, infos, shape, params, wait=()):
        name = _NAMES[spv] if isinstance(spv, int) else str(spv)
        v = [b.arr if isinstance(b, FakeBuffer) else np.asarray(b) for b in infos]
        p = params
        if name in BIN:
            v[2][:] = orc.binary(name, v[0], v[1])
        elif name[0] == "i" and name[1:] in BIN:
            v[0][:] = orc.binary(name[1:], v[0], v[1])
        elif name.endswith("_scalar"):
            base = name[:-7]
            if base in BIN:
                v[1][:] = orc.scalar(base, v[0], p.scalar)
            elif base[0] == "r" and base[1:] in BIN:
                v[1][:] = orc.scalar(base[1:], v[0], p.scalar, reverse=True)
            elif base[0] == "i" and base[1:] in BIN:
                v[0][:] = orc.scalar(base[1:], v[0], p.scalar)
            else:
                raise RuntimeError("Unknown Operation")
        elif name.endswith("_broadcast"):
            base, nd = name[:-10], int(p.ndim)
            if base in BIN:       # A, B, C, [shapeA | shapeB | shapeC]
                sa, sb, sc = (tuple(int(s) for s in v[3][i * nd:(i + 1) * nd]) for i in range(3))
                assert v[0].size == p.size0 and v[1].size == p.size1 and v[2].size == p.size2
                out = orc.broadcast_binary(base, v[0].reshape(sa), v[1].reshape(sb))
                assert out.shape == sc
                v[2][:] = out.reshape(-1)
            else:                 # in place: A, B, [shapeA | shapeB]
                sa, sb = (tuple(int(s) for s in v[2][i * nd:(i + 1) * nd]) for i in range(2))
                out = orc.broadcast_binary(base[1:], v[0].reshape(sa), v[1].reshape(sb))
                assert out.shape == sa
                v[0][:] = out.reshape(-1)
        elif name == "broadcast":  # A, B, shapeA, shapeB
            sa, sb = tuple(int(s) for s in v[2]), tuple(int(s) for s in v[3])
            v[1][:] = orc.broadcast_to(v[0].reshape(sa), sb).reshape(-1)
        elif name in UN:
            v[1][:] = orc.unary(name, v[0])
        elif name[0] == "i" and name[1:] in UN:
            v[0][:] = orc.unary(name[1:], v[0])
        elif name in ("clamp", "iclamp"):
            v[-1 if name == "clamp" else 0][:] = orc.clamp(v[0], v[1], v[2])
        elif name in ("clamp_sv", "iclamp_sv"):      # array max, scalar min
            v[-1 if name == "clamp_sv" else 0][:] = orc.clamp(v[0], F(p.scalar), v[1])
        elif name in ("clamp_vs", "iclamp_vs"):      # array min, scalar max
            v[-1 if name == "clamp_vs" else 0][:] = orc.clamp(v[0], v[1], F(p.scalar))
        elif name in ("clamp_ss", "iclamp_ss"):
            v[-1 if name == "clamp_ss" else 0][:] = orc.clamp(v[0], F(p.scalar0), F(p.scalar1))
        elif name in RED or (name.endswith("_v1.3") and name[:-5] in RED):
            op = name[:-5] if name.endswith("_v1.3") else name
            nb = 1 if name.endswith("_v1.3") else int(p.size1)
            for i in range(nb):   # literal sum.comp semantics: b[i] = reduce a[i::sizeB]
                v[1][i] = orc.reduce_full_exact(op, v[0][i::nb]) if op in ("sum", "prod") else \
                    (v[0][i::nb].max() if op == "maximum" else v[0][i::nb].min())
        elif name.endswith("_axis_rebroadcast"):
            v[1][:] = orc.reduce_axis(name[:-17], _shape3(v[0], p), 1, rebroadcast=True).reshape(-1)
        elif name.endswith("_axis") and name[:-5] in RED:
            v[1][:] = orc.reduce_axis(name[:-5], _shape3(v[0], p), 1).reshape(-1)
        elif name == "gather":
            v[2][:] = orc.gather(v[0], v[1])
        elif name == "gather_axis":   # A [prev, axis, post], idx, C [idx, prev, post]
            a3 = v[0].reshape(p.prev_prod, p.axis_size, p.post_prod)
            v[2][:] = np.moveaxis(a3[:, v[1].astype(np.int64), :], 1, 0).reshape(-1)
        elif name == "matmul":
            v[2][:] = orc.matmul(v[0].reshape(p.rowA, p.contractSize), v[1].reshape(p.contractSize, p.columnB)).reshape(-1)
        elif name == "batch_affine":  # W, b, X, Y
            w = v[0].reshape(p.output_size, p.input_size)
            v[3][:] = orc.batch_affine(w, v[1], v[2].reshape(p.batch_size, p.input_size)).reshape(-1)
        elif name == "nn_cross_entropy":
            v[2][:] = orc.cross_entropy(v[0], v[1])
        elif name == "nn_cross_entropy_backward":
            v[2][:] = orc.cross_entropy_backward(v[0], v[1])
        elif name == "prng_box_muller":
            n = int(p.size)
            v[1][:n] = orc.box_muller(v[0], n, p.scalar0, p.scalar1)
        elif name == "prng_ibox_muller":
            n = int(p.size)
            v[0][:n] = orc.box_muller(v[0].copy(), n, p.scalar0, p.scalar1)
        elif name == "prng_randrange":
            v[1][:] = orc.randrange_shader(v[0], int(p.low), int(p.high))
        else:
            raise RuntimeError("Unknown Operation")
        return self._done(name, len(infos))

    # -- entry points outside the shader list ------------------------------------------------------
    def ew_chain(self, inputs, out, ops, srcs, scalars):
        """vkp_ew_chain: the recorded chain evaluated step by step with the oracle's one-rounding ops."""
        v = [b.arr for b in inputs]
        acc, tmp = v[0].copy(), None
        names = list(BIN) + ["rsub", "rdiv", "rpow"]
        for op, src, s in zip(ops, srcs, scalars):
            if op == 31:
                tmp = acc.copy()
                continue
            if op >= 11:
                acc = orc.unary(UN[op - 11], acc)
                continue
            rev = op >= 7
            base = names[op][1:] if rev else names[op]
            if src == 0:
                acc = orc.scalar(base, acc, s, reverse=rev)
            else:
                b = tmp if src == 4 else v[src]
                acc = orc.binary(base, b, acc) if rev else orc.binary(base, acc, b)
        out.arr[:] = acc
        return self._done("ew_chain", len(inputs) + 1)

    def fill(self, buf, bits):
        buf.arr.view(np.uint32)[:] = np.uint32(bits & 0xFFFFFFFF)
        return self._done("fill", 1)

    def fill_many(self, bufs, bits):
        for b in bufs:
            b.arr.view(np.uint32)[:] = np.uint32(bits & 0xFFFFFFFF)
        return self._done("fill_many", len(bufs))

    def gemm(self, transA, transB, M, N, K, A, B, Cbuf, bias=None, flags=0, relu_mask=None):
        a = A.arr.reshape((K, M) if transA else (M, K)).astype(np.float64)
        b = B.arr.reshape((N, K) if transB else (K, N)).astype(np.float64)
        c = (a.T if transA else a) @ (b.T if transB else b)
        c = c.astype(F)
        if bias is not None:
            c = (c + bias.arr[None, :]).astype(F)
        if flags & 4:
            c = (Cbuf.arr.reshape(M, N) + c).astype(F)
        if flags & 8:                       # VKP_GEMM_RELU: x.max(0.0)
            c = orc.scalar("max", c, 0.0)
        if relu_mask is not None:           # ReLU.backward: max(sign(y), 0) * dx
            c = (np.maximum(np.sign(relu_mask.arr.reshape(M, N)), F(0)) * c).astype(F)
        Cbuf.arr[:] = c.reshape(-1)
        return self._done("gemm", 3)

    def argreduce(self, op, src, dst, prev, axis, post):
        a = src.arr.reshape(prev, axis, post)
        dst.arr[:] = (orc.argmax if op == 0 else orc.argmin)(a, 1).reshape(-1)
        return self._done("argreduce", 2)

    def argsort_u32(self, keys, dst):
        dst.arr[:] = orc.permutation_from_keys(keys.arr)
        return self._done("argsort_u32", 2)

    def nn_adam(self, grad, m, v, diff, b1, omb1, b2, omb2, c1, c2, eps, neg_lr):
        g = grad.arr
        m.arr[:] = (m.arr * F(b1)).astype(F)
        m.arr[:] = (m.arr + (F(omb1) * g).astype(F)).astype(F)
        v.arr[:] = (v.arr * F(b2)).astype(F)
        v.arr[:] = (v.arr + (F(omb2) * (g * g).astype(F)).astype(F)).astype(F)
        mh = (m.arr / F(c1)).astype(F)
        vh = (np.sqrt((v.arr / F(c2)).astype(F)).astype(F) + F(eps)).astype(F)
        diff.arr[:] = ((mh * F(neg_lr)).astype(F) / vh).astype(F)
        return self._done("nn_adam", 4)

    def nn_adam_apply_many(self, grads, ms, vs, values, scalars):
        for g, m, v, val, s in zip(grads, ms, vs, values, scalars):
            diff = FakeBuffer(self, g.size(), F)
            self.nn_adam(g, m, v, diff, *s)
            self.launches -= 1
            self.log.pop()
            val.arr[:] = (val.arr + diff.arr).astype(F)
        return self._done("nn_adam_apply_many", 4 * len(grads))

    def nn_activation_backward(self, kind, y, dy, dx):
        if kind == 0:
            dx.arr[:] = (np.maximum(orc.unary("sign", y.arr), F(0)) * dy.arr).astype(F)
        else:
            dx.arr[:] = (((F(1) - y.arr).astype(F) * y.arr).astype(F) * dy.arr).astype(F)
        return self._done("nn_activation_backward", 3)

    def nn_softmax_forward(self, x, y, rows, cols):
        xr = x.arr.reshape(rows, cols)
        e = orc.unary("exp", (xr - xr.max(axis=1, keepdims=True)).astype(F))
        s = orc.reduce_axis("sum", e, 1)
        y.arr[:] = (e / s[:, None]).astype(F).reshape(-1)
        return self._done("nn_softmax_forward", 2)

    def nn_softmax_ce_train(self, z, t, p, L, dz, rows, cols, scale):
        self.nn_softmax_forward(z, p, rows, cols)
        self.launches -= 1
        self.log.pop()
        L.arr[:] = orc.cross_entropy(p.arr, t.arr)
        g = orc.cross_entropy_backward(p.arr, t.arr)
        if scale is not None:
            g = (g * F(scale)).astype(F)
        dz.arr[:] = (((F(1) - p.arr).astype(F) * p.arr).astype(F) * g).astype(F)
        return self._done("nn_softmax_ce_train", 5)


class FakeRng:
    """``_backend.Xoshiro128pp`` over the oracle generator."""

    def __init__(self, gpu, spv_uint32="", spv_float="", size=64, seed=None):
        self.size = int(size)
        self._o = orc.Xoshiro128pp(self.size, 0 if seed is None else int(seed))
        self._dev = gpu

    def random_uint32(self, n, info):
        info.arr[:n] = self._o.randint(int(n))
        return self._dev._done("rng_uint32", 1)

    def random_float(self, n, info):
        info.arr[:n] = self._o.random(int(n))
        return self._dev._done("rng_float", 1)

    def normal(self, n, info, mean, stddev):
        info.arr[:n] = self._o.normal(int(n), mean, stddev)
        return self._dev._done("rng_normal", 1)

    def advance(self, n):
        if n:
            self._o.randint(int(n))

    def state(self):
        return np.asarray(self._o.state).copy()


def install(monkeypatch):
    """Route ``vk.GPU()`` and ``vk.random.Xoshiro128pp`` to the stand-ins for one test."""
    import vulkpy_b200.vkarray as vkarray
    monkeypatch.setattr(vkarray, "createGPU", lambda idx, priority: FakeDevice())
    monkeypatch.setattr(_b, "Xoshiro128pp", FakeRng)
