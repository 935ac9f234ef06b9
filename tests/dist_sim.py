"""CPU stand-ins used ONLY by tests/test_dist_gloo.py to drive the sharding logic of
vulkpy_b200.dist without a GPU: a NumPy-backed local array with the vk.Array method surface that
ShardedArray / DataParallel call, and a gloo transport.  Test infrastructure, not product."""
import numpy as np
import torch
import torch.distributed as td

F = np.float32
_RED = {"sum": np.sum, "prod": np.prod, "maximum": np.max, "minimum": np.min}
_TD = {"sum": td.ReduceOp.SUM, "prod": td.ReduceOp.PRODUCT, "maximum": td.ReduceOp.MAX, "minimum": td.ReduceOp.MIN}
_UN = {"abs": np.abs, "sign": np.sign, "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin,
       "acos": np.arccos, "atan": np.arctan, "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh,
       "asinh": np.arcsinh, "acosh": np.arccosh, "atanh": np.arctanh, "exp": np.exp, "log": np.log,
       "exp2": np.exp2, "log2": np.log2, "sqrt": np.sqrt, "invsqrt": lambda x: 1 / np.sqrt(x)}


def _v(x):
    return x.a if isinstance(x, NumpyLocal) else x


class NumpyLocal:
    def __init__(self, a):
        self.a = np.array(a, dtype=F)
        self.job = None

    @property
    def shape(self):
        return self.a.shape

    def wait(self):
        pass

    def reshape(self, shape):
        self.a = self.a.reshape(shape)

    def __array__(self, dtype=None, copy=None):
        return self.a

    def _b(self, o, f):
        return NumpyLocal(np.asarray(f(self.a, _v(o)), dtype=F))

    def __add__(self, o): return self._b(o, np.add)
    def __sub__(self, o): return self._b(o, np.subtract)
    def __mul__(self, o): return self._b(o, np.multiply)
    def __truediv__(self, o): return self._b(o, np.divide)
    def __pow__(self, o): return self._b(o, np.power)
    def __radd__(self, o): return NumpyLocal(F(o) + self.a)
    def __rsub__(self, o): return NumpyLocal(F(o) - self.a)
    def __rmul__(self, o): return NumpyLocal(F(o) * self.a)
    def __rtruediv__(self, o): return NumpyLocal(F(o) / self.a)
    def __matmul__(self, o): return NumpyLocal(self.a @ _v(o))

    def _i(self, o, f):
        self.a = np.asarray(f(self.a, _v(o)), dtype=F)
        return self

    def __iadd__(self, o): return self._i(o, np.add)
    def __isub__(self, o): return self._i(o, np.subtract)
    def __imul__(self, o): return self._i(o, np.multiply)
    def __itruediv__(self, o): return self._i(o, np.divide)

    def max(self, o, inplace=False):
        return self._i(o, np.maximum) if inplace else self._b(o, np.maximum)

    def min(self, o, inplace=False):
        return self._i(o, np.minimum) if inplace else self._b(o, np.minimum)

    def clamp(self, lo, hi, inplace=False):
        r = np.minimum(np.maximum(self.a, _v(lo)), _v(hi)).astype(F)
        if inplace:
            self.a = r
            return self
        return NumpyLocal(r)

    def broadcast_to(self, shape):
        return NumpyLocal(np.broadcast_to(self.a, shape))

    def _red(self, name, axis, keepdims, rebroadcast):
        f = _RED[name]
        if rebroadcast:
            return NumpyLocal(np.broadcast_to(f(self.a, axis=axis, keepdims=True), self.a.shape))
        if axis is None:
            r = np.asarray(f(self.a), dtype=F).reshape((1,) * self.a.ndim if keepdims else (1,))
            return NumpyLocal(r)
        ax = tuple(int(x) for x in np.asarray(axis).reshape(-1))
        return NumpyLocal(f(self.a, axis=ax, keepdims=keepdims))

    def sum(self, axis=None, keepdims=False, rebroadcast=False): return self._red("sum", axis, keepdims, rebroadcast)
    def prod(self, axis=None, keepdims=False, rebroadcast=False): return self._red("prod", axis, keepdims, rebroadcast)
    def maximum(self, axis=None, keepdims=False, rebroadcast=False): return self._red("maximum", axis, keepdims, rebroadcast)
    def minimum(self, axis=None, keepdims=False, rebroadcast=False): return self._red("minimum", axis, keepdims, rebroadcast)

    def gather(self, idx, axis=None):
        return NumpyLocal(self.a.reshape(-1)[np.asarray(idx, dtype=np.int64)])


for _name, _f in _UN.items():
    def _mk(f):
        def method(self, inplace=False):
            r = f(self.a).astype(F)
            if inplace:
                self.a = r
                return self
            return NumpyLocal(r)
        return method
    setattr(NumpyLocal, _name, _mk(_f))


class GlooTransport:
    """torch.distributed (gloo) collectives on the NumPy locals."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.calls = []          # (kind, number of elements) log: tests assert WHICH ops communicate

    def allreduce(self, arr, op):
        t = torch.from_numpy(np.ascontiguousarray(arr.a))
        td.all_reduce(t, op=_TD[op])
        arr.a = t.numpy().reshape(arr.a.shape)
        self.calls.append(("allreduce", arr.a.size))
        return arr

    def allgather(self, arr, out):
        parts = [torch.empty_like(torch.from_numpy(arr.a)) for _ in range(self.world)]
        td.all_gather(parts, torch.from_numpy(np.ascontiguousarray(arr.a)))
        out.a = np.concatenate([p.numpy() for p in parts], axis=0).reshape(out.a.shape)
        self.calls.append(("allgather", arr.a.size))
        return out

    def new_array(self, shape):
        return NumpyLocal(np.zeros(shape, F))


class GlooFusedTransport(GlooTransport):
    """Same, plus the fused entry point of the NCCL transport (``reduce_allreduce``): the local
    ``[prev, axis, post] -> [prev, post]`` reduction and its exchange as one call."""

    def reduce_allreduce(self, arr, op, prev, axis, post, out_shape):
        assert arr.a.size == prev * axis * post
        part = _RED[op](arr.a.reshape(prev, axis, post), axis=1) if axis else np.full((prev, post), {"sum": 0, "prod": 1, "maximum": -np.inf, "minimum": np.inf}[op], F)
        t = torch.from_numpy(np.ascontiguousarray(part, dtype=F))
        td.all_reduce(t, op=_TD[op])
        self.calls.append(("reduce_allreduce", int(t.numel())))
        return NumpyLocal(t.numpy().reshape(out_shape))


class GlooBucketTransport(GlooTransport):
    """Same, plus the bucket entry point of the NCCL transport (``allreduce_many``): every tensor of
    the bucket crosses in ONE collective, then ``x *= scale`` in float32 on each."""

    def allreduce_many(self, arrs, op, scale=1.0):
        flat = torch.from_numpy(np.concatenate([np.ascontiguousarray(a.a).reshape(-1) for a in arrs]))
        td.all_reduce(flat, op=_TD[op])
        self.calls.append(("allreduce_many", int(flat.numel()), len(arrs)))
        off = 0
        for a in arrs:
            n = a.a.size
            a.a = np.asarray(flat[off:off + n].numpy().reshape(a.a.shape) * F(scale), dtype=F)
            off += n
        return arrs



class GlooPullTransport(GlooTransport):
    """Stand-in for ``NcclTransport.matmul_allgather``: the row-sharded matmul as the device kernel
    does it -- start on the local K range, then walk the peers' ranges in ring order
    (rank+1, rank+2, ...), accumulating in that order.  Shapes it "does not take" (fewer than
    ``min_rows`` local rows) return ``None`` so that the caller's all-gather fallback is exercised."""
    min_rows = 4

    def matmul_allgather(self, a, b_shard, n_cols):
        M, K = a.a.shape
        kc = K // self.world
        if K % self.world or b_shard.a.shape[0] != kc or M < self.min_rows:
            return None
        parts = [torch.empty_like(torch.from_numpy(b_shard.a)) for _ in range(self.world)]
        td.all_gather(parts, torch.from_numpy(np.ascontiguousarray(b_shard.a)))     # the "peer memory"
        self.calls.append(("pull", b_shard.a.size * (self.world - 1)))
        acc = np.zeros((M, n_cols), F)
        for j in range(self.world):
            s = (self.rank + j) % self.world
            acc = acc + a.a[:, s * kc:(s + 1) * kc] @ parts[s].numpy()
        return NumpyLocal(acc)
