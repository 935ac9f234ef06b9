"""
Multi-GPU sharding of the array hot path (:mod:`vulkpy_b200.dist`) -- additive, single-GPU
behaviour is untouched (the reference has no multi-device surface: ``GPU(idx)`` only picks a
device, vulkpy/vkarray.py:89-101, vulkpy/_vkarray.cc:485).

One process per GPU (``torchrun``), arrays sharded by LEADING axis in contiguous row blocks
(C order: every shard is one contiguous range).  Communication only where the path has a real
exchange step:

==============================  ==================================================
op class                         exchange
==============================  ==================================================
element-wise / scalar / unary    none
broadcast binary                 none (small operand replicated or sharded alike)
reduction over axes >= 1         none
full reduction, axis-0 reduction all-reduce (sum / prod / max / min) of the partials
``A @ B`` with B row(K)-sharded  all-gather of B, then local GEMM
gather (table replicated)        none
Xoshiro128pp                     none: every rank jumps its lanes ahead (GF(2) matrices)
nn data parallel                 all-reduce(sum) of gradients, scaled by 1/world
==============================  ==================================================

The collectives run through ``vkp_comm_*`` (NCCL on the context stream).  ``Transport`` is the
only seam: tests drive the same sharding logic on CPU with a gloo transport and NumPy-backed
local arrays (tests/dist_sim.py); the product only ever constructs ``NcclTransport``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence, Tuple, Union

import numpy as np

__all__ = ["shard_bounds", "Group", "ShardedArray", "DataParallel", "NcclTransport", "gather_replicated"]

_OPS = {"sum": 0, "prod": 1, "maximum": 2, "minimum": 3}


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows ``[lo, hi)`` of a length-``n`` leading axis owned by ``rank``: contiguous blocks whose
    sizes differ by at most one (the first ``n % world`` ranks get the extra row)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class NcclTransport:
    """Collectives on device buffers through the C ABI (``vkp_comm_*``: NCCL over NVLink)."""

    def __init__(self, gpu, rank: int, world: int, unique_id: bytes):
        from . import _backend as b
        self._b = b
        self.gpu, self.rank, self.world = gpu, rank, world
        buf = C.create_string_buffer(unique_id, b.COMM_ID_BYTES)
        b._check(b.lib.vkp_comm_init(gpu.gpu._ctx, world, rank, buf))

    @staticmethod
    def new_unique_id() -> bytes:
        from . import _backend as b
        buf = C.create_string_buffer(b.COMM_ID_BYTES)
        b._check(b.lib.vkp_comm_unique_id(buf))
        return buf.raw

    def allreduce(self, arr, op: str):
        """In-place all-reduce of a local ``vk.Array``; returns it (job attached)."""
        b = self._b
        job = C.c_void_p()
        b._check(b.lib.vkp_comm_allreduce(self.gpu.gpu._ctx, arr.buffer.ptr, arr.buffer.ptr,
                                          arr.buffer.size(), _OPS[op], C.byref(job)))
        arr.job = b.Job(job.value)
        return arr

    def allreduce_many(self, arrs, op: str, scale: float = 1.0):
        """In-place all-reduce of up to 16 local arrays as ONE grouped NCCL launch, then
        ``x *= scale`` on all of them in one kernel (``vkp_comm_allreduce_multi``)."""
        b = self._b
        for i in range(0, len(arrs), 16):
            part = arrs[i:i + 16]
            n = len(part)
            ptrs = (C.c_void_p * n)(*[a.buffer.ptr for a in part])
            counts = (C.c_size_t * n)(*[a.buffer.size() for a in part])
            job = C.c_void_p()
            b._check(b.lib.vkp_comm_allreduce_multi(self.gpu.gpu._ctx, ptrs, counts, n, _OPS[op],
                                                    float(np.float32(scale)), C.byref(job)))
            j = b.Job(job.value)
            for a in part:
                a.job = j
        return arrs

    def reduce_allreduce(self, arr, op: str, prev: int, axis: int, post: int, out_shape):
        """Local ``[prev, axis, post] -> [prev, post]`` reduction of ``arr`` fused with its exchange
        (``vkp_comm_reduce_allreduce``): the partials land in this rank's peer mailbox and one kernel
        folds every rank's partials over NVLink.  Returns the replicated result (job attached)."""
        import vulkpy_b200 as vk
        b = self._b
        out = vk.Array(self.gpu, shape=out_shape)
        job = C.c_void_p()
        b._check(b.lib.vkp_comm_reduce_allreduce(self.gpu.gpu._ctx, _OPS[op], arr.buffer.ptr, out.buffer.ptr,
                                                 int(prev), int(axis), int(post), C.byref(job)))
        out.job = b.Job(job.value)
        out._keep = [arr]
        return out

    def barrier(self):
        """Stream-ordered barrier over the ranks (mailbox flags over NVLink, else a one-word NCCL
        all-reduce); returns the job."""
        b = self._b
        job = C.c_void_p()
        b._check(b.lib.vkp_comm_barrier(self.gpu.gpu._ctx, C.byref(job)))
        return b.Job(job.value)

    def peer_mode(self, mode: int = -1) -> bool:
        """0: route every later collective through NCCL, 1: peer mailbox where it applies (default),
        -1: query.  Returns whether the mailbox path is mapped and selected.  Collective."""
        b = self._b
        active = C.c_int(0)
        b._check(b.lib.vkp_comm_peer_mode(self.gpu.gpu._ctx, int(mode), C.byref(active)))
        return bool(active.value)

    def allgather(self, arr, out):
        """``out`` (world * len(arr) elements) receives every rank's ``arr`` in rank order."""
        b = self._b
        job = C.c_void_p()
        b._check(b.lib.vkp_comm_allgather(self.gpu.gpu._ctx, arr.buffer.ptr, out.buffer.ptr,
                                          arr.buffer.nbytes, C.byref(job)))
        out.job = b.Job(job.value)
        out._keep = [arr]
        return out

    def matmul_allgather(self, a, b_shard, n_cols: int):
        """``a`` [M, K] local rows, ``b_shard`` [K/world, N] local rows of B -> local rows of A @ B.
        One tcgen05 GEMM that starts on the local K range while the copy engines pull the peers'
        shards over NVLink (``vkp_comm_matmul_allgather``).  Returns ``None`` when the shape does
        not fit that kernel (the caller then all-gathers B and multiplies)."""
        b = self._b
        M, K = a.shape
        kc = K // self.world
        if (self._fused_broken or K % self.world or kc % 32 or b_shard.shape[0] != kc or M < 128 or n_cols < 128
                or M % 4 or n_cols % 4):
            return None
        import vulkpy_b200 as vk
        out = vk.Array(self.gpu, shape=(M, n_cols))
        job = C.c_void_p()
        try:
            b._check(b.lib.vkp_comm_matmul_allgather(self.gpu.gpu._ctx, M, n_cols, K, a.buffer.ptr,
                                                     b_shard.buffer.ptr, out.buffer.ptr, C.byref(job)))
        except RuntimeError as e:      # e.g. CUDA IPC not permitted in this container: collective fallback
            if "cudaIpc" not in str(e):
                raise
            self._fused_broken = True
            return None
        out.job = b.Job(job.value)
        out._keep = [a, b_shard]
        return out

    _fused_broken = False

    def new_array(self, shape):
        import vulkpy_b200 as vk
        return vk.Array(self.gpu, shape=shape)

    def close(self):
        self._b.lib.vkp_comm_destroy(self.gpu.gpu._ctx)


def _exchange_unique_id(rank: int, world: int) -> bytes:
    """Rank 0 creates the NCCL id; everyone reads it from a TCP store on MASTER_ADDR (torchrun
    exports it) -- torch is plumbing here, never on the data path."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        obj = [NcclTransport.new_unique_id() if rank == 0 else None]
        td.broadcast_object_list(obj, src=0)
        return obj[0]
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + 23
    store = td.TCPStore(addr, port, world, rank == 0)
    if rank == 0:
        store.set("vkp_nccl_id", NcclTransport.new_unique_id())
    return store.get("vkp_nccl_id")


class Group:
    """The set of ranks a sharded array lives on."""

    def __init__(self, transport, rank: int, world: int, gpu=None):
        self.t, self.rank, self.world, self.gpu = transport, int(rank), int(world), gpu

    @classmethod
    def from_env(cls) -> "Group":
        """One process per GPU under torchrun: RANK / WORLD_SIZE / LOCAL_RANK from the environment."""
        import vulkpy_b200 as vk
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        gpu = vk.GPU(int(os.environ.get("LOCAL_RANK", "0")))
        uid = _exchange_unique_id(rank, world)
        return cls(NcclTransport(gpu, rank, world, uid), rank, world, gpu)

    def bounds(self, n: int) -> Tuple[int, int]:
        return shard_bounds(n, self.world, self.rank)

    # -- constructors ------------------------------------------------------------------------
    def shard(self, data) -> "ShardedArray":
        """Shard a full host array by rows; each rank uploads only its own block."""
        data = np.asarray(data)
        lo, hi = self.bounds(data.shape[0])
        import vulkpy_b200 as vk
        return ShardedArray(self, vk.Array(self.gpu, data=data[lo:hi]), data.shape)

    def wrap(self, local, global_shape) -> "ShardedArray":
        return ShardedArray(self, local, tuple(global_shape))

    def random(self, rng, shape: Sequence[int], kind: str = "random", **kw) -> "ShardedArray":
        """This rank's rows of ``rng.<kind>(shape=shape)``: same values as one GPU would produce.

        Every rank holds an identically seeded generator.  The rows of a rank are a contiguous
        range of the flat output, i.e. (when it starts and ends on a multiple of ``rng.size``)
        a range of chunks; the lanes jump ahead over the chunks of the lower ranks, draw, and
        jump over the rest so that all generators end in the single-GPU state."""
        shape = tuple(int(s) for s in shape)
        row = int(np.prod(shape[1:], dtype=np.int64))
        lo, hi = self.bounds(shape[0])
        size = rng.rng.size
        total = shape[0] * row
        if kind == "normal" and total % 2:
            raise ValueError("sharded normal() needs an even number of elements")
        if self.world > 1 and ((lo * row) % size or ((hi * row) % size and hi != shape[0])):
            raise ValueError("shard boundaries must fall on multiples of the generator's lane count")
        rng.rng.advance(lo * row)
        local = getattr(rng, kind)(shape=(hi - lo,) + shape[1:], **kw)
        if total - hi * row:
            rng.rng.advance(total - hi * row)
        return ShardedArray(self, local, shape)


def _is_sharded(x) -> bool:
    return isinstance(x, ShardedArray)


class ShardedArray:
    """Rows ``bounds(global_shape[0])`` of a global float32 array, held as one local array."""

    def __init__(self, group: Group, local, global_shape: Sequence[int]):
        self.group, self.local = group, local
        self.shape = tuple(int(s) for s in global_shape)
        lo, hi = group.bounds(self.shape[0])
        if tuple(local.shape) != (hi - lo,) + self.shape[1:]:
            raise ValueError(f"local shape {tuple(local.shape)} is not rows [{lo},{hi}) of {self.shape}")

    # -- element-wise: no exchange ---------------------------------------------------------------
    def _other(self, other):
        if _is_sharded(other):
            if other.shape[0] != self.shape[0]:
                raise ValueError(f"Incompatible shapes: {self.shape} vs {other.shape}")
            return other.local
        if hasattr(other, "shape") and len(other.shape) == len(self.shape) and other.shape[0] not in (1,):
            raise ValueError("an operand with a full leading axis must be a ShardedArray")
        return other            # scalar or replicated small operand broadcast along the rows

    def _like(self, local) -> "ShardedArray":
        return ShardedArray(self.group, local, (self.shape[0],) + tuple(local.shape[1:]))

    def __add__(self, o): return self._like(self.local + self._other(o))
    def __sub__(self, o): return self._like(self.local - self._other(o))
    def __mul__(self, o): return self._like(self.local * self._other(o))
    def __truediv__(self, o): return self._like(self.local / self._other(o))
    def __pow__(self, o): return self._like(self.local ** self._other(o))
    def __radd__(self, o): return self._like(o + self.local)
    def __rsub__(self, o): return self._like(o - self.local)
    def __rmul__(self, o): return self._like(o * self.local)
    def __rtruediv__(self, o): return self._like(o / self.local)
    def __rpow__(self, o): return self._like(o ** self.local)

    def __iadd__(self, o):
        self.local += self._other(o)
        return self

    def __isub__(self, o):
        self.local -= self._other(o)
        return self

    def __imul__(self, o):
        self.local *= self._other(o)
        return self

    def __itruediv__(self, o):
        self.local /= self._other(o)
        return self

    def max(self, o, inplace: bool = False):
        r = self.local.max(self._other(o), inplace=inplace)
        return self if inplace else self._like(r)

    def min(self, o, inplace: bool = False):
        r = self.local.min(self._other(o), inplace=inplace)
        return self if inplace else self._like(r)

    def clamp(self, min, max, inplace: bool = False):
        r = self.local.clamp(self._other(min), self._other(max), inplace=inplace)
        return self if inplace else self._like(r)

    def __getattr__(self, name):
        if name in _UNARY:
            def method(inplace: bool = False):
                r = getattr(self.local, name)(inplace=inplace)
                return self if inplace else self._like(r)
            return method
        raise AttributeError(name)

    def wait(self):
        self.local.wait()

    # -- reductions --------------------------------------------------------------------------------
    def _reduce0(self, name: str, local, keep_shape):
        """Reduce ``local`` over its leading axis and over the ranks: replicated ``[1, ...]`` /
        ``[...]`` array.  One fused call when the transport has it, else reduce + all-reduce."""
        fused = getattr(self.group.t, "reduce_allreduce", None)
        rest = tuple(local.shape[1:])
        if fused is not None:
            return fused(local, name, 1, int(local.shape[0]), int(np.prod(rest, dtype=np.int64)), keep_shape)
        part = getattr(local, name)(axis=0)
        part.reshape(keep_shape)
        return self.group.t.allreduce(part, name)

    def _reduce(self, name: str, axis, keepdims: bool, rebroadcast: bool):
        nd = len(self.shape)
        if rebroadcast:
            if not isinstance(axis, (int, np.integer)):
                raise ValueError("When `rebroadcast` is specified, `axis` must be `int`")
            a = axis % nd
            if a != 0:
                return self._like(getattr(self.local, name)(axis=a, rebroadcast=True))
            part = self._reduce0(name, self.local, (1,) + tuple(self.local.shape[1:]))     # [1, ...]
            return self._like(part.broadcast_to(self.local.shape))
        if axis is None:
            fused = getattr(self.group.t, "reduce_allreduce", None)
            n_local = int(np.prod(self.local.shape, dtype=np.int64))
            if fused is not None:
                part = fused(self.local, name, 1, n_local, 1, (1,))               # one float per rank
            else:
                part = getattr(self.local, name)()                                # (1,) partial
                self.group.t.allreduce(part, name)
            if keepdims:
                part.reshape((1,) * nd)
            return part                                                        # replicated
        axes = sorted({int(a) % nd for a in np.asarray(axis).reshape(-1)})
        if 0 not in axes:
            local = getattr(self.local, name)(axis=axes, keepdims=keepdims)
            return ShardedArray(self.group, local, (self.shape[0],) + tuple(local.shape[1:]))
        inner = [a for a in axes if a != 0]
        local = getattr(self.local, name)(axis=inner, keepdims=True) if inner else self.local
        shape = [1 if a in axes else s for a, s in enumerate(self.shape)]          # keepdims form
        out = self._reduce0(name, local, tuple(shape))                             # `post` floats
        if not keepdims:
            out.reshape(tuple(s for a, s in enumerate(shape) if a not in axes) or (1,))
        return out                                                             # replicated

    def sum(self, axis=None, keepdims=False, rebroadcast=False): return self._reduce("sum", axis, keepdims, rebroadcast)
    def prod(self, axis=None, keepdims=False, rebroadcast=False): return self._reduce("prod", axis, keepdims, rebroadcast)
    def maximum(self, axis=None, keepdims=False, rebroadcast=False): return self._reduce("maximum", axis, keepdims, rebroadcast)
    def minimum(self, axis=None, keepdims=False, rebroadcast=False): return self._reduce("minimum", axis, keepdims, rebroadcast)

    def mean(self, axis=None, keepdims=False, rebroadcast=False):
        """Global sum, then the reference's single rescale with GLOBAL element counts
        (vulkpy/vkarray.py:1420-1432)."""
        n_before = int(np.prod(self.shape, dtype=np.int64))
        ret = self.sum(axis, keepdims, rebroadcast)
        if rebroadcast:
            ret /= self.shape[axis % len(self.shape)]
            return ret
        n_after = int(np.prod(ret.shape, dtype=np.int64))
        ret *= (n_after / n_before)
        return ret

    # -- contraction / gather ----------------------------------------------------------------------------
    def __matmul__(self, other):
        """Row-sharded ``A @ B``.  ``B`` replicated: local GEMM.  ``B`` sharded by its rows (= K):
        all-gather B first, then the local GEMM; the result is row-sharded like ``A``."""
        if _is_sharded(other):
            if other.shape[0] != self.shape[-1]:
                raise ValueError(f"Incompatible shapes: {self.shape} vs {other.shape}")
            fused = getattr(self.group.t, "matmul_allgather", None)
            if fused is not None and len(self.local.shape) == 2 and len(other.shape) == 2:
                out = fused(self.local, other.local, int(other.shape[1]))
                if out is not None:
                    return self._like(out)
            other = other.allgather()
        return self._like(self.local @ other)

    def gather(self, indices, axis=None) -> "ShardedArray":
        """``self`` is a REPLICATED-table view is not needed: see :func:`gather_replicated`."""
        raise TypeError("gather on a sharded table needs an exchange; replicate the table and use "
                        "dist.gather_replicated(group, table, local_indices, n_global)")

    def allgather(self):
        """Replicated local array with the full global contents (equal shards only for NCCL)."""
        g = self.group
        rows = [shard_bounds(self.shape[0], g.world, r) for r in range(g.world)]
        if len({hi - lo for lo, hi in rows}) != 1:
            raise ValueError("allgather needs equal shards (leading axis divisible by the world size)")
        out = g.t.new_array(self.shape)
        return g.t.allgather(self.local, out)

    def to_numpy(self) -> np.ndarray:
        return np.asarray(self.allgather()).reshape(self.shape).copy()


def gather_replicated(group: Group, table, local_indices, n_global: int) -> ShardedArray:
    """Flat gather with the table replicated on every rank and the uint32 indices sharded by rows:
    purely local; the result is sharded like the indices."""
    out = table.gather(local_indices)
    return ShardedArray(group, out, (int(n_global),) + tuple(out.shape[1:]))


_UNARY = ("abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
          "asinh", "acosh", "atanh", "exp", "log", "exp2", "log2", "sqrt", "invsqrt")


class DataParallel:
    """Data-parallel wrapper of ``nn.Sequence``: parameters and optimizer state replicated, batch
    sharded; after the local backward pass every gradient is all-reduced (sum) and scaled by
    1/world so that ``reduce="mean"`` losses keep their meaning (vulkpy/nn/losses.py:41-46); the
    identical optimizer step then runs on every replica."""

    def __init__(self, net, group: Group):
        self.net, self.group = net, group

    def parameters(self) -> List:
        ps = []
        for layer in self.net.L:
            for name in ("w", "b"):
                p = getattr(layer, name, None)
                if p is not None and getattr(p, "grad", None) is not None:
                    ps.append(p)
        return ps

    def train(self, x, y):
        net, g = self.net, self.group
        forward_loss = getattr(net, "_forward_loss", None)      # Sequence: the fused Softmax + CrossEntropy tail
        if forward_loss is not None:
            pred, loss = forward_loss(x, y)
        else:
            pred = net._forward(x)
            loss = net.loss(pred, y)
        net._zero_grad()
        net._backward()
        inv = 1.0 / g.world
        bucket = [p.grad for p in self.parameters()] + [loss]
        many = getattr(g.t, "allreduce_many", None)
        if many is not None:            # one grouped exchange + one scale kernel for the whole bucket
            many(bucket, "sum", inv)
        else:
            for t in bucket:
                g.t.allreduce(t, "sum")
                t *= inv
        net._update()
        return pred, loss

    def predict(self, x, y=None):
        return self.net.predict(x, y)
