"""
Array core of the B200 backend (:mod:`vulkpy_b200.vkarray`)

Same public surface as the reference's ``vulkpy.vkarray`` (reference: vulkpy/vkarray.py):
``GPU``, ``Array`` (float32), ``U32Array`` / ``Shape`` (uint32) and ``zeros``.  Every operator
allocates its result, enqueues ONE kernel on the device's in-order stream through
``GPU._submit`` and returns immediately; the result carries the producing ``job`` and keeps
its inputs alive in ``_keep`` until ``wait()`` (reference job model: vkarray.py:139-161).

Differences that are supersets of the reference behaviour:
in-place broadcasting follows NumPy for any rank (Q1), full reductions are correct for any
size (Q2), ``__setitem__`` waits for pending work first (Q5) and ``a[:] = scalar`` is a device
fill instead of a host loop (Q17).
"""
from __future__ import annotations

import contextlib
import logging
import math
import operator
import os
import weakref
from typing import Iterable, List, Optional, Tuple, Union

import numpy as np

from . import _backend as _b
from ._backend import (  # re-exported: nn code imports these from here (nn/layers.py:9)
    createGPU, DataShape, Job, Buffer, Shape as U32Buffer,
    VectorParams, MultiVector2Params, VectorScalarParams, VectorScalar2Params, MatMulParams,
    AxisReductionParams, BroadcastParams, Multi3BroadcastParams, BatchAffineParams,
    AxisGatherParams, VectorRangeParams,
)
from .vktyping import Resource

__all__ = ["GPU", "U32Array", "Shape", "Array", "zeros", "fuse"]

logger = logging.getLogger("vulkpy")

# host reads of at least this many bytes migrate in one bulk prefetch instead of page faults
_PREFETCH_BYTES = 64 * 1024


# op id -> index of the first shape binding (add_broadcast.comp binding 3, iadd_broadcast.comp
# binding 2, broadcast.comp bindings 2-3)
_SHAPE_BINDING = {}
for _name, _id in list(_b.OPS.items()):
    if _name == "broadcast" or (_name.endswith("_broadcast") and _name.startswith("i")):
        _SHAPE_BINDING[_id] = 2
    elif _name.endswith("_broadcast"):
        _SHAPE_BINDING[_id] = 3


# ---- lazy element-wise fusion -------------------------------------------------
# Inside `with vk.fuse():` (or with VULKPY_FUSE=1) same-shape / scalar element-wise operators do not
# launch: the result carries a recorded chain over its concrete inputs and ONE `vkp_ew_chain` launch
# evaluates it when something needs the values (any other operation, wait(), a host view).  Every
# step is the reference shader's own operation with its own rounding, so results are bit-identical
# to the op-by-op sequence; intermediates nobody looks at are never written to memory.
_CHAIN_BIN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "max": 4, "min": 5, "pow": 6, "rsub": 7, "rdiv": 8, "rpow": 9}
_CHAIN_FLIP = {0: 0, 2: 2, 4: 4, 5: 5, 1: 7, 3: 8, 6: 9}     # concrete (op) lazy == lazy (flipped op) concrete
_CHAIN_UNARY0 = 11
_CHAIN_SAVE = 31
_CHAIN_MAX_STEPS = 16
_CHAIN_MAX_INPUTS = 4
_fuse_depth = 1 if os.environ.get("VULKPY_FUSE", "0") == "1" else 0


@contextlib.contextmanager
def fuse():
    """Record same-shape element-wise operators instead of launching them one by one; chains are
    evaluated by one kernel each on first use.  Additive (the reference has no counterpart); do not
    write through a NumPy view obtained BEFORE an operation was recorded on that array."""
    global _fuse_depth
    _fuse_depth += 1
    try:
        yield
    finally:
        _fuse_depth -= 1


class _Lazy:
    """Recorded chain: inputs[0] is the running value's source, steps = (op, src, scalar)."""
    __slots__ = ("inputs", "steps")

    def __init__(self, inputs, steps):
        self.inputs, self.steps = inputs, steps

    def extended(self, op: int, other):
        """New chain = this one + (op, other); None when it would not fit one launch."""
        if len(self.steps) >= _CHAIN_MAX_STEPS:
            return None
        inputs = self.inputs
        if other is None:
            step = (op, 0, 0.0)
        elif isinstance(other, _GPUArray):
            for k, a in enumerate(inputs):
                if a is other and k > 0:
                    break
            else:
                if len(inputs) >= _CHAIN_MAX_INPUTS:
                    return None
                inputs = inputs + [other]
                k = len(inputs) - 1
            step = (op, k, 0.0)
        else:
            step = (op, 0, float(other))
        return _Lazy(inputs, self.steps + [step])


def _full_slice(key) -> bool:
    if key is Ellipsis:
        return True
    if isinstance(key, slice):
        return key.start is None and key.stop is None and key.step is None
    if isinstance(key, tuple):
        return all(_full_slice(k) for k in key)
    return False


def _prod(shape) -> int:
    """Product of a shape tuple as a Python int (math.prod: ~0.1 us; np.prod costs ~5 us per call,
    which at ~40 small launches per nn step was a quarter of the host time)."""
    return math.prod(shape)


def _as_shape(shape) -> Tuple[int, ...]:
    """Shape argument (int, sequence of ints, or integer ndarray) as a tuple of Python ints."""
    if isinstance(shape, (tuple, list)):
        try:
            return tuple(operator.index(s) for s in shape)
        except TypeError:
            pass
    return tuple(int(s) for s in np.asarray(shape, dtype=int).reshape(-1))

class GPU:
    """
    One B200 (reference: vkarray.py:68-137).

    ``GPU(idx)`` selects CUDA device ``idx``; objects with the same index compare equal and
    share one context / stream.
    """

    def __init__(self, idx: int = 0, priority: float = 0.0):
        self._idx = idx
        self.gpu = createGPU(idx, priority)
        self.canSubgroupArithmetic = self.gpu.canSubgroupArithmetic()
        logger.info("GPU %d: CUDA sm_100a backend, %d SMs", idx, self.gpu.sm_count())

    def __eq__(self, other: object):
        if not isinstance(other, GPU):
            return NotImplemented
        return self._idx == other._idx

    def __hash__(self):
        return hash(("vulkpy.GPU", self._idx))

    def _submit(self, spv, local_size_x: int, local_size_y: int, local_size_z: int,
                arrays: Iterable, shape, params) -> Job:
        """Enqueue kernel ``spv`` over ``arrays`` (reference: vkarray.py:110-119).

        Dependencies need no host-side wait: the stream is in order.  Bindings may be
        ``Array``/``U32Array`` objects or, for broadcast shape bindings, host uint32 arrays.
        """
        infos = []
        for a in arrays:
            if isinstance(a, _GPUArray):
                infos.append(a.buffer)
            elif isinstance(a, np.ndarray):
                infos.append(a)
            else:
                infos.append(a._info())
        op = spv if isinstance(spv, int) else _b.op_id(spv)
        first_shape = _SHAPE_BINDING.get(op)
        if first_shape is not None:
            # shape bindings are consumed on the host at submit time; accept Shape arrays too
            arrays = list(arrays)
            for i in range(first_shape, len(infos)):
                if isinstance(infos[i], _b._BufferBase):
                    arrays[i].wait()
                    infos[i] = np.array(arrays[i].array, dtype=np.uint32).reshape(-1)
        return self.gpu.submit(op, local_size_x, local_size_y, local_size_z, infos, shape, params)

    def flush(self, arrays: Iterable["_GPUArray"]):
        """No-op: host writes to managed memory are coherent (reference: vkarray.py:121-130)."""
        self.gpu.flush([a.buffer.range() for a in arrays])

    def wait(self):
        """Wait for everything enqueued on this GPU."""
        self.gpu.wait()


class _GPUArray(Resource):
    """Job / keep-alive / host-view plumbing shared by Array and U32Array."""
    _np_dtype = np.float32

    def __init__(self, gpu: GPU):
        self._gpu: GPU = gpu
        self.job: Optional[Job] = None
        self._keep: List[Resource] = []
        self._view: Optional[np.ndarray] = None
        self._buffer = None
        self._lazy: Optional[_Lazy] = None        # recorded, not yet launched chain (vk.fuse())
        self._consumers = None                    # weak set of lazy arrays that read this one

    # The device buffer.  Reading it is what "needs the values": a recorded chain is launched first,
    # and recorded chains that READ this array are launched before anyone can overwrite it.
    @property
    def buffer(self):
        if self._consumers:
            self._flush_consumers()
        if self._lazy is not None:
            self._materialize()
        return self._buffer

    @buffer.setter
    def buffer(self, b):
        self._buffer = b

    def _flush_consumers(self):
        cons, self._consumers = self._consumers, None
        for c in list(cons):
            if c._lazy is not None:
                c._materialize()

    def _materialize(self):
        lz, self._lazy = self._lazy, None
        dev = self._gpu.gpu
        src0 = lz.inputs[0]
        for a in lz.inputs:
            if a._lazy is not None:
                a._materialize()
        # an in-place chain (`a += 1; a.exp(inplace=True)`) writes back into its own source buffer;
        # chains recorded from the value it is about to overwrite go first
        if src0 is self._inplace_src and src0._consumers:
            src0._flush_consumers()
        out = src0._buffer if src0 is self._inplace_src else self._create(dev, _prod(self.shape))
        self.job = dev.ew_chain([a._buffer for a in lz.inputs], out, [st[0] for st in lz.steps],
                                [st[1] for st in lz.steps], [st[2] for st in lz.steps])
        self._buffer = out
        self._keep = [a for a in lz.inputs if a is not self._inplace_src]
        self._inplace_src = None
        for a in lz.inputs:
            if a._consumers is not None:
                a._consumers.discard(self)

    _inplace_src = None

    def _watch(self, lz: _Lazy):
        """Register this (lazy) array with the inputs of its chain."""
        for a in lz.inputs:
            if a is self._inplace_src:
                continue
            if a._consumers is None:
                a._consumers = weakref.WeakSet()
            a._consumers.add(self)

    def _alloc(self, data, shape):
        dev = self._gpu.gpu
        if data is not None:
            host = np.asarray(data)
            self.shape = host.shape
            host = np.ascontiguousarray(host, dtype=self._np_dtype)
            self.buffer = self._create(dev, host.size)
            if host.size:
                self.buffer.upload(host.reshape(-1))
        else:
            if shape is None:
                raise ValueError("`data` or `shape` must not be `None`.")
            self.shape = _as_shape(shape)
            self.buffer = self._create(dev, _prod(self.shape))

    def wait(self):
        """Wait for the job that writes this array."""
        if self._lazy is not None:
            self._materialize()
        job = self.job
        if job is not None:
            job.wait()
            self.job = None
        self._keep = []

    def flush(self):
        self._gpu.flush([self])

    def _info(self):
        return self.buffer.info()

    # -- host view --------------------------------------------------------------------------
    def _host(self, prefetch: bool, write: bool = False) -> np.ndarray:
        self.buffer.host_acquire(prefetch and self.buffer.nbytes >= _PREFETCH_BYTES, write)
        v = self._view
        if v is None:
            v = np.asarray(self.buffer)
            v.shape = self.shape
            self._view = v
        return v

    @property
    def array(self) -> np.ndarray:
        """Live NumPy view of the buffer (reference attribute ``array``: vkarray.py:430-431)."""
        return self._host(False)

    def __getitem__(self, key):
        self.wait()
        return self._host(True)[key]

    def __setitem__(self, key, value):
        if not isinstance(value, _GPUArray) and _full_slice(key) and np.ndim(value) == 0:
            # whole-array scalar assignment: device fill, ordered on the stream
            bits = int(np.asarray(value, dtype=self._np_dtype).view(np.uint32))
            self.job = self._gpu.gpu.fill(self.buffer, bits)
            self._keep = []
            return
        self.wait()
        self._host(False, write=True)[key] = value

    def __repr__(self) -> str:
        return f"<{self.__class__.__name__}(shape={tuple(self.shape)})>"

    def __str__(self) -> str:
        self.wait()
        return str(self._host(True))

    def __array__(self, dtype=None, copy=None) -> np.ndarray:
        self.wait()
        v = self._host(True)
        if dtype is not None and np.dtype(dtype) != v.dtype:
            return v.astype(dtype)
        return v.copy() if copy else v

    def to_host(self, out: Optional[np.ndarray] = None, wait: bool = True) -> np.ndarray:
        """Copy the contents into host memory with one stream-ordered device-to-host transfer
        (additive helper: unlike ``np.asarray(array)`` no managed page migrates, so the array
        stays resident in HBM; pass a ``pinned_empty`` buffer as ``out`` for full PCIe rate).

        ``wait=False`` (``out`` must be a ``pinned_empty`` array) runs the transfer on the
        device-to-host copy engine and returns at once; ``self.wait()`` completes it.  Uploads
        of the next step (``from_host``) and kernels overlap it."""
        if out is None:
            assert wait, "an asynchronous download needs a pinned `out`"
            out = np.empty(self.shape, dtype=self._np_dtype)
        assert out.nbytes == self.buffer.nbytes and out.flags.c_contiguous
        if not wait:
            self.job = self.buffer.download_async(out.reshape(-1))
            self._keep = [out]
            return out
        _b._check(_b.lib.vkp_download(self._gpu.gpu._ctx, out.ctypes.data, self.buffer.ptr, out.nbytes))
        self.job = None
        self._keep = []
        return out

    @classmethod
    def from_host(cls, gpu: "GPU", pinned: np.ndarray):
        """Array filled from a ``pinned_empty`` host array by the host-to-device copy engine.

        Additive helper for pipelined steps: returns at once, the transfer overlaps kernels and
        downloads already enqueued, operations on the result are ordered after it.  Unlike
        ``Array(gpu, data=...)`` the source is read asynchronously: do not rewrite ``pinned``
        before ``wait()`` on the result (or on anything computed from it) has returned."""
        self = cls.__new__(cls)
        _GPUArray.__init__(self, gpu)
        host = np.asarray(pinned)
        assert host.dtype == cls._np_dtype and host.flags.c_contiguous
        self.shape = host.shape
        self.buffer = (_b.Shape if cls._np_dtype == np.uint32 else _b.Buffer)(gpu.gpu, host.size, for_upload=True)
        self.job = self.buffer.upload_async(host.reshape(-1))
        self._keep = [host]
        return self

    def _set_shape(self, shape):
        shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
        n = _prod(self.shape) if self._lazy is not None else self.buffer.size()
        if shape.count(-1) == 1:
            rest = -_prod(shape)
            if rest > 0 and n % rest == 0:
                shape = tuple(n // rest if s == -1 else s for s in shape)
        if any(s < 0 for s in shape) or _prod(shape) != n:
            raise ValueError(f"cannot reshape array of size {n} into shape {shape}")
        self.shape = shape
        if self._view is not None:
            self._view = self._view.reshape(shape)


class U32Array(_GPUArray):
    """uint32 array for indices / labels / shapes (reference: vkarray.py:191-257)."""
    _np_dtype = np.uint32

    def __init__(self, gpu: GPU, *, data: Optional[Iterable[int]] = None,
                 shape: Optional[Iterable[int]] = None):
        super().__init__(gpu)
        if data is None and shape is None:
            raise ValueError("One of `data` or `shape` must be specified.")
        self._alloc(data, shape)

    @staticmethod
    def _create(dev, n):
        return dev.createU32Buffer(n)

    def to_onehot(self, num_classes: int) -> "Array":
        """Rows of the identity selected by the labels (reference: vkarray.py:239-253)."""
        return Array(self._gpu, data=np.identity(num_classes, dtype=np.float32)).gather(self, axis=0)


class Shape(U32Array):
    """uint32 vector holding a shape (reference: vkarray.py:259-278)."""

    def __init__(self, gpu: GPU, *, data: Optional[Iterable[int]] = None, ndim: Optional[int] = None):
        super().__init__(gpu, data=data, shape=(ndim,) if ndim is not None else None)


_BINARY = ("add", "sub", "mul", "div", "max", "min", "pow")
_UNARY = ("abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh",
          "asinh", "acosh", "atanh", "exp", "log", "exp2", "log2", "sqrt", "invsqrt")
_REDUCE = ("sum", "prod", "maximum", "minimum")

Scalar = Union[int, float]


class Array(_GPUArray):
    """float32 device array (reference: vkarray.py:281-1521)."""

    def __init__(self, gpu: GPU, *, data=None, shape: Optional[Iterable[int]] = None):
        super().__init__(gpu)
        self._alloc(data, shape)

    @staticmethod
    def _create(dev, n):
        return dev.createBuffer(n)

    # -- submit helpers -----------------------------------------------------------------------
    def _check_shape(self, other):
        if tuple(self.shape) != tuple(other.shape):
            raise ValueError(f"Incompatible shapes: {self.shape} vs {other.shape}")

    def _new(self, shape=None) -> "Array":
        return Array(self._gpu, shape=self.shape if shape is None else shape)

    def _run(self, ret: "Array", spv: str, arrays, params, keep) -> "Array":
        n = ret.buffer.size()
        ret.job = self._gpu._submit(spv, 64, 1, 1, arrays, DataShape(n, 1, 1), params)
        ret._keep = keep
        return ret

    # -- lazy element-wise fusion ---------------------------------------------------------------------
    def _shell(self) -> "Array":
        """A second handle on this array's current buffer (the source of an in-place chain)."""
        sh = Array.__new__(Array)
        _GPUArray.__init__(sh, self._gpu)
        sh.shape = self.shape
        sh._buffer, sh.job, sh._keep = self._buffer, self.job, self._keep
        return sh

    def _record(self, op: int, other, inplace: bool) -> Optional["Array"]:
        """Record `self (op) other` instead of launching it; None when it cannot join a chain."""
        if isinstance(other, _GPUArray):
            if type(other) is not Array or tuple(other.shape) != tuple(self.shape) or other._gpu is not self._gpu:
                return None
            if other is self and inplace:
                return None
        if _prod(self.shape) == 0:
            return None
        if (not inplace and self._lazy is None and isinstance(other, Array) and other._lazy is not None
                and op in _CHAIN_FLIP and other is not self):
            return other._record(_CHAIN_FLIP[op], self, False)      # continue the other operand's chain
        if inplace and self._consumers:          # recorded readers of the OLD value go first
            self._flush_consumers()
        if self._lazy is not None:
            lz = self._lazy.extended(op, other)
            if lz is None:                       # chain full: launch it, start a new one on its result
                self._materialize()
        if self._lazy is None:
            if inplace:
                src = self._shell()
                lz = _Lazy([src], []).extended(op, other)
            else:
                lz = _Lazy([self], []).extended(op, other)
        if inplace:
            if self._lazy is None:
                self._inplace_src = src
                self._buffer, self.job, self._keep, self._view = None, None, [], None
            self._lazy = lz
            self._watch(lz)
            return self
        ret = Array.__new__(Array)
        _GPUArray.__init__(ret, self._gpu)
        ret.shape = tuple(self.shape)
        ret._lazy = lz
        ret._watch(lz)
        return ret

    def _op(self, other, name: str) -> "Array":
        """Out-of-place binary op: same shape, scalar or broadcast (reference: vkarray.py:492-519)."""
        if _fuse_depth > 0:
            r = self._record(_CHAIN_BIN[name], other, False)
            if r is not None:
                return r
        n = self.buffer.size()
        if not isinstance(other, Array):
            ret = self._new()
            return self._run(ret, name + "_scalar", [self, ret], VectorScalarParams(n, float(other)), [self])
        if tuple(self.shape) == tuple(other.shape):
            ret = self._new()
            return self._run(ret, name, [self, other, ret], VectorParams(n), [self, other])
        shape = np.broadcast_shapes(self.shape, other.shape)  # ValueError if incompatible
        ndim = len(shape)
        sh = np.ones(3 * ndim, dtype=np.uint32)
        sh[ndim - len(self.shape):ndim] = self.shape
        sh[2 * ndim - len(other.shape):2 * ndim] = other.shape
        sh[2 * ndim:] = shape
        ret = self._new(shape)
        p = Multi3BroadcastParams(n, other.buffer.size(), ret.buffer.size(), ndim)
        return self._run(ret, name + "_broadcast", [self, other, ret, sh], p, [self, other])

    def _iop(self, other, name: str) -> "Array":
        """In-place binary op (reference: vkarray.py:533-559; NumPy rules for any rank, Q1)."""
        if _fuse_depth > 0:
            r = self._record(_CHAIN_BIN[name], other, True)
            if r is not None:
                return r
        n = self.buffer.size()
        if not isinstance(other, Array):
            return self._run(self, "i" + name + "_scalar", [self], VectorScalarParams(n, float(other)), [])
        if tuple(self.shape) == tuple(other.shape):
            return self._run(self, "i" + name, [self, other], VectorParams(n), [other])
        shape = np.broadcast_shapes(self.shape, other.shape)
        if tuple(shape) != tuple(self.shape):
            raise ValueError(f"Incompatible shape. {shape} vs {self.shape}")
        ndim = len(shape)
        sh = np.ones(2 * ndim, dtype=np.uint32)
        sh[:ndim] = shape
        if len(other.shape) > 0:
            sh[2 * ndim - len(other.shape):] = other.shape
        p = BroadcastParams(n, other.buffer.size(), ndim)
        return self._run(self, "i" + name + "_broadcast", [self, other, sh], p, [other])

    def _rop(self, other: Scalar, name: str) -> "Array":
        if _fuse_depth > 0:
            r = self._record(_CHAIN_BIN[name], other, False)
            if r is not None:
                return r
        ret = self._new()
        return self._run(ret, name + "_scalar", [self, ret], VectorScalarParams(self.buffer.size(), float(other)),
                         [self])

    def _unary(self, name: str, inplace: bool) -> "Array":
        if _fuse_depth > 0:
            r = self._record(_CHAIN_UNARY0 + _UNARY.index(name), None, inplace)
            if r is not None:
                return r
        p = VectorParams(self.buffer.size())
        if inplace:
            return self._run(self, "i" + name, [self], p, [])
        ret = self._new()
        return self._run(ret, name, [self, ret], p, [self])

    # -- arithmetic operators (reference: vkarray.py:521-583, 1100-1108) -------------------------
    def __add__(self, other): return self._op(other, "add")
    def __sub__(self, other): return self._op(other, "sub")
    def __mul__(self, other): return self._op(other, "mul")
    def __truediv__(self, other): return self._op(other, "div")
    def __pow__(self, other): return self._op(other, "pow")
    def __iadd__(self, other): return self._iop(other, "add")
    def __isub__(self, other): return self._iop(other, "sub")
    def __imul__(self, other): return self._iop(other, "mul")
    def __itruediv__(self, other): return self._iop(other, "div")
    def __ipow__(self, other): return self._iop(other, "pow")
    def __radd__(self, other): return self._rop(other, "add")
    def __rsub__(self, other): return self._rop(other, "rsub")
    def __rmul__(self, other): return self._rop(other, "mul")
    def __rtruediv__(self, other): return self._rop(other, "rdiv")
    def __rpow__(self, other): return self._rop(other, "rpow")

    def max(self, other: Union["Array", float], inplace: bool = False) -> "Array":
        """Element-wise maximum with an array (broadcast) or a scalar."""
        return self._iop(other, "max") if inplace else self._op(other, "max")

    def min(self, other: Union["Array", float], inplace: bool = False) -> "Array":
        """Element-wise minimum with an array (broadcast) or a scalar."""
        return self._iop(other, "min") if inplace else self._op(other, "min")

    def __matmul__(self, other: "Array") -> "Array":
        """1-D / 2-D matrix product, fp32 (reference: vkarray.py:585-605)."""
        if len(self.shape) > 3 or len(other.shape) > 3 or self.shape[-1] != other.shape[0]:
            raise ValueError(f"Incompatible shapes: {self.shape} vs {other.shape}")
        shape = tuple(self.shape)[:-1] + tuple(other.shape)[1:]
        if len(shape) == 0:
            shape = (1,)
        rowA = self.shape[0] if len(self.shape) > 1 else 1
        contract = self.shape[-1]
        colB = other.shape[1] if len(other.shape) > 1 else 1
        ret = self._new(shape)
        ret.job = self._gpu._submit("matmul", 1, 64, 1, [self, other, ret], DataShape(rowA, colB, 1),
                                    MatMulParams(rowA, contract, colB))
        ret._keep = [self, other]
        return ret

    def reshape(self, shape: Iterable[int]):
        """Reshape in place; ``ValueError`` if the size does not match (reference: vkarray.py:607-622)."""
        self._set_shape(shape)

    # -- clamp (reference: vkarray.py:1110-1191) ---------------------------------------------------
    def clamp(self, min: Union["Array", float], max: Union["Array", float], inplace: bool = False) -> "Array":
        """``min(max(x, lo), hi)`` with array (broadcast) or scalar bounds."""
        lo, hi = min, max
        lo_arr, hi_arr = isinstance(lo, Array), isinstance(hi, Array)
        src = self
        if lo_arr or hi_arr:
            shapes = [self.shape] + ([lo.shape] if lo_arr else []) + ([hi.shape] if hi_arr else [])
            shape = tuple(np.broadcast_shapes(*shapes))
            if shape != tuple(self.shape):
                if inplace:
                    raise ValueError("Incompatible shape")
                src = self.broadcast_to(shape)
            if lo_arr and tuple(lo.shape) != shape:
                lo = lo.broadcast_to(shape)
            if hi_arr and tuple(hi.shape) != shape:
                hi = hi.broadcast_to(shape)
        n = src.buffer.size()
        ret = self if inplace else src._new()
        tail = [] if inplace else [ret]
        pre = "i" if inplace else ""
        keep = [] if inplace else [src]
        if lo_arr and hi_arr:
            return src._run(ret, pre + "clamp", [src, lo, hi] + tail, VectorParams(n), keep + [lo, hi])
        if hi_arr:
            return src._run(ret, pre + "clamp_sv", [src, hi] + tail, VectorScalarParams(n, float(lo)), keep + [hi])
        if lo_arr:
            return src._run(ret, pre + "clamp_vs", [src, lo] + tail, VectorScalarParams(n, float(hi)), keep + [lo])
        return src._run(ret, pre + "clamp_ss", [src] + tail, VectorScalar2Params(n, float(lo), float(hi)), keep)

    # -- reductions (reference: vkarray.py:1194-1432) -----------------------------------------------
    def _norm_axis(self, axis) -> List[int]:
        nd = len(self.shape)
        if isinstance(axis, int):          # the common case without a round trip through NumPy
            ax = (axis,)
        else:
            ax = np.unique(np.asarray(axis, dtype=int).reshape(-1))
        out = sorted({int(a) + nd if a < 0 else int(a) for a in ax}, reverse=True)
        for a in out:
            if not 0 <= a < nd:
                raise ValueError(f"axis {a} is out of bounds for array of dimension {nd}")
        return out

    def _reduce(self, name: str, axis, keepdims: bool, rebroadcast: bool) -> "Array":
        if rebroadcast:
            if not isinstance(axis, (int, np.integer)):
                raise ValueError("When `rebroadcast` is specified, `axis` must be `int`")
            (a,) = self._norm_axis(axis)
            prev = _prod(self.shape[:a])
            post = _prod(self.shape[a + 1:])
            ret = self._new()
            ret.job = self._gpu._submit(name + "_axis_rebroadcast", 1, 64, 1, [self, ret],
                                        DataShape(prev, post, 1),
                                        AxisReductionParams(prev, int(self.shape[a]), post))
            ret._keep = [self]
            return ret
        if axis is None:
            # one launch pair replaces the reference's log64(n) dependent jobs (vkarray.py:1246-1274)
            ret = self._new((1,))
            n = self.buffer.size()
            ret.job = self._gpu._submit(name, 64, 1, 1, [self, ret], DataShape(64, 1, 1),
                                        MultiVector2Params(n, 1))
            ret._keep = [self]
            if keepdims:
                ret.reshape((1,) * len(self.shape))
            return ret
        axes = self._norm_axis(axis)
        tmp = self
        for a in axes:  # descending, one pass per axis like the reference (vkarray.py:1194-1222)
            prev = _prod(tmp.shape[:a])
            post = _prod(tmp.shape[a + 1:])
            ret = self._new(tuple(tmp.shape[:a]) + tuple(tmp.shape[a + 1:]))
            ret.job = self._gpu._submit(name + "_axis", 1, 64, 1, [tmp, ret], DataShape(prev, post, 1),
                                        AxisReductionParams(prev, int(tmp.shape[a]), post))
            ret._keep = [tmp]
            tmp = ret
        if keepdims:
            shape = list(self.shape)
            for a in axes:
                shape[a] = 1
            tmp.reshape(shape)
        return tmp

    def sum(self, axis=None, keepdims: bool = False, rebroadcast: bool = False) -> "Array":
        """Sum over all elements, one axis or several axes."""
        return self._reduce("sum", axis, keepdims, rebroadcast)

    def prod(self, axis=None, keepdims: bool = False, rebroadcast: bool = False) -> "Array":
        """Product over all elements, one axis or several axes."""
        return self._reduce("prod", axis, keepdims, rebroadcast)

    def maximum(self, axis=None, keepdims: bool = False, rebroadcast: bool = False) -> "Array":
        """Maximum over all elements, one axis or several axes."""
        return self._reduce("maximum", axis, keepdims, rebroadcast)

    def minimum(self, axis=None, keepdims: bool = False, rebroadcast: bool = False) -> "Array":
        """Minimum over all elements, one axis or several axes."""
        return self._reduce("minimum", axis, keepdims, rebroadcast)

    def mean(self, axis=None, keepdims: bool = False, rebroadcast: bool = False) -> "Array":
        """Mean = sum, then one scalar rescale exactly as the reference does (vkarray.py:1420-1432)."""
        n_before = self.buffer.size()
        ret = self.sum(axis, keepdims, rebroadcast)
        if rebroadcast:
            (a,) = self._norm_axis(axis)
            ret /= self.shape[a]
        else:
            ret *= (ret.buffer.size() / n_before)
        return ret

    # -- argmax / argmin: listed as missing by the reference (README.md:73; example/02-nn.py:96 uses
    #    np.argmax on the host); NumPy semantics, uint32 indices ------------------------------------
    def _argreduce(self, op: int, axis: Optional[int]) -> U32Array:
        if self.buffer.size() == 0:
            raise ValueError("attempt to get argmax of an empty sequence")
        if axis is None:
            prev, n, post, shape = 1, self.buffer.size(), 1, (1,)
        else:
            (a,) = self._norm_axis(axis)
            prev = _prod(self.shape[:a])
            post = _prod(self.shape[a + 1:])
            n, shape = int(self.shape[a]), tuple(self.shape[:a]) + tuple(self.shape[a + 1:])
        ret = U32Array(self._gpu, shape=shape if len(shape) else (1,))
        ret.job = self._gpu.gpu.argreduce(op, self.buffer, ret.buffer, prev, n, post)
        ret._keep = [self]
        return ret

    def argmax(self, axis: Optional[int] = None) -> U32Array:
        """Index of the first maximum (NaN counts as maximum) over everything (flat index, shape
        ``(1,)``) or along ``axis`` -- ``np.argmax`` as ``uint32``."""
        return self._argreduce(0, axis)

    def argmin(self, axis: Optional[int] = None) -> U32Array:
        """Index of the first minimum (NaN counts as minimum); see :meth:`argmax`."""
        return self._argreduce(1, axis)

    def shuffle(self, rng) -> "Array":
        """Rows (leading axis) in a random order drawn from ``rng`` (a
        ``vulkpy.random.Xoshiro128pp``): ``self.gather(rng.permutation(len), axis=0)``.  The
        reference lists shuffle as missing (README.md:77; example/02-nn.py:82 shuffles on the host)."""
        return self.gather(rng.permutation(int(self.shape[0])), axis=0)

    # -- broadcast / gather (reference: vkarray.py:1434-1521) ----------------------------------------
    def broadcast_to(self, shape: Iterable[int]) -> "Array":
        """Materialise the array broadcast to ``shape``; ``ValueError`` if not broadcastable."""
        shape = tuple(int(s) for s in np.asarray(shape, dtype=int).reshape(-1))
        if tuple(np.broadcast_shapes(self.shape, shape)) != shape:
            raise ValueError(f"Cannot broadcast to {shape}")
        ret = self._new(shape)
        shA = np.ones(len(shape), dtype=np.uint32)
        if len(self.shape):
            shA[len(shape) - len(self.shape):] = self.shape
        shB = np.asarray(shape, dtype=np.uint32)
        p = BroadcastParams(self.buffer.size(), ret.buffer.size(), len(shape))
        return self._run(ret, "broadcast", [self, ret, shA, shB], p, [self])

    def gather(self, indices: U32Array, axis: Optional[int] = None) -> "Array":
        """``a.flat[indices]`` or, with ``axis``, ``take(a, indices, axis)`` with the index
        dimensions leading: result shape ``indices.shape + prev + post``."""
        size = indices.buffer.size()
        if axis is None:
            ret = self._new(indices.shape)
            ret.job = self._gpu._submit("gather", 64, 1, 1, [self, indices, ret], DataShape(size, 1, 1),
                                        VectorParams(size))
        else:
            (a,) = self._norm_axis(axis)
            prev_shape, post_shape = tuple(self.shape[:a]), tuple(self.shape[a + 1:])
            ret = self._new(tuple(indices.shape) + prev_shape + post_shape)
            prev = _prod(prev_shape)
            post = _prod(post_shape)
            ret.job = self._gpu._submit("gather_axis", 1, 64, 1, [self, indices, ret],
                                        DataShape(prev, post, size),
                                        AxisGatherParams(prev, post, int(self.shape[a]), size))
        ret._keep = [self, indices]
        return ret


def _install_unary(name: str):
    def method(self, inplace: bool = False) -> Array:
        return self._unary(name, inplace)
    method.__name__ = name
    method.__qualname__ = f"Array.{name}"
    method.__doc__ = (f"Element-wise ``{name}`` (reference shader {name}.comp / i{name}.comp). "
                      "With ``inplace=True`` the array is overwritten and returned.")
    setattr(Array, name, method)


for _n in _UNARY:
    _install_unary(_n)
del _n


def zeros(gpu: GPU, shape: Iterable[int]) -> Array:
    """Zero-initialised array (reference: vkarray.py:1524-1542); the fill runs on the device."""
    z = Array(gpu, shape=shape)
    z[:] = 0.0
    return z
