"""
ctypes binding of ``libvulkpy_b200.so`` (C ABI declared in ``include/vulkpy_b200.h``).

This module plays the role of the reference's pybind11 extension ``vulkpy._vkarray``
(reference: vulkpy/_vkarray.cc:756-898): it exports ``createGPU``, the ``*Params``
parameter blocks, ``DataShape``, ``Job``, ``Buffer``/``Shape`` and ``Xoshiro128pp`` with the
same names and argument order, implemented over CUDA instead of Vulkan.

There is no CPU or PyTorch fallback: if the shared library or a CUDA device is missing the
import (or ``createGPU``) fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvulkpy_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python vulkpy_b200/build.py` "
        "(nvcc, sm_100a). vulkpy_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp = C.c_void_p
_u32 = C.c_uint32
_u64 = C.c_uint64
_sz = C.c_size_t

# name -> (restype, argtypes); every symbol of include/vulkpy_b200.h
PROTOTYPES = {
    "vkp_abi_version": (C.c_int, []),
    "vkp_last_error": (C.c_char_p, []),
    "vkp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "vkp_ctx_create": (C.c_int, [C.c_int, C.c_float, C.POINTER(_vp)]),
    "vkp_ctx_destroy": (C.c_int, [_vp]),
    "vkp_ctx_sync": (C.c_int, [_vp]),
    "vkp_ctx_device": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "vkp_ctx_sm_count": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "vkp_ctx_set_debug_sync": (C.c_int, [_vp, C.c_int]),
    "vkp_ctx_launch_count": (C.c_int, [_vp, C.POINTER(_u64)]),
    "vkp_ctx_mem_info": (C.c_int, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "vkp_ctx_trim": (C.c_int, [_vp]),
    "vkp_alloc": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "vkp_free": (C.c_int, [_vp, _vp]),
    "vkp_alloc_for_upload": (C.c_int, [_vp, _sz, C.POINTER(_vp)]),
    "vkp_upload_async": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "vkp_download_async": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "vkp_host_view": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "vkp_argreduce": (C.c_int, [_vp, C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "vkp_argsort_u32": (C.c_int, [_vp, _vp, _vp, C.c_uint32, C.POINTER(_vp)]),
    "vkp_upload": (C.c_int, [_vp, _vp, _vp, _sz]),
    "vkp_download": (C.c_int, [_vp, _vp, _vp, _sz]),
    "vkp_host_acquire": (C.c_int, [_vp, _vp, _sz, C.c_int]),
    "vkp_host_alloc": (C.c_int, [_sz, C.POINTER(_vp)]),
    "vkp_host_free": (C.c_int, [_vp]),
    "vkp_op_id": (C.c_int, [C.c_char_p]),
    "vkp_op_name": (C.c_char_p, [C.c_int]),
    "vkp_op_count": (C.c_int, []),
    "vkp_submit": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.c_int, _vp, _sz, C.POINTER(_vp)]),
    "vkp_fill_u32": (C.c_int, [_vp, _vp, _sz, _u32, C.POINTER(_vp)]),
    "vkp_ew_chain": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), _vp, _sz, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                               C.POINTER(C.c_float), C.POINTER(_vp)]),
    "vkp_gemm": (C.c_int, [_vp, C.c_int, C.c_int, _u32, _u32, _u32, _vp, _vp, _vp, _vp, C.c_int,
                           C.POINTER(_vp)]),
    "vkp_gemm_fused": (C.c_int, [_vp, C.c_int, C.c_int, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, C.c_int,
                                 C.POINTER(_vp)]),
    "vkp_nn_adam": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _sz] + [C.c_float] * 8 + [C.POINTER(_vp)]),
    "vkp_nn_activation_backward": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "vkp_nn_softmax_forward": (C.c_int, [_vp, _vp, _vp, _u32, _u32, C.POINTER(_vp)]),
    "vkp_nn_softmax_ce_train": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _u32, _u32, C.c_float, C.c_int, C.POINTER(_vp)]),
    "vkp_job_wait": (C.c_int, [_vp, _u64]),
    "vkp_job_done": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "vkp_job_release": (C.c_int, [_vp]),
    "vkp_rng_create": (C.c_int, [_vp, _u32, _u64, C.c_int, C.POINTER(_vp)]),
    "vkp_rng_destroy": (C.c_int, [_vp]),
    "vkp_rng_uint32": (C.c_int, [_vp, _vp, _u32, C.POINTER(_vp)]),
    "vkp_rng_float": (C.c_int, [_vp, _vp, _u32, C.POINTER(_vp)]),
    "vkp_rng_normal": (C.c_int, [_vp, _vp, _u32, C.c_float, C.c_float, C.POINTER(_vp)]),
    "vkp_rng_state": (C.c_int, [_vp, _vp]),
    "vkp_rng_advance": (C.c_int, [_vp, _u64]),
    "vkp_timer_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "vkp_timer_record": (C.c_int, [_vp]),
    "vkp_timer_elapsed_ms": (C.c_int, [_vp, _vp, C.POINTER(C.c_float)]),
    "vkp_timer_destroy": (C.c_int, [_vp]),
    "vkp_nn_adam_apply_many": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp),
                                         C.POINTER(_sz), C.POINTER(C.c_float), C.POINTER(_vp)]),
    "vkp_fill_many_u32": (C.c_int, [_vp, C.c_int, C.POINTER(_vp), C.POINTER(_sz), C.c_uint32, C.POINTER(_vp)]),
    "vkp_comm_unique_id": (C.c_int, [_vp]),
    "vkp_comm_init": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "vkp_comm_destroy": (C.c_int, [_vp]),
    "vkp_comm_allreduce": (C.c_int, [_vp, _vp, _vp, _sz, C.c_int, C.POINTER(_vp)]),
    "vkp_comm_allgather": (C.c_int, [_vp, _vp, _vp, _sz, C.POINTER(_vp)]),
    "vkp_comm_allreduce_multi": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_sz), C.c_int, C.c_int, C.c_float, C.POINTER(_vp)]),
    "vkp_comm_reduce_allreduce": (C.c_int, [_vp, C.c_int, _vp, _vp, _u32, _u32, _u32, C.POINTER(_vp)]),
    "vkp_comm_barrier": (C.c_int, [_vp, C.POINTER(_vp)]),
    "vkp_comm_peer_mode": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "vkp_comm_matmul_allgather": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp, C.POINTER(_vp)]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _f = getattr(lib, _name)  # AttributeError here = header and library out of sync
    _f.restype = _res
    _f.argtypes = _args

ABI_VERSION = 1
if lib.vkp_abi_version() != ABI_VERSION:
    raise ImportError("libvulkpy_b200.so ABI version mismatch; rebuild with `python vulkpy_b200/build.py`")

UINT64_MAX = (1 << 64) - 1
COMM_ID_BYTES = 128


def _check(rc: int):
    if rc != 0:
        msg = lib.vkp_last_error()
        raise RuntimeError(msg.decode("utf-8", "replace") if msg else f"vulkpy_b200 error {rc}")


# ---------------------------------------------------------------------------------------
# parameter blocks -- same names / positional constructors as _vkarray.cc:835-874
# ---------------------------------------------------------------------------------------
def _params(name: str, fields: Sequence[tuple]):
    return type(name, (C.Structure,), {"_fields_": list(fields), "__slots__": ()})


VectorParams = _params("VectorParams", [("size", _u32)])
MultiVector2Params = _params("MultiVector2Params", [("size0", _u32), ("size1", _u32)])
VectorScalarParams = _params("VectorScalarParams", [("size", _u32), ("scalar", C.c_float)])
VectorScalar2Params = _params("VectorScalar2Params",
                              [("size", _u32), ("scalar0", C.c_float), ("scalar1", C.c_float)])
MatMulParams = _params("MatMulParams", [("rowA", _u32), ("contractSize", _u32), ("columnB", _u32)])
AxisReductionParams = _params("AxisReductionParams",
                              [("prev_prod", _u32), ("axis_size", _u32), ("post_prod", _u32)])
BroadcastParams = _params("BroadcastParams", [("size0", _u32), ("size1", _u32), ("ndim", _u32)])
Multi3BroadcastParams = _params("Multi3BroadcastParams",
                                [("size0", _u32), ("size1", _u32), ("size2", _u32), ("ndim", _u32)])
BatchAffineParams = _params("BatchAffineParams",
                            [("batch_size", _u32), ("input_size", _u32), ("output_size", _u32)])
VectorRangeParams = _params("VectorRangeParams", [("size", _u32), ("low", _u32), ("high", _u32)])
AxisGatherParams = _params("AxisGatherParams", [("prev_prod", _u32), ("post_prod", _u32),
                                                ("axis_size", _u32), ("index_size", _u32)])
ShiftVectorParams = _params("ShiftVectorParams", [("shift", _u32), ("size", _u32)])
DataShape = _params("DataShape", [("x", _u32), ("y", _u32), ("z", _u32)])


# ---------------------------------------------------------------------------------------
# op ids: the reference identifies a kernel by the path of its .spv (util.py:58-72)
# ---------------------------------------------------------------------------------------
OPS = {lib.vkp_op_name(i).decode(): i for i in range(lib.vkp_op_count())}


def op_id(spv: str) -> int:
    """Map a shader name (``"add"``, ``"add.spv"`` or a path ending in it) to its op id."""
    i = OPS.get(spv)
    if i is None:
        i = lib.vkp_op_id(spv.encode())
        if i < 0:
            raise RuntimeError("Unknown Operation")  # _vkarray.cc:752
        OPS[spv] = i
    return i


class Job:
    """Handle of one submitted operation (reference ``Job``: _vkarray.cc:392-457, :876-879)."""
    __slots__ = ("_h",)

    def __init__(self, handle: int):
        self._h = handle

    def wait(self, timeout_ns: Optional[int] = None):
        """Block until the operation finished; ``RuntimeError`` on timeout or device error."""
        if self._h:
            _check(lib.vkp_job_wait(self._h, UINT64_MAX if timeout_ns is None else int(timeout_ns)))

    def done(self) -> bool:
        if not self._h:
            return True
        d = C.c_int(0)
        _check(lib.vkp_job_done(self._h, C.byref(d)))
        return bool(d.value)

    def __del__(self):
        h, self._h = self._h, None
        if h:
            try:
                lib.vkp_job_release(h)
            except Exception:  # interpreter shutdown
                pass


class _BufferBase:
    """Device buffer visible to the host through the same pointer (managed memory).

    Reference: ``Buffer<T>`` (_vkarray.cc:38-130) exported with the buffer protocol
    (:797-833).  ``np.asarray(buffer)`` gives a zero-copy 1-D view that keeps the buffer alive.
    """
    _dtype = np.dtype(np.float32)
    __slots__ = ("_dev", "ptr", "_n", "__weakref__")

    def __init__(self, dev: "Device", n: int, for_upload: bool = False):
        self._dev = dev
        self._n = int(n)
        p = _vp()
        _check((lib.vkp_alloc_for_upload if for_upload else lib.vkp_alloc)(dev._ctx, self._n * 4, C.byref(p)))
        self.ptr = p.value

    def size(self) -> int:
        return self._n

    def info(self):  # BufferInfo of the reference: here the buffer itself
        return self

    def range(self):  # MemoryRange of the reference
        return self

    @property
    def nbytes(self) -> int:
        return self._n * 4

    def host_view(self):
        """Make the buffer host-visible (first call moves it from device memory into a managed
        block; ``ptr`` changes once and then stays put for the life of the buffer)."""
        p = _vp()
        _check(lib.vkp_host_view(self._dev._ctx, self.ptr, C.byref(p)))
        self.ptr = p.value

    @property
    def __array_interface__(self):
        self.host_view()
        return {"shape": (self._n,), "typestr": self._dtype.str, "data": (self.ptr, False), "version": 3}

    def host_acquire(self, prefetch: bool = False, write: bool = False):
        self.host_view()
        _check(lib.vkp_host_acquire(self._dev._ctx, self.ptr, self._n * 4,
                                    (1 if prefetch else 0) | (2 if write else 0)))

    def upload(self, host: np.ndarray):
        """Stream-ordered copy of a contiguous host array into the buffer."""
        assert host.nbytes == self._n * 4 and host.flags.c_contiguous
        _check(lib.vkp_upload(self._dev._ctx, self.ptr, host.ctypes.data, host.nbytes))

    def upload_async(self, pinned: np.ndarray) -> Job:
        """Copy-engine upload from a ``pinned_empty`` array; overlaps compute and downloads."""
        assert pinned.nbytes == self._n * 4 and pinned.flags.c_contiguous
        job = _vp()
        _check(lib.vkp_upload_async(self._dev._ctx, self.ptr, pinned.ctypes.data, pinned.nbytes, C.byref(job)))
        return Job(job.value)

    def download_async(self, pinned: np.ndarray) -> Job:
        """Copy-engine download into a ``pinned_empty`` array; overlaps compute and uploads."""
        assert pinned.nbytes == self._n * 4 and pinned.flags.c_contiguous
        job = _vp()
        _check(lib.vkp_download_async(self._dev._ctx, pinned.ctypes.data, self.ptr, pinned.nbytes, C.byref(job)))
        return Job(job.value)

    def __del__(self):
        p, self.ptr = self.ptr, None
        if p:
            try:
                lib.vkp_free(self._dev._ctx, p)
            except Exception:
                pass


class Buffer(_BufferBase):
    """float32 buffer (reference class ``Buffer``)."""
    __slots__ = ()


class Shape(_BufferBase):
    """uint32 buffer (reference class ``Shape``)."""
    _dtype = np.dtype(np.uint32)
    __slots__ = ()


class Device:
    """One CUDA device + in-order stream (reference class ``GPU``: _vkarray.cc:460-574, :765-795)."""

    def __init__(self, idx: int, priority: float):
        h = _vp()
        _check(lib.vkp_ctx_create(int(idx), float(priority), C.byref(h)))
        self._ctx = h.value
        self.index = int(idx)

    # -- buffers ------------------------------------------------------------------------
    def createBuffer(self, n: int) -> Buffer:
        return Buffer(self, n)

    def createU32Buffer(self, n: int) -> Shape:
        return Shape(self, n)

    def toBuffer(self, data) -> Buffer:
        host = np.ascontiguousarray(data, dtype=np.float32).ravel()
        b = Buffer(self, host.size)
        b.upload(host)
        return b

    def toU32Buffer(self, data) -> Shape:
        host = np.ascontiguousarray(data, dtype=np.uint32).ravel()
        b = Shape(self, host.size)
        b.upload(host)
        return b

    # -- ops ----------------------------------------------------------------------------
    def submit(self, spv, x: int, y: int, z: int, infos: Iterable, shape, params,
               wait: Iterable[Job] = ()) -> Job:
        """``GPU.submit(spv, x, y, z, infos, DataShape, Params, wait)`` (_vkarray.cc:770-791).

        ``infos`` are buffers (or, for the shape bindings of the broadcast family, host
        ``uint32`` arrays).  ``x, y, z``, ``shape`` and ``wait`` are accepted for signature
        compatibility: grids derive from ``params`` and the stream is in-order.
        """
        op = spv if isinstance(spv, int) else op_id(spv)
        ptrs = []
        for b in infos:
            if isinstance(b, _BufferBase):
                ptrs.append(b.ptr)
            elif isinstance(b, np.ndarray):
                ptrs.append(b.ctypes.data)
            else:
                raise TypeError(f"cannot bind {type(b)!r}")
        arr = (_vp * len(ptrs))(*ptrs)
        job = _vp()
        _check(lib.vkp_submit(self._ctx, op, arr, len(ptrs), C.byref(params), C.sizeof(params),
                              C.byref(job)))
        return Job(job.value)

    def fill(self, buf: _BufferBase, bits: int) -> Job:
        job = _vp()
        _check(lib.vkp_fill_u32(self._ctx, buf.ptr, buf.size(), bits & 0xFFFFFFFF, C.byref(job)))
        return Job(job.value)

    def ew_chain(self, inputs, out: Buffer, ops, srcs, scalars) -> Job:
        """One launch for a chain of same-shape element-wise operations (``vkp_ew_chain``)."""
        n, k = len(inputs), len(ops)
        ptrs = (_vp * n)(*[b.ptr for b in inputs])
        job = _vp()
        _check(lib.vkp_ew_chain(self._ctx, n, ptrs, out.ptr, out.size(), k, (C.c_int * k)(*ops), (C.c_int * k)(*srcs),
                                (C.c_float * k)(*scalars), C.byref(job)))
        return Job(job.value)

    def gemm(self, transA: bool, transB: bool, M: int, N: int, K: int, A: Buffer, B: Buffer, Cbuf: Buffer,
             bias: Optional[Buffer] = None, flags: int = 0, relu_mask: Optional[Buffer] = None) -> Job:
        job = _vp()
        _check(lib.vkp_gemm_fused(self._ctx, int(transA), int(transB), M, N, K, A.ptr, B.ptr, Cbuf.ptr,
                                  bias.ptr if bias is not None else None,
                                  relu_mask.ptr if relu_mask is not None else None, flags, C.byref(job)))
        return Job(job.value)

    def argreduce(self, op: int, src: Buffer, dst: "Shape", prev: int, axis: int, post: int) -> Job:
        job = _vp()
        _check(lib.vkp_argreduce(self._ctx, op, src.ptr, dst.ptr, prev, axis, post, C.byref(job)))
        return Job(job.value)

    def argsort_u32(self, keys: "Shape", dst: "Shape") -> Job:
        job = _vp()
        _check(lib.vkp_argsort_u32(self._ctx, keys.ptr, dst.ptr, keys.size(), C.byref(job)))
        return Job(job.value)

    def nn_adam(self, grad: Buffer, m: Buffer, v: Buffer, diff: Buffer, *scalars: float) -> Job:
        job = _vp()
        _check(lib.vkp_nn_adam(self._ctx, grad.ptr, m.ptr, v.ptr, diff.ptr, grad.size(), *scalars, C.byref(job)))
        return Job(job.value)

    def nn_adam_apply_many(self, grads, ms, vs, values, scalars) -> Job:
        """One launch for the Adam step + ``value += diff`` of up to 16 parameters."""
        n = len(grads)
        arr = lambda bufs: (_vp * n)(*[b.ptr for b in bufs])
        counts = (_sz * n)(*[b.size() for b in grads])
        sc = (C.c_float * (8 * n))(*[x for row in scalars for x in row])
        job = _vp()
        _check(lib.vkp_nn_adam_apply_many(self._ctx, n, arr(grads), arr(ms), arr(vs), arr(values), counts, sc, C.byref(job)))
        return Job(job.value)

    def fill_many(self, bufs, bits: int) -> Job:
        n = len(bufs)
        ptrs = (_vp * n)(*[b.ptr for b in bufs])
        counts = (_sz * n)(*[b.size() for b in bufs])
        job = _vp()
        _check(lib.vkp_fill_many_u32(self._ctx, n, ptrs, counts, bits & 0xFFFFFFFF, C.byref(job)))
        return Job(job.value)

    def nn_activation_backward(self, kind: int, y: Buffer, dy: Buffer, dx: Buffer) -> Job:
        job = _vp()
        _check(lib.vkp_nn_activation_backward(self._ctx, kind, y.ptr, dy.ptr, dx.ptr, y.size(), C.byref(job)))
        return Job(job.value)

    def nn_softmax_forward(self, x: Buffer, y: Buffer, rows: int, cols: int) -> Job:
        job = _vp()
        _check(lib.vkp_nn_softmax_forward(self._ctx, x.ptr, y.ptr, rows, cols, C.byref(job)))
        return Job(job.value)

    def nn_softmax_ce_train(self, z: Buffer, t: Buffer, p: Buffer, L: Buffer, dz: Buffer, rows: int, cols: int,
                            scale: Optional[float]) -> Job:
        job = _vp()
        _check(lib.vkp_nn_softmax_ce_train(self._ctx, z.ptr, t.ptr, p.ptr, L.ptr, dz.ptr, rows, cols,
                                           0.0 if scale is None else float(scale), 0 if scale is None else 1, C.byref(job)))
        return Job(job.value)

    def wait(self):
        _check(lib.vkp_ctx_sync(self._ctx))

    def flush(self, ranges: List) -> None:
        """Host writes are coherent (managed memory); nothing to flush (_vkarray.cc:551-558)."""
        return None

    def canSubgroupArithmetic(self) -> bool:
        return True  # warp shuffles: the sum_v1.3 family is always available

    # -- extras -------------------------------------------------------------------------
    def launch_count(self) -> int:
        n = _u64(0)
        _check(lib.vkp_ctx_launch_count(self._ctx, C.byref(n)))
        return n.value

    def sm_count(self) -> int:
        n = C.c_int(0)
        _check(lib.vkp_ctx_sm_count(self._ctx, C.byref(n)))
        return n.value

    def mem_info(self):
        a, b = _sz(0), _sz(0)
        _check(lib.vkp_ctx_mem_info(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def trim(self):
        _check(lib.vkp_ctx_trim(self._ctx))

    def set_debug_sync(self, on: bool):
        _check(lib.vkp_ctx_set_debug_sync(self._ctx, int(bool(on))))


def createGPU(n: int, priority: float) -> Device:
    """``createGPU(n, priority)`` (_vkarray.cc:759-763)."""
    return Device(n, priority)


def device_count() -> int:
    n = C.c_int(0)
    _check(lib.vkp_device_count(C.byref(n)))
    return n.value


class Xoshiro128pp:
    """``_vkarray.Xoshiro128pp(gpu, spv_uint32, spv_float, size[, seed])`` (_vkarray.cc:884-897)."""

    def __init__(self, gpu: Device, spv_uint32: str = "", spv_float: str = "", size: int = 64,
                 seed: Optional[int] = None):
        self._dev = gpu
        self.size = int(size)
        h = _vp()
        _check(lib.vkp_rng_create(gpu._ctx, self.size, 0 if seed is None else int(seed) & UINT64_MAX,
                                  0 if seed is None else 1, C.byref(h)))
        self._h = h.value

    def random_uint32(self, n: int, info: _BufferBase) -> Job:
        job = _vp()
        _check(lib.vkp_rng_uint32(self._h, info.ptr, int(n), C.byref(job)))
        return Job(job.value)

    def random_float(self, n: int, info: _BufferBase) -> Job:
        job = _vp()
        _check(lib.vkp_rng_float(self._h, info.ptr, int(n), C.byref(job)))
        return Job(job.value)

    def normal(self, n: int, info: _BufferBase, mean: float, stddev: float) -> Job:
        job = _vp()
        _check(lib.vkp_rng_normal(self._h, info.ptr, int(n), float(mean), float(stddev), C.byref(job)))
        return Job(job.value)

    def advance(self, n: int):
        """Discard ``n`` draws exactly as ``random_uint32(n)`` would consume them."""
        _check(lib.vkp_rng_advance(self._h, int(n)))

    def state(self) -> np.ndarray:
        out = np.empty((self.size, 4), dtype=np.uint32)
        _check(lib.vkp_rng_state(self._h, out.ctypes.data))
        return out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib.vkp_rng_destroy(h)
            except Exception:
                pass


class Timer:
    """CUDA event on the context stream (measurement only)."""

    def __init__(self, dev: Device):
        h = _vp()
        _check(lib.vkp_timer_create(dev._ctx, C.byref(h)))
        self._h = h.value

    def record(self):
        _check(lib.vkp_timer_record(self._h))

    def elapsed_ms(self, stop: "Timer") -> float:
        ms = C.c_float(0)
        _check(lib.vkp_timer_elapsed_ms(self._h, stop._h, C.byref(ms)))
        return ms.value

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib.vkp_timer_destroy(h)
            except Exception:
                pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """Page-locked host array (full-rate PCIe staging for ``Array(data=...)``)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    p = _vp()
    _check(lib.vkp_host_alloc(n * dtype.itemsize, C.byref(p)))

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr
            self.__array_interface__ = {"shape": (n,), "typestr": dtype.str, "data": (ptr, False), "version": 3}

        def __del__(self):
            try:
                lib.vkp_host_free(self.ptr)
            except Exception:
                pass

    return np.asarray(_Owner(p.value)).reshape(shape)
