"""
Drop-in for the reference's pybind11 extension ``vulkpy._vkarray`` (_vkarray.cc:756-898), written
against nothing but the C ABI of ``include/vulkpy_b200.h`` (ctypes; no import from the rest of this
package, so the file can be copied into the reference tree as ``vulkpy/_vkarray.py``).

Same names, argument order and error behaviour as the extension:

=====================================================  ==========================================
``_vkarray`` symbol (_vkarray.cc)                       C ABI
=====================================================  ==========================================
``createGPU(n, priority)`` :759-763                     ``vkp_ctx_create``
``GPU.createBuffer / createU32Buffer`` :767-769         ``vkp_alloc``
``GPU.toBuffer / toU32Buffer`` :766-768                 ``vkp_alloc`` + ``vkp_upload``
``Buffer`` / ``Shape`` buffer protocol :797-833         ``vkp_host_view`` + ``vkp_host_acquire``
``GPU.submit(spv, x, y, z, infos, shape, params, wait)``  ``vkp_op_id`` + ``vkp_submit``
``*Params``, ``DataShape`` :835-874                     ``vkp_*_params`` (same field order)
``Job.wait([timeout_ns])`` :876-879                     ``vkp_job_wait`` / ``vkp_job_release``
``GPU.wait / flush / canSubgroupArithmetic`` :792-795   ``vkp_ctx_sync`` / no-op / ``True``
``Xoshiro128pp(gpu, spv_u32, spv_f32, size[, seed])``   ``vkp_rng_create``
``.random_uint32(n, info) / .random_float(n, info)``    ``vkp_rng_uint32`` / ``vkp_rng_float``
=====================================================  ==========================================

Errors surface as ``RuntimeError`` with the library's message, as the extension's C++ exceptions do
(_vkarray.cc:32,286,446-456,508,752).  ``wait`` lists are accepted and ignored: the context stream is
in order.  Shape bindings of the broadcast family (add_broadcast.comp binding 3, iadd_broadcast.comp
binding 2, broadcast.comp bindings 2-3) are consumed on the host at submit time, so the shim hands the
library the host view of those ``Shape`` buffers.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFAULT = os.path.join(os.path.dirname(_HERE), "libvulkpy_b200.so")
_lib = C.CDLL(os.environ.get("VULKPY_B200_LIB", _DEFAULT if os.path.exists(_DEFAULT) else "libvulkpy_b200.so"))
_lib.vkp_last_error.restype = C.c_char_p
_lib.vkp_op_id.argtypes = [C.c_char_p]
_vp, _u32, _u64, _sz, _f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t, C.c_float
_lib.vkp_ctx_create.argtypes = [C.c_int, _f32, C.POINTER(_vp)]
_lib.vkp_ctx_sync.argtypes = [_vp]
_lib.vkp_alloc.argtypes = [_vp, _sz, C.POINTER(_vp)]
_lib.vkp_free.argtypes = [_vp, _vp]
_lib.vkp_upload.argtypes = [_vp, _vp, _vp, _sz]
_lib.vkp_host_view.argtypes = [_vp, _vp, C.POINTER(_vp)]
_lib.vkp_host_acquire.argtypes = [_vp, _vp, _sz, C.c_int]
_lib.vkp_submit.argtypes = [_vp, C.c_int, C.POINTER(_vp), C.c_int, _vp, _sz, C.POINTER(_vp)]
_lib.vkp_job_wait.argtypes = [_vp, _u64]
_lib.vkp_job_release.argtypes = [_vp]
_lib.vkp_rng_create.argtypes = [_vp, _u32, _u64, C.c_int, C.POINTER(_vp)]
_lib.vkp_rng_destroy.argtypes = [_vp]
_lib.vkp_rng_uint32.argtypes = [_vp, _vp, _u32, C.POINTER(_vp)]
_lib.vkp_rng_float.argtypes = [_vp, _vp, _u32, C.POINTER(_vp)]


def _ck(rc):
    if rc:
        raise RuntimeError((_lib.vkp_last_error() or b"vulkpy_b200 error").decode("utf-8", "replace"))


def _params(name, fields):
    return type(name, (C.Structure,), {"_fields_": fields})


VectorParams = _params("VectorParams", [("size", _u32)])
MultiVector2Params = _params("MultiVector2Params", [("size0", _u32), ("size1", _u32)])
VectorScalarParams = _params("VectorScalarParams", [("size", _u32), ("scalar", _f32)])
VectorScalar2Params = _params("VectorScalar2Params", [("size", _u32), ("scalar0", _f32), ("scalar1", _f32)])
MatMulParams = _params("MatMulParams", [("rowA", _u32), ("contractSize", _u32), ("columnB", _u32)])
AxisReductionParams = _params("AxisReductionParams", [("prev_prod", _u32), ("axis_size", _u32), ("post_prod", _u32)])
BroadcastParams = _params("BroadcastParams", [("size0", _u32), ("size1", _u32), ("ndim", _u32)])
Multi3BroadcastParams = _params("Multi3BroadcastParams", [("size0", _u32), ("size1", _u32), ("size2", _u32), ("ndim", _u32)])
BatchAffineParams = _params("BatchAffineParams", [("batch_size", _u32), ("input_size", _u32), ("output_size", _u32)])
VectorRangeParams = _params("VectorRangeParams", [("size", _u32), ("low", _u32), ("high", _u32)])
AxisGatherParams = _params("AxisGatherParams", [("prev_prod", _u32), ("post_prod", _u32), ("axis_size", _u32), ("index_size", _u32)])
ShiftVectorParams = _params("ShiftVectorParams", [("shift", _u32), ("size", _u32)])
DataShape = _params("DataShape", [("x", _u32), ("y", _u32), ("z", _u32)])

_UINT64_MAX = (1 << 64) - 1


class Job:
    """_vkarray.cc:392-457, :876-879"""

    def __init__(self, handle):
        self._h = handle

    def wait(self, timeout_ns=_UINT64_MAX):
        if self._h:
            _ck(_lib.vkp_job_wait(self._h, _u64(int(timeout_ns))))

    def __del__(self):
        h, self._h = self._h, None
        if h:
            try:
                _lib.vkp_job_release(h)
            except Exception:
                pass


class _Buf:
    """Buffer<T> with the buffer protocol (_vkarray.cc:38-130, :797-833)."""
    _typestr = "<f4"

    def __init__(self, gpu, n):
        self._gpu, self._n = gpu, int(n)
        p = _vp()
        _ck(_lib.vkp_alloc(gpu._ctx, 4 * self._n, C.byref(p)))
        self.ptr = p.value

    def _host_ptr(self):
        p = _vp()      # first view: the buffer moves into host-visible memory (the pointer changes once)
        _ck(_lib.vkp_host_view(self._gpu._ctx, self.ptr, C.byref(p)))
        self.ptr = p.value
        _ck(_lib.vkp_host_acquire(self._gpu._ctx, self.ptr, 4 * self._n, 1))
        return self.ptr

    @property
    def __array_interface__(self):
        return {"shape": (self._n,), "typestr": self._typestr, "data": (self._host_ptr(), False), "version": 3}

    def info(self):
        return self

    def range(self):
        return self

    def size(self):
        return self._n

    def __del__(self):
        p, self.ptr = self.ptr, None
        if p:
            try:
                _lib.vkp_free(self._gpu._ctx, p)
            except Exception:
                pass


class Buffer(_Buf):
    pass


class Shape(_Buf):
    _typestr = "<u4"


def _first_shape_binding(name):
    """Index of the first binding that carries a shape (host-consumed), or None."""
    base = os.path.basename(name)
    if base.endswith(".spv"):
        base = base[:-4]
    if base == "broadcast":
        return 2
    if base.endswith("_broadcast"):
        return 2 if base.startswith("i") else 3
    return None


class GPU:
    """_vkarray.cc:460-574, :765-795"""

    def __init__(self, n, priority):
        h = _vp()
        _ck(_lib.vkp_ctx_create(int(n), _f32(priority), C.byref(h)))
        self._ctx = h.value

    def createBuffer(self, n):
        return Buffer(self, n)

    def createU32Buffer(self, n):
        return Shape(self, n)

    def _to(self, cls, data, dtype):
        host = np.ascontiguousarray(data, dtype=dtype).ravel()
        b = cls(self, host.size)
        if host.size:
            _ck(_lib.vkp_upload(self._ctx, b.ptr, host.ctypes.data, host.nbytes))
        return b

    def toBuffer(self, data):
        return self._to(Buffer, data, np.float32)

    def toU32Buffer(self, data):
        return self._to(Shape, data, np.uint32)

    def submit(self, spv, x, y, z, infos, shape, params, wait=()):
        op = _lib.vkp_op_id(str(spv).encode())
        if op < 0:
            raise RuntimeError("Unknown Operation")                 # _vkarray.cc:752
        infos = list(infos)
        first = _first_shape_binding(str(spv))
        ptrs = []
        for i, b in enumerate(infos):
            if first is not None and i >= first:
                _ck(_lib.vkp_ctx_sync(self._ctx))                   # the shape buffer may still be uploading
                ptrs.append(b._host_ptr())
            else:
                ptrs.append(b.ptr)
        arr = (_vp * len(ptrs))(*ptrs)
        job = _vp()
        _ck(_lib.vkp_submit(self._ctx, op, arr, len(ptrs), C.byref(params), C.sizeof(params), C.byref(job)))
        return Job(job.value)

    def wait(self):
        _ck(_lib.vkp_ctx_sync(self._ctx))

    def flush(self, ranges):
        return None

    def canSubgroupArithmetic(self):
        return True


def createGPU(n, priority):
    return GPU(n, priority)


class Xoshiro128pp:
    """_vkarray.cc:577-719, :884-897"""

    def __init__(self, gpu, spv_uint32, spv_float, size, seed=None):
        h = _vp()
        _ck(_lib.vkp_rng_create(gpu._ctx, int(size), (0 if seed is None else int(seed)) & _UINT64_MAX,
                                0 if seed is None else 1, C.byref(h)))
        self._h = h.value

    def random_uint32(self, n, info):
        j = _vp()
        _ck(_lib.vkp_rng_uint32(self._h, info.ptr, int(n), C.byref(j)))
        return Job(j.value)

    def random_float(self, n, info):
        j = _vp()
        _ck(_lib.vkp_rng_float(self._h, info.ptr, int(n), C.byref(j)))
        return Job(j.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.vkp_rng_destroy(h)
            except Exception:
                pass
