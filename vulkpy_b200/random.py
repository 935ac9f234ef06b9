"""
Random Module (:mod:`vulkpy_b200.random`; reference: vulkpy/random.py)

``Xoshiro128pp(gpu, size=64, *, seed=None)`` produces exactly the reference's xoshiro128++
stream: same seeding, same lane layout ``out[c*size + lane]``, same [0, 1) mapping, and the
state persists across calls.  With ``seed=0`` (reference docstring, random.py:12-24):

>>> r = vk.random.Xoshiro128pp(gpu, seed=0)
>>> print(r.random(shape=(3,)))
[0.42977667 0.8235899  0.90622926]
>>> print(r.normal(shape=(3,)))
[-2.3403292  0.7247794  0.7118352]

``normal`` is fused (uniforms never touch memory) when ``size`` is even; it consumes ``n``
uniforms for even ``n`` and ``n + 1`` for odd ``n`` like the reference (random.py:105-124).
"""
from __future__ import annotations

from typing import Iterable, Optional

import math

import numpy as np

from . import _backend as _b
from . import vkarray as vk
from .vktyping import Resource

__all__ = ["Xoshiro128pp"]


def _target(cls, gpu, shape, buffer):
    """Output array of a generator call: a fresh one of ``shape`` or the caller's ``buffer``."""
    if buffer is None:
        if shape is None:
            raise ValueError("One of `shape` and `buffer` must be specified.")
        return cls(gpu, shape=shape)
    buffer.wait()  # reference waits on the output's pending job (random.py:98-100)
    return buffer


class PRNG(Resource):
    """Distribution transforms shared by generators (reference: random.py:40-189)."""
    _2p32 = 1 << 32

    def __init__(self, gpu: vk.GPU):
        self._gpu = gpu

    def random(self, *, shape=None, buffer=None) -> vk.Array:
        raise NotImplementedError

    def randint(self, *, shape=None, buffer=None) -> vk.U32Array:
        raise NotImplementedError

    def normal(self, *, shape: Optional[Iterable[int]] = None, buffer: Optional[vk.Array] = None,
               mean: float = 0.0, stddev: float = 1.0) -> vk.Array:
        """Box-Muller over ``random()`` output: in place for even n, via an (n+1)-element
        temporary for odd n (reference: random.py:60-124)."""
        out = _target(vk.Array, self._gpu, shape, buffer)
        n = math.prod(out.shape)
        p = _b.VectorScalar2Params(n, float(mean), float(stddev))
        d = _b.DataShape(n // 2, 1, 1)
        if n % 2 == 0:
            out = self.random(buffer=out)
            out.job = self._gpu._submit("prng_ibox_muller", 64, 1, 1, [out], d, p)
            out._keep = []
        else:
            u = self.random(shape=(n + 1,))
            out.job = self._gpu._submit("prng_box_muller", 64, 1, 1, [u, out], d, p)
            out._keep = [u]
        return out

    def randrange(self, *, shape: Optional[Iterable[int]] = None, buffer: Optional[vk.U32Array] = None,
                  low: int = 0, high: int = 1 << 32) -> vk.U32Array:
        """Integers in ``[low, high)`` as ``low + uint(float(high-low) * u)`` with ``u`` from
        ``random()`` (reference: random.py:126-186, prng_randrange.comp:20-27)."""
        if low < 0:
            raise ValueError(f"`low` must be non negative integer, but {low}")
        if high > self._2p32:
            raise ValueError(f"`high` must not be greater than 2^32, but {high}")
        if low >= high:
            raise ValueError(f"`low` must be smaller than `high`, but {low}, {high}")
        if low == 0 and high == self._2p32:
            return self.randint(shape=shape, buffer=buffer)
        out = _target(vk.U32Array, self._gpu, shape, buffer)
        size = out.buffer.size()
        u = self.random(shape=out.shape)
        out.job = self._gpu._submit("prng_randrange", 64, 1, 1, [u, out], _b.DataShape(size, 1, 1),
                                    _b.VectorRangeParams(size, low, high - 1))
        out._keep = [u]
        return out

    def permutation(self, n: int) -> vk.U32Array:
        """Random permutation of ``0 .. n-1``: the indices stably sorted by ``n`` keys from
        ``randint`` (``np.argsort(keys, kind="stable")``).  Additive: the reference lists shuffle
        as missing (README.md:77) and example/02-nn.py:82 shuffles an index array on the host."""
        n = int(n)
        if n < 0:
            raise ValueError(f"`n` must be non negative, but {n}")
        keys = self.randint(shape=(n,))
        out = vk.U32Array(self._gpu, shape=(n,))
        out.job = self._gpu.gpu.argsort_u32(keys.buffer, out.buffer)
        out._keep = [keys]
        return out

    def wait(self):
        pass


class Xoshiro128pp(PRNG):
    """xoshiro128++ with ``size`` parallel lanes spaced by the reference's jump
    (reference: random.py:192-312, _vkarray.cc:577-719)."""

    def __init__(self, gpu: vk.GPU, size: int = 64, *, seed: Optional[int] = None):
        super().__init__(gpu)
        self.rng = _b.Xoshiro128pp(gpu.gpu, "prng_xoshiro128pp_uint32", "prng_xoshiro128pp_float",
                                   size, seed)

    def random(self, *, shape: Optional[Iterable[int]] = None,
               buffer: Optional[vk.Array] = None) -> vk.Array:
        """Uniform float32 in [0, 1)."""
        out = _target(vk.Array, self._gpu, shape, buffer)
        out.job = self.rng.random_float(math.prod(out.shape), out.buffer.info())
        out._keep = [self]
        return out

    def randint(self, *, shape: Optional[Iterable[int]] = None,
                buffer: Optional[vk.U32Array] = None) -> vk.U32Array:
        """Uniform uint32 in [0, 2^32)."""
        out = _target(vk.U32Array, self._gpu, shape, buffer)
        out.job = self.rng.random_uint32(math.prod(out.shape), out.buffer.info())
        out._keep = [self]
        return out

    def normal(self, *, shape: Optional[Iterable[int]] = None, buffer: Optional[vk.Array] = None,
               mean: float = 0.0, stddev: float = 1.0) -> vk.Array:
        """Gaussian numbers; one fused kernel when the lane count is even."""
        if self.rng.size % 2:
            return super().normal(shape=shape, buffer=buffer, mean=mean, stddev=stddev)
        out = _target(vk.Array, self._gpu, shape, buffer)
        n = math.prod(out.shape)
        out.job = self.rng.normal(n, out.buffer.info(), mean, stddev)
        out._keep = [self]
        return out
