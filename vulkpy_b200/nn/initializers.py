"""
Parameter initializers (reference: vulkpy/nn/initializers.py).
"""
from __future__ import annotations

from typing import Iterable, Optional

import numpy as np

from ..vkarray import GPU, Array
from ..random import Xoshiro128pp

__all__ = ["Constant", "HeNormal"]


class Initializer:
    def __call__(self, gpu: GPU, shape: Iterable[int]) -> Array:
        raise NotImplementedError


class Constant(Initializer):
    """Every element equals ``value`` (device fill; reference: initializers.py:21-49)."""

    def __init__(self, value: float):
        self.value = value

    def __call__(self, gpu: GPU, shape: Iterable[int]) -> Array:
        p = Array(gpu, shape=shape)
        p[:] = self.value
        return p


class HeNormal(Initializer):
    """N(0, 2/input_dim) from the initializer's own ``Xoshiro128pp(gpu, seed=seed)``
    (reference: initializers.py:52-89)."""

    def __init__(self, gpu: GPU, input_dim: int, *, seed: Optional[int] = None):
        self.rng = Xoshiro128pp(gpu, seed=seed)
        self.stddev = np.sqrt(2 / input_dim)

    def __call__(self, gpu: GPU, shape: Iterable[int]) -> Array:
        return self.rng.normal(shape=shape, stddev=self.stddev)
