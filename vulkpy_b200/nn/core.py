"""
Abstract bases of :mod:`vulkpy_b200.nn` (reference: vulkpy/nn/core.py).
"""
from __future__ import annotations

from typing import Iterable

from ..vkarray import GPU, Array

__all__ = ["OptimizerState", "Optimizer", "Regularizer", "Loss", "Module"]


class OptimizerState:
    """Per-parameter mutable optimizer state; ``grad2diff`` maps the accumulated gradient to
    the update that ``Parameter.update`` adds to the value (reference: core.py:24-62)."""

    def grad2diff(self, grad: Array) -> Array:
        raise NotImplementedError


class Optimizer:
    """Holds hyper-parameters and creates one ``OptimizerState`` per parameter
    (reference: core.py:65-121)."""

    def init_state(self, shape: Iterable[int]) -> OptimizerState:
        raise NotImplementedError


class Loss:
    """``loss(x, y)`` computes the loss and remembers its inputs, ``grad()`` returns dL/dx
    (reference: core.py:124-178)."""

    def __call__(self, x: Array, y: Array) -> Array:
        raise NotImplementedError

    def grad(self) -> Array:
        raise NotImplementedError


class Regularizer:
    """Penalty on a parameter array (reference: core.py:181-231)."""

    def loss(self, param: Array) -> Array:
        raise NotImplementedError

    def grad(self, param: Array) -> Array:
        raise NotImplementedError


class Module:
    """Layer base class.  Calling a module runs ``forward`` and keeps the input (``_x``) and the
    output (``_y``) for ``backward`` (reference: core.py:234-337)."""

    def __call__(self, x: Array) -> Array:
        if len(x.shape) < 2:
            raise ValueError("Input must have at least 2-dimensions.")
        self._x = x
        self._y = self.forward(x)
        return self._y

    def forward(self, x: Array) -> Array:
        raise NotImplementedError

    def backward(self, dy: Array) -> Array:
        raise NotImplementedError

    def zero_grad(self):
        pass

    def update(self):
        pass
