"""
Trainable parameter (reference: vulkpy/nn/parameters.py).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

from ..vkarray import GPU, Array, zeros
from .core import Optimizer, OptimizerState, Regularizer
from .optimizers import Adam

__all__ = ["Parameter"]


class Parameter:
    """Value + accumulated gradient + optimizer state (+ optional regularizer).

    ``initializer`` defaults to zeros and ``opt`` to ``Adam(gpu)`` for trainable parameters
    (reference: parameters.py:14-67)."""

    def __init__(self, gpu: GPU, shape: Iterable[int], trainable: bool = True,
                 opt: Optional[Optimizer] = None,
                 initializer: Optional[Callable[[GPU, Iterable[int]], Array]] = None,
                 regularizer: Optional[Regularizer] = None):
        self.value: Array = (initializer or zeros)(gpu, shape)
        self.grad: Optional[Array] = None
        self.opt_state: Optional[OptimizerState] = None
        if trainable:
            self.grad = zeros(gpu, shape=shape)
            self.opt_state = (opt or Adam(gpu)).init_state(shape)
        self.R: Optional[Regularizer] = regularizer
        # True: `grad` stands for zeros that were never written (Sequence's launch-free zero_grad); the next
        # contribution overwrites.  Every public path below resolves the mark first.
        self._fresh = False

    def is_trainable(self) -> bool:
        return self.grad is not None

    def _materialize_zero(self):
        if self._fresh:
            self._fresh = False
            if self.grad is not None:
                self.grad[:] = 0.0

    def _take_grad(self, grad: Array):
        """``add_grad`` of a temporary nobody else holds: on a fresh gradient the temporary BECOMES the
        gradient (0 + x = x), otherwise ``grad += temporary``."""
        if self.grad is None:
            return
        if self._fresh and tuple(grad.shape) == tuple(self.grad.shape):
            self._fresh = False
            self.grad = grad
            return
        self.add_grad(grad)

    def add_grad(self, grad: Array):
        if self.grad is not None:
            self._materialize_zero()
            self.grad += grad

    def zero_grad(self):
        """Device fill (the reference zeroes through the host view: parameters.py:81-86)."""
        self._fresh = False
        if self.grad is not None:
            self.grad[:] = 0.0

    def update(self):
        if self.grad is not None:
            self._materialize_zero()
            self.value += self.opt_state.grad2diff(self.grad)

    def regular_loss(self) -> Array:
        if self.R is not None:
            return self.R.loss(self.value)
        return zeros(self.value._gpu, shape=(1,))

    def regular_grad(self):
        if self.R is not None:
            self.add_grad(self.R.grad(self.value))
