"""
Losses (reference: vulkpy/nn/losses.py).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Tuple

import contextlib

from ..vkarray import Array, DataShape, VectorParams, fuse
from . import optimizers as _opt
from .core import Loss
from .layers import Softmax

__all__ = ["CrossEntropyLoss", "SoftmaxCrossEntropyLoss", "MSELoss", "HuberLoss", "MixLoss"]


def _chains():
    """Element-wise chains of a loss (`L = y - x; L **= 2`, `dx = x - y; dx *= 2; dx *= 1/B` ...) are recorded
    and issued as one launch each unless the op-by-op reference order is requested (VULKPY_NN_UNFUSED=1)."""
    return contextlib.nullcontext() if _opt.UNFUSED else fuse()


class ReduceLoss(Loss):
    """Per-sample loss (sum over axis 1) reduced over the batch by ``"mean"`` or ``"sum"``;
    with ``"mean"`` the gradient is scaled by 1/batch (reference: losses.py:40-92)."""

    def __init__(self, reduce: str = "mean"):
        if reduce == "mean":
            self.reduce = lambda L: L.mean(axis=0)
            self.scale_backward = lambda dx: 1 / dx.shape[0]
        elif reduce == "sum":
            self.reduce = lambda L: L.sum(axis=0)
            self.scale_backward = None
        else:
            raise KeyError(reduce)

    def __call__(self, x: Array, y: Array) -> Array:
        self._x, self._y = x, y
        with _chains():
            return self.reduce(self.forward(x, y))

    def grad(self) -> Array:
        with _chains():
            dx = self.backward()
            if self.scale_backward is not None:
                dx *= self.scale_backward(dx)
            return dx

    def forward(self, x: Array, y: Array) -> Array:
        raise NotImplementedError

    def backward(self) -> Array:
        raise NotImplementedError


class CrossEntropyLoss(ReduceLoss):
    """L = -sum_j y_j log(x_j + 1e-8); dL/dx = -y / (x + 1e-8)
    (reference: losses.py:95-176, nn_cross_entropy.comp:25, nn_cross_entropy_backward.comp:25)."""

    def _elementwise(self, spv: str, x: Array, y: Array) -> Array:
        n = x.buffer.size()
        out = Array(x._gpu, shape=x.shape)
        out.job = x._gpu._submit(spv, 64, 1, 1, [x, y, out], DataShape(n, 1, 1), VectorParams(n))
        out._keep.extend([x, y])
        return out

    def forward(self, x: Array, y: Array) -> Array:
        return self._elementwise("nn_cross_entropy", x, y).sum(axis=1)

    def backward(self) -> Array:
        return self._elementwise("nn_cross_entropy_backward", self._x, self._y)


class SoftmaxCrossEntropyLoss(CrossEntropyLoss):
    """Softmax followed by cross entropy; gradient softmax(x) - y (reference: losses.py:179-249)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._sm = Softmax()

    def forward(self, x: Array, y: Array) -> Array:
        return super().forward(self._sm(x), y)

    def backward(self) -> Array:
        return self._sm._y - self._y


class MSELoss(ReduceLoss):
    """sum_j (y_j - x_j)^2; gradient 2 (x - y) (reference: losses.py:252-320)."""

    def forward(self, x: Array, y: Array) -> Array:
        L = y - x
        L **= 2.0
        return L.sum(axis=1)

    def backward(self) -> Array:
        dx = self._x - self._y
        dx *= 2
        return dx


class HuberLoss(ReduceLoss):
    """0.5 * min(|d|, d^2) summed over axis 1; gradient clamp(x - y, -1, 1)
    (reference: losses.py:323-393)."""

    def forward(self, x: Array, y: Array) -> Array:
        d = y - x
        d.abs(inplace=True)
        d.min(d ** 2.0, inplace=True)
        d *= 0.5
        return d.sum(axis=1)

    def backward(self) -> Array:
        d = self._x - self._y
        d.clamp(-1.0, 1.0, inplace=True)
        return d


class MixLoss(Loss):
    """Weighted sum of losses (reference: losses.py:396-458)."""

    def __init__(self, losses: Iterable[Tuple[float, Loss]]):
        self.L: Tuple[Tuple[float, Loss], ...] = tuple(losses)
        if len(self.L) < 1:
            raise ValueError("losses should not empty")

    def _sum(self, f: Callable[[Loss], Array]) -> Array:
        coeff, loss = self.L[0]
        total = coeff * f(loss)
        for coeff, loss in self.L[1:]:
            total += coeff * f(loss)
        return total

    def __call__(self, x: Array, y: Array) -> Array:
        return self._sum(lambda loss: loss(x, y))

    def grad(self) -> Array:
        return self._sum(lambda loss: loss.grad())
