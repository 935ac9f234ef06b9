"""
Sequential model (reference: vulkpy/nn/models.py).
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple, Union

from ..vkarray import Array
from .core import Module, Loss

__all__ = ["Sequence"]


class Sequence:
    """Layers applied in order plus a loss; ``train`` = forward, loss, zero_grad, backward,
    update (reference: models.py:15-105; regularizers are not applied, as in the reference)."""

    def __init__(self, layers: Iterable[Module], loss: Loss):
        self.L: Tuple[Module, ...] = tuple(layers)
        self.loss: Loss = loss

    def _forward(self, x: Array) -> Array:
        for layer in self.L:
            x = layer(x)
        return x

    def _backward(self):
        dx = self.loss.grad()
        for layer in reversed(self.L):
            dx = layer.backward(dx)

    def _zero_grad(self):
        for layer in self.L:
            layer.zero_grad()

    def _update(self):
        for layer in self.L:
            layer.update()

    def train(self, x: Array, y: Array) -> Tuple[Array, Array]:
        pred = self._forward(x)
        loss = self.loss(pred, y)
        self._zero_grad()
        self._backward()
        self._update()
        return pred, loss

    def predict(self, x: Array, y: Optional[Array] = None) -> Union[Array, Tuple[Array, Array]]:
        pred = self._forward(x)
        if y is None:
            return pred
        return pred, self.loss(pred, y)
