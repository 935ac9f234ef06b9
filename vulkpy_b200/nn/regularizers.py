"""
Regularizers (reference: vulkpy/nn/regularizers.py).
"""
from __future__ import annotations

import logging

from ..vkarray import Array
from .core import Regularizer

__all__ = ["Lasso", "Ridge", "Elastic"]

logger = logging.getLogger("vulkpy")


class Lasso(Regularizer):
    """L1: ``coeff * sum|W|`` (reference: regularizers.py:23-77)."""

    def __init__(self, coeff: float = 1.0):
        logger.debug("Lasso(L1=%s)", coeff)
        self.coeff = coeff

    def loss(self, param: Array) -> Array:
        L = param.abs().sum()
        L *= self.coeff
        return L

    def grad(self, param: Array) -> Array:
        return self.coeff * param.sign()


class Ridge(Regularizer):
    """L2: ``coeff * sum W^2`` (reference: regularizers.py:80-134)."""

    def __init__(self, coeff: float = 1.0):
        logger.debug("Ridge(L2=%s)", coeff)
        self.coeff = coeff

    def loss(self, param: Array) -> Array:
        L = (param ** 2).sum()
        L *= self.coeff
        return L

    def grad(self, param: Array) -> Array:
        return (2 * self.coeff) * param


class Elastic(Regularizer):
    """L1 + L2 (reference: regularizers.py:137-192)."""

    def __init__(self, L1: float = 1.0, L2: float = 1.0):
        self.L1 = Lasso(L1)
        self.L2 = Ridge(L2)

    def loss(self, param: Array) -> Array:
        return self.L1.loss(param) + self.L2.loss(param)

    def grad(self, param: Array) -> Array:
        return self.L1.grad(param) + self.L2.grad(param)
