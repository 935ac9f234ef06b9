"""
Optimizers (reference: vulkpy/nn/optimizers.py).

Each ``grad2diff`` issues the same element-wise operations in the same order as the reference,
so that every intermediate is rounded to float32 exactly where the reference rounds it.
"""
from __future__ import annotations

import logging
import os
from typing import Iterable

from ..vkarray import GPU, Array, zeros, fuse
from .core import Optimizer, OptimizerState

__all__ = ["SGD", "SGDState", "AdaGrad", "AdaGradState", "Adam", "AdamState", "Optimizer", "OptimizerState"]

logger = logging.getLogger("vulkpy")

# VULKPY_NN_UNFUSED=1 keeps the reference's op-by-op compositions (tests compare both paths)
UNFUSED = os.environ.get("VULKPY_NN_UNFUSED", "0") == "1"


class SGDState(OptimizerState):
    """diff = -lr * grad (reference: optimizers.py:24-53)."""

    def __init__(self, opt: "SGD"):
        self.opt = opt

    def grad2diff(self, grad: Array) -> Array:
        return (-self.opt.lr) * grad


class SGD(Optimizer):
    def __init__(self, lr: float):
        self.lr = lr
        logger.debug("SGD(lr=%f)", self.lr)

    def init_state(self, shape: Iterable[int]) -> SGDState:
        return SGDState(self)


class AdaGradState(OptimizerState):
    """h += g^2; diff = -lr * g / (sqrt(h) + eps) (reference: optimizers.py:99-140)."""

    def __init__(self, opt: "AdaGrad", shape: Iterable[int], tau: float):
        self.opt = opt
        self.h: Array = zeros(opt.gpu, shape=shape)
        self.h[:] = tau

    def grad2diff(self, grad: Array) -> Array:
        if not UNFUSED:
            with fuse():     # h += g^2 and diff = g / (sqrt(h) + eps) * -lr: two chain launches instead of five jobs
                self.h += grad ** 2
                self.h.wait()
                root = self.h.sqrt()
                root += self.opt.eps
                diff = grad / root
                diff *= -self.opt.lr
                return diff
        self.h += grad ** 2
        root = self.h.sqrt()
        root += self.opt.eps
        diff = grad / root
        diff *= -self.opt.lr
        return diff


class AdaGrad(Optimizer):
    def __init__(self, gpu: GPU, *, lr: float = 0.01, tau: float = 0.0, eps: float = 1e-8):
        self.gpu, self.lr, self.tau, self.eps = gpu, lr, tau, eps
        logger.debug("AdaGrad(lr=%f, tau=%f, eps=%f)", lr, tau, eps)

    def init_state(self, shape: Iterable[int]) -> AdaGradState:
        return AdaGradState(opt=self, shape=shape, tau=self.tau)


class AdamState(OptimizerState):
    """Adam moments ``m``, ``v`` and the running powers ``beta1t``, ``beta2t``
    (reference: optimizers.py:200-253)."""

    def __init__(self, opt: "Adam", shape: Iterable[int]):
        self.opt = opt
        self.m: Array = zeros(opt.gpu, shape=shape)
        self.v: Array = zeros(opt.gpu, shape=shape)
        self.beta1t: float = 1.0
        self.beta2t: float = 1.0

    def grad2diff(self, grad: Array) -> Array:
        o = self.opt
        if not UNFUSED:
            # one kernel with the same operations, order and float32 roundings as the chain below
            self.beta1t *= o.beta1
            self.beta2t *= o.beta2
            diff = Array(grad._gpu, shape=grad.shape)
            diff.job = grad._gpu.gpu.nn_adam(grad.buffer, self.m.buffer, self.v.buffer, diff.buffer,
                                              o.beta1, 1 - o.beta1, o.beta2, 1 - o.beta2,
                                              1 - self.beta1t, 1 - self.beta2t, o.eps, -o.lr)
            diff._keep = [grad, self.m, self.v]
            self.m.job = self.v.job = diff.job
            return diff
        self.m *= o.beta1
        self.m += (1 - o.beta1) * grad
        self.v *= o.beta2
        self.v += (1 - o.beta2) * (grad ** 2)
        self.beta1t *= o.beta1
        self.beta2t *= o.beta2
        mhat = self.m / (1 - self.beta1t)
        vhat = self.v / (1 - self.beta2t)
        vhat.sqrt(inplace=True)
        vhat += o.eps
        mhat *= -o.lr
        mhat /= vhat
        return mhat


class Adam(Optimizer):
    def __init__(self, gpu: GPU, *, lr: float = 0.001, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8):
        self.gpu, self.lr, self.beta1, self.beta2, self.eps = gpu, lr, beta1, beta2, eps
        logger.debug("Adam(lr=%f, beta1=%f, beta2=%f, eps=%f)", lr, beta1, beta2, eps)

    def init_state(self, shape: Iterable[int]) -> AdamState:
        return AdamState(opt=self, shape=shape)
