"""
Layers (reference: vulkpy/nn/layers.py).
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional

from ..vkarray import GPU, Array, DataShape, BatchAffineParams, fuse
from .core import Module, Optimizer, Regularizer
from .parameters import Parameter
from .initializers import HeNormal
from . import optimizers as _opt

__all__ = ["Dense", "ReLU", "Sigmoid", "Softmax"]


def _fused_backward(kind: int, y: Array, dy: Array) -> Array:
    dx = Array(dy._gpu, shape=y.shape)
    dx.job = dy._gpu.gpu.nn_activation_backward(kind, y.buffer, dy.buffer, dx.buffer)
    dx._keep = [y, dy]
    return dx

Init = Callable[[GPU, Iterable[int]], Array]
_GEMM_ACCUMULATE = 4   # VKP_GEMM_ACCUMULATE (include/vulkpy_b200.h)
_GEMM_RELU = 8         # VKP_GEMM_RELU


class Dense(Module):
    """Fully connected layer ``y = x W^T + b`` with ``W`` of shape ``(output_dim, input_dim)``
    (reference: layers.py:18-155)."""

    def __init__(self, gpu: GPU, input_dim: int, output_dim: int, *,
                 w_init: Optional[Init] = None, b_init: Optional[Init] = None,
                 w_opt: Optional[Optimizer] = None, b_opt: Optional[Optimizer] = None,
                 w_reg: Optional[Regularizer] = None, b_reg: Optional[Regularizer] = None):
        self.input_dim = int(input_dim)
        self.output_dim = int(output_dim)
        if w_init is None:
            w_init = HeNormal(gpu, self.input_dim)
        self.w = Parameter(gpu, shape=(self.output_dim, self.input_dim), initializer=w_init,
                           opt=w_opt, regularizer=w_reg)
        self.b = Parameter(gpu, shape=(self.output_dim,), initializer=b_init, opt=b_opt,
                           regularizer=b_reg)

    def forward(self, x: Array) -> Array:
        """One ``batch_affine`` kernel: GEMM with the bias added in its epilogue
        (reference: layers.py:71-102, batch_affine.comp:25-39)."""
        batch = x.shape[0]
        y = Array(x._gpu, shape=(batch, self.output_dim))
        y.job = x._gpu._submit("batch_affine", 1, 64, 1, [self.w.value, self.b.value, x, y],
                               DataShape(batch, self.output_dim, 1),
                               BatchAffineParams(batch, x.shape[1], self.output_dim))
        y._keep.extend([self.w.value, self.b.value, x])
        return y

    def _forward_relu(self, x: Array) -> Array:
        """``ReLU()(self(x))`` as one kernel: the GEMM epilogue adds the bias and applies ``max(., 0)`` -- the
        float32 operations of batch_affine.comp:25-39 followed by max_scalar.comp (``x.max(0.0)``, layers.py:186)."""
        batch = x.shape[0]
        y = Array(x._gpu, shape=(batch, self.output_dim))
        y.job = x._gpu.gpu.gemm(False, True, batch, self.output_dim, x.shape[1], x.buffer, self.w.value.buffer, y.buffer,
                                self.b.value.buffer, _GEMM_RELU)
        y._keep.extend([self.w.value, self.b.value, x])
        return y

    def backward(self, dy: Array, relu_y: Optional[Array] = None) -> Array:
        """db = sum_batch dy, dW = dy^T x, dx = dy W.

        The reference forms dW by materialising ``dy[:, :, None] * x[:, None, :]``
        (batch x out x in) and summing over the batch (layers.py:126-141); the same
        contraction is one GEMM here.  ``relu_y`` (additive): the output of a ReLU that feeds this
        layer; its backward ``max(sign(y), 0) * dx`` (layers.py:207-210) then runs in the epilogue
        of the ``dx`` GEMM and the returned gradient is already the ReLU's input gradient."""
        self._backward_params(dy)
        if relu_y is None:
            return dy @ self.w.value
        dx = Array(dy._gpu, shape=(dy.shape[0], self.input_dim))
        dx.job = dy._gpu.gpu.gemm(False, False, dy.shape[0], self.input_dim, self.output_dim, dy.buffer,
                                  self.w.value.buffer, dx.buffer, None, 0, relu_mask=relu_y.buffer)
        dx._keep = [dy, self.w.value, relu_y]
        return dx

    def _backward_params(self, dy: Array):
        """Parameter gradients only (what ``Sequence`` needs from its first layer)."""
        if _opt.UNFUSED:
            self.b.add_grad(dy.sum(axis=0))
        else:
            self.b._take_grad(dy.sum(axis=0))
        x = self._x
        dev = dy._gpu.gpu
        g = self.w.grad
        if g is not None and not _opt.UNFUSED:
            # `grad += dy^T x` in the GEMM epilogue: the accumulated float32 product is added to
            # the old value with one rounding, exactly what add_grad's `+=` does to a temporary; a gradient
            # still marked zero (Sequence._zero_grad) is simply overwritten
            fresh, self.w._fresh = self.w._fresh, False
            g.job = dev.gemm(True, False, self.output_dim, self.input_dim, dy.shape[0],
                             dy.buffer, x.buffer, g.buffer, None, 0 if fresh else _GEMM_ACCUMULATE)
            g._keep = [dy, x]
            return
        dW = Array(dy._gpu, shape=(self.output_dim, self.input_dim))
        dW.job = dev.gemm(True, False, self.output_dim, self.input_dim, dy.shape[0],
                          dy.buffer, x.buffer, dW.buffer)
        dW._keep = [dy, x]
        self.w.add_grad(dW)

    def _parameters(self):
        return [self.w, self.b]

    def zero_grad(self):
        self.w.zero_grad()
        self.b.zero_grad()

    def update(self):
        self.w.update()
        self.b.update()


class ReLU(Module):
    """max(x, 0); backward passes dy where the output is positive (reference: layers.py:158-210)."""

    def forward(self, x: Array) -> Array:
        return x.max(0.0)

    def backward(self, dy: Array) -> Array:
        if not _opt.UNFUSED:
            return _fused_backward(0, self._y, dy)
        dx = self._y.sign()
        dx.max(0.0, inplace=True)
        dx *= dy
        return dx


class Sigmoid(Module):
    """1 / (1 + exp(-x)) (reference: layers.py:213-267)."""

    def forward(self, x: Array) -> Array:
        if not _opt.UNFUSED:
            # the reference's four jobs (layers.py:239-243) recorded and issued as ONE chain launch
            # (vkp_ew_chain): same operations, same roundings, 8 B per element instead of 32
            with fuse():
                y = 0.0 - x
                y.exp(inplace=True)
                y += 1.0
                return 1.0 / y
        y = 0.0 - x
        y.exp(inplace=True)
        y += 1.0
        return 1.0 / y

    def backward(self, dy: Array) -> Array:
        if not _opt.UNFUSED:
            return _fused_backward(1, self._y, dy)
        dx = 1.0 - self._y
        dx *= self._y
        dx *= dy
        return dx


class Softmax(Module):
    """Row-wise softmax over axis 1 with the max subtracted first; backward uses the diagonal
    of the Jacobian only, like the reference (layers.py:270-323)."""

    def forward(self, x: Array) -> Array:
        if not _opt.UNFUSED and len(x.shape) == 2:
            y = Array(x._gpu, shape=x.shape)
            y.job = x._gpu.gpu.nn_softmax_forward(x.buffer, y.buffer, x.shape[0], x.shape[1])
            y._keep = [x]
            return y
        e = x - x.maximum(axis=1, rebroadcast=True)
        e.exp(inplace=True)
        e /= e.sum(axis=1, rebroadcast=True)
        return e

    def backward(self, dy: Array) -> Array:
        if not _opt.UNFUSED:
            return _fused_backward(1, self._y, dy)
        dx = 1.0 - self._y
        dx *= self._y
        dx *= dy
        return dx
