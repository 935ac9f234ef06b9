"""
Sequential model (reference: vulkpy/nn/models.py).
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple, Union

from ..vkarray import Array
from .core import Module, Loss
from . import optimizers as _opt
from .optimizers import AdamState

__all__ = ["Sequence"]


class Sequence:
    """Layers applied in order plus a loss; ``train`` = forward, loss, zero_grad, backward,
    update (reference: models.py:15-105; regularizers are not applied, as in the reference)."""

    def __init__(self, layers: Iterable[Module], loss: Loss):
        self.L: Tuple[Module, ...] = tuple(layers)
        self.loss: Loss = loss

    @staticmethod
    def _stock(layer, cls) -> bool:
        """`layer` is exactly the library's `cls` (a subclass may override forward / backward)."""
        return type(layer) is cls

    def _forward(self, x: Array) -> Array:
        from .layers import Dense, ReLU
        L = self.L
        i = 0
        while i < len(L):
            layer = L[i]
            if (not _opt.UNFUSED and i + 1 < len(L) and self._stock(layer, Dense) and self._stock(L[i + 1], ReLU)
                    and len(x.shape) == 2):
                # Dense followed by ReLU: one GEMM whose epilogue adds the bias and clamps at zero;
                # both modules remember what their own forward would have (the ReLU's input is not kept:
                # its backward only reads its output, layers.py:207-210)
                y = layer._forward_relu(x)
                layer._x, layer._y = x, y
                L[i + 1]._x = L[i + 1]._y = y
                x = y
                i += 2
                continue
            x = layer(x)
            i += 1
        return x

    def _backward(self):
        """Layers in reverse.  The reference also asks the FIRST layer for the gradient of the
        network input and drops it (models.py:46-53); a layer that can produce its parameter
        gradients alone (``_backward_params``) is spared that contraction."""
        from .layers import Dense, ReLU
        dx = self.loss.grad()
        L = self.L
        first = L[0] if L else None
        i = len(L) - 1
        while i >= 0:
            layer = L[i]
            if layer is first and not _opt.UNFUSED and hasattr(layer, "_backward_params"):
                layer._backward_params(dx)
            elif (not _opt.UNFUSED and i >= 1 and self._stock(layer, Dense) and self._stock(L[i - 1], ReLU)
                  and len(dx.shape) == 2 and layer.input_dim % 4 == 0):
                # the ReLU in front of this Dense: its backward runs in the epilogue of the dx GEMM
                dx = layer.backward(dx, relu_y=L[i - 1]._y)
                i -= 1
            else:
                dx = layer.backward(dx)
            i -= 1

    def _known_parameters(self):
        """(parameters the one-launch zero_grad / update may own, the other layers).

        Only layers whose ``zero_grad`` / ``update`` / ``_parameters`` are the stock ``Dense`` ones are
        taken over: a subclass that overrides them (frozen layer, clipping, a regularizer step) keeps
        being called, as the reference's ``Sequence`` always does (models.py:46-78).  A parameter
        shared by two layers (weight tying) is listed once."""
        from .layers import Dense
        params, rest, seen = [], [], set()
        for layer in self.L:
            t = type(layer)
            stock = (getattr(t, "_parameters", None) is Dense._parameters and t.update is Dense.update
                     and t.zero_grad is Dense.zero_grad)
            if not stock:
                rest.append(layer)
                continue
            for p in layer._parameters():
                if p.grad is not None and id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        return params, rest

    def _zero_grad(self):
        if _opt.UNFUSED:
            for layer in self.L:
                layer.zero_grad()
            return
        params, rest = self._known_parameters()
        for layer in rest:
            layer.zero_grad()
        # all gradients in one device fill (the reference zeroes each through its host view)
        for i in range(0, len(params), 16):
            part = [p.grad for p in params[i:i + 16]]
            job = part[0]._gpu.gpu.fill_many([g.buffer for g in part], 0)
            for g in part:
                g.job = job
                g._keep = []

    def _update(self):
        if _opt.UNFUSED:
            for layer in self.L:
                layer.update()
            return
        params, rest = self._known_parameters()
        for layer in rest:
            layer.update()
        adam = [p for p in params if type(p.opt_state) is AdamState]
        for p in params:
            if type(p.opt_state) is not AdamState:
                p.update()
        # Adam step and `value += diff` of every parameter in one launch, same float32 operations
        # in the same order as AdamState.grad2diff + Parameter.update (nn/optimizers.py:235-253)
        for i in range(0, len(adam), 16):
            part = adam[i:i + 16]
            rows = []
            for p in part:
                st, o = p.opt_state, p.opt_state.opt
                st.beta1t *= o.beta1
                st.beta2t *= o.beta2
                rows.append((o.beta1, 1 - o.beta1, o.beta2, 1 - o.beta2, 1 - st.beta1t, 1 - st.beta2t, o.eps, -o.lr))
            dev = part[0].value._gpu.gpu
            job = dev.nn_adam_apply_many([p.grad.buffer for p in part], [p.opt_state.m.buffer for p in part],
                                         [p.opt_state.v.buffer for p in part], [p.value.buffer for p in part], rows)
            for p in part:
                p.value.job = p.opt_state.m.job = p.opt_state.v.job = job
                p.value._keep = [p.grad]

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
