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

    def _forward(self, x: Array, upto: Optional[int] = None) -> Array:
        from .layers import Dense, ReLU
        L = self.L if upto is None else self.L[:upto]
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

    def _forward_loss(self, x: Array, y: Array) -> Tuple[Array, Array]:
        """``pred = forward(x); loss = self.loss(pred, y)`` of a training step.

        A stock ``Softmax`` last layer under a stock ``CrossEntropyLoss`` runs as ONE launch
        (``vkp_nn_softmax_ce_train``) that also produces the gradient with respect to the Softmax input --
        the float32 operations of Softmax.forward (layers.py:297-300), nn_cross_entropy.comp:25,
        nn_cross_entropy_backward.comp:25 with the 1/batch scale of losses.py:58-68 and Softmax.backward
        (layers.py:320-323), in that order, bit-identical to the five launches it replaces; the modules are
        left in the state their own calls would have produced.  ``_backward`` picks the gradient up."""
        from .layers import Softmax
        from .losses import CrossEntropyLoss, _chains
        self._tail_dz = None
        L, loss = self.L, self.loss
        if (not _opt.UNFUSED and len(L) >= 2 and self._stock(L[-1], Softmax) and type(loss) is CrossEntropyLoss
                and len(y.shape) == 2):
            z = self._forward(x, upto=len(L) - 1)
            if len(z.shape) == 2 and tuple(z.shape) == tuple(y.shape):
                rows, cols = z.shape
                p, Lij, dz = (Array(z._gpu, shape=z.shape) for _ in range(3))
                scale = loss.scale_backward(z) if loss.scale_backward is not None else None
                job = z._gpu.gpu.nn_softmax_ce_train(z.buffer, y.buffer, p.buffer, Lij.buffer, dz.buffer, rows, cols, scale)
                for a in (p, Lij, dz):
                    a.job = job
                    a._keep = [z, y]
                sm = L[-1]
                sm._x, sm._y = z, p
                loss._x, loss._y = p, y
                with _chains():
                    value = loss.reduce(Lij.sum(axis=1))
                self._tail_dz = dz
                return p, value
            pred = L[-1](z)
        else:
            pred = self._forward(x)
        return pred, loss(pred, y)

    def _backward(self):
        """Layers in reverse.  The reference also asks the FIRST layer for the gradient of the
        network input and drops it (models.py:46-53); a layer that can produce its parameter
        gradients alone (``_backward_params``) is spared that contraction."""
        from .layers import Dense, ReLU
        L = self.L
        first = L[0] if L else None
        i = len(L) - 1
        dz, self._tail_dz = getattr(self, "_tail_dz", None), None
        if dz is not None:          # loss.grad() and Softmax.backward already ran inside _forward_loss's launch
            dx = dz
            i -= 1
        else:
            dx = self.loss.grad()
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
        if not _opt.UNFUSED:        # parameters no layer wrote a gradient to: the deferred zero fill happens now
            for p in self._known_parameters()[0]:
                p._materialize_zero()

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
        # No launch at all: the gradients are marked "zero" and the first contribution of the backward pass
        # that follows OVERWRITES instead of accumulating (0 + x = x in float32; the reference zeroes each
        # gradient through its host view, parameters.py:81-86).  Whatever received nothing is filled at the
        # end of `_backward`; any other access goes through Parameter.add_grad / zero_grad, which honour the mark.
        for p in params:
            p._fresh = True

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
        pred, loss = self._forward_loss(x, y)
        self._zero_grad()
        self._backward()
        self._update()
        return pred, loss

    def predict(self, x: Array, y: Optional[Array] = None) -> Union[Array, Tuple[Array, Array]]:
        pred = self._forward(x)
        if y is None:
            return pred
        return pred, self.loss(pred, y)
