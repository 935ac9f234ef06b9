"""GPU tests of vulkpy.nn (layers, losses, optimizers, regularizers, Sequence): known answers in
the style of the reference's test/test_nn.py with its tolerances (rtol = atol = 1e-7 against
float64 for activations and losses) plus oracle-backed checks of Dense and of a training step."""
import numpy as np
import pytest

import vulkpy_b200 as vk
from vulkpy_b200 import nn
from vulkpy_b200.nn.parameters import Parameter
from oracle import vulkpy_oracle as orc

pytestmark = pytest.mark.gpu
F = np.float32
TOL = dict(rtol=1e-7, atol=1e-7)


def A(gpu, x):
    return vk.Array(gpu, data=x)


def softmax64(x):
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def test_relu_sigmoid(gpu):
    x = np.asarray([[-0.2, 0.0, 0.2], [1.5, -3.0, 0.1]])
    relu = nn.ReLU()
    np.testing.assert_allclose(relu(A(gpu, x)), np.maximum(x, 0), **TOL)
    dy = np.asarray([[0.7, 0.8, 0.9], [1.0, 1.1, 1.2]])
    np.testing.assert_allclose(relu.backward(A(gpu, dy)), dy * (x > 0), **TOL)
    sig = nn.Sigmoid()
    y = 1 / (1 + np.exp(-x))
    np.testing.assert_allclose(sig(A(gpu, x)), y, **TOL)
    np.testing.assert_allclose(sig.backward(A(gpu, dy)), dy * y * (1 - y), **TOL)
    with pytest.raises(ValueError):
        relu(A(gpu, [1, 2, 3]))       # modules need at least 2-D input (core.py:247-275)


def test_softmax(gpu):
    sm = nn.Softmax()
    x = np.asarray([[1.0, 1.0]])
    np.testing.assert_allclose(sm(A(gpu, x)), [[0.5, 0.5]], **TOL)
    x = np.asarray([[0.0, 100.0]])
    np.testing.assert_allclose(sm(A(gpu, x)), [[0.0, 1.0]], **TOL)
    # exp(-100) is a float32 subnormal: flushed, so the result is EXACTLY [1, 0] (test/test_nn.py:120-126)
    np.testing.assert_allclose(sm(A(gpu, [[100.0, 0.0]])), [[1.0, 0.0]], rtol=1e-7, atol=0)
    x = np.asarray([[0.1, 0.7, -1.3, 2.0], [3.0, 3.0, 2.0, -8.0]])
    y = softmax64(x)
    np.testing.assert_allclose(sm(A(gpu, x)), y, **TOL)
    dy = np.asarray([[0.1, 0.2, 0.3, 0.4], [1.0, -1.0, 0.5, 2.0]])
    # diagonal of the Jacobian only, as the reference (layers.py:302-323)
    np.testing.assert_allclose(sm.backward(A(gpu, dy)), dy * y * (1 - y), **TOL)


def test_dense_forward_known_answers(gpu):
    d = nn.Dense(gpu, 2, 3, w_init=nn.Constant(0.0), b_init=nn.Constant(0.0))
    np.testing.assert_allclose(d(A(gpu, [[1, 2], [3, 4]])), np.zeros((2, 3)))
    d = nn.Dense(gpu, 2, 3, w_init=nn.Constant(0.0), b_init=nn.Constant(1.5))
    np.testing.assert_allclose(d(A(gpu, [[1, 2], [3, 4]])), np.full((2, 3), 1.5))
    d = nn.Dense(gpu, 2, 2, w_init=lambda g, s: vk.Array(g, data=[[1, 2], [3, 4]]), b_init=nn.Constant(0.5))
    np.testing.assert_allclose(d(A(gpu, [[1, 1], [2, 0]])), [[3.5, 7.5], [2.5, 6.5]])


def test_dense_backward_known_answers(gpu):
    W = np.asarray([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])           # (out=2, in=3)
    d = nn.Dense(gpu, 3, 2, w_init=lambda g, s: vk.Array(g, data=W), b_init=nn.Constant(0.0),
                 w_opt=nn.SGD(0.1), b_opt=nn.SGD(0.1))
    x = np.asarray([[1.0, 0.0, -1.0], [2.0, 1.0, 0.5]])
    dy = np.asarray([[1.0, -1.0], [0.5, 2.0]])
    d(A(gpu, x))
    dx = d.backward(A(gpu, dy))
    np.testing.assert_allclose(dx, dy @ W, rtol=1e-6)
    np.testing.assert_allclose(d.w.grad, dy.T @ x, rtol=1e-6)
    np.testing.assert_allclose(d.b.grad, dy.sum(axis=0), rtol=1e-6)
    assert d._x.shape == (2, 3)
    d.update()
    np.testing.assert_allclose(d.w.value, W - 0.1 * (dy.T @ x), rtol=1e-6)
    d.zero_grad()
    np.testing.assert_array_equal(np.asarray(d.w.grad), np.zeros((2, 3)))


def test_dense_vs_oracle(gpu, rs):
    w = rs.normal(size=(37, 53)).astype(F)
    b = rs.normal(size=37).astype(F)
    x = rs.normal(size=(29, 53)).astype(F)
    d = nn.Dense(gpu, 53, 37, w_init=lambda g, s: vk.Array(g, data=w), b_init=lambda g, s: vk.Array(g, data=b))
    np.testing.assert_allclose(d(A(gpu, x)), orc.batch_affine(w, b, x), rtol=0, atol=2e-5)


@pytest.mark.parametrize("reduce", ["mean", "sum"])
def test_cross_entropy(gpu, reduce):
    x = np.asarray([[0.5, 0.5], [0.2, 0.8], [0.9, 0.1]])
    y = np.asarray([[1.0, 0.0], [0.0, 1.0], [1.0, 0.0]])
    L = nn.CrossEntropyLoss(reduce=reduce)
    want = -(y * np.log(x + 1e-8)).sum(axis=1)
    want = want.mean() if reduce == "mean" else want.sum()
    np.testing.assert_allclose(L(A(gpu, x), A(gpu, y)), want, rtol=1e-6)
    g = -y / (x + 1e-8)
    np.testing.assert_allclose(L.grad(), g / 3 if reduce == "mean" else g, rtol=1e-6)
    np.testing.assert_allclose(nn.CrossEntropyLoss()(A(gpu, [[0.5, 0.5]]), A(gpu, [[1, 0]])), 0.6931472, rtol=1e-6)


@pytest.mark.parametrize("reduce", ["mean", "sum"])
def test_softmax_cross_entropy(gpu, reduce):
    x = np.asarray([[1.0, 2.0, 3.0], [1.0, 1.0, 1.0]])
    y = np.asarray([[0.0, 0.0, 1.0], [1.0, 0.0, 0.0]])
    L = nn.SoftmaxCrossEntropyLoss(reduce=reduce)
    p = softmax64(x)
    want = -(y * np.log(p + 1e-8)).sum(axis=1)
    want = want.mean() if reduce == "mean" else want.sum()
    np.testing.assert_allclose(L(A(gpu, x), A(gpu, y)), want, rtol=1e-6)
    g = p - y
    np.testing.assert_allclose(L.grad(), g / 2 if reduce == "mean" else g, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("reduce", ["mean", "sum"])
def test_mse_and_huber(gpu, reduce):
    x = np.asarray([[1.0, 2.0], [0.5, -3.0], [4.0, 4.0]])
    y = np.asarray([[1.5, 0.0], [0.5, -1.0], [0.0, 4.5]])
    red = (lambda v: v.mean()) if reduce == "mean" else (lambda v: v.sum())
    scale = 1 / 3 if reduce == "mean" else 1.0
    L = nn.MSELoss(reduce=reduce)
    np.testing.assert_allclose(L(A(gpu, x), A(gpu, y)), red(((y - x) ** 2).sum(axis=1)), **TOL)
    np.testing.assert_allclose(L.grad(), 2 * (x - y) * scale, rtol=1e-6)
    H = nn.HuberLoss(reduce=reduce)
    d = np.abs(y - x)
    np.testing.assert_allclose(H(A(gpu, x), A(gpu, y)), red((0.5 * np.minimum(d, d ** 2)).sum(axis=1)), rtol=1e-6)
    np.testing.assert_allclose(H.grad(), np.clip(x - y, -1, 1) * scale, rtol=1e-6)


def test_mix_loss(gpu):
    x, y = A(gpu, [[1.0, 2.0]]), A(gpu, [[0.0, 0.0]])
    m = nn.MixLoss([(0.5, nn.MSELoss()), (2.0, nn.HuberLoss())])
    np.testing.assert_allclose(m(x, y), 0.5 * 5.0 + 2.0 * (0.5 * (1 + 2)), rtol=1e-6)
    np.testing.assert_allclose(m.grad(), 0.5 * np.asarray([[2.0, 4.0]]) + 2.0 * np.asarray([[1.0, 1.0]]), rtol=1e-6)
    with pytest.raises(ValueError):
        nn.MixLoss([])


def test_optimizers(gpu):
    g = A(gpu, [1.0, -2.0, 0.5])
    np.testing.assert_allclose(nn.SGD(0.01).init_state((3,)).grad2diff(g), [-0.01, 0.02, -0.005], rtol=1e-6)
    ada = nn.AdaGrad(gpu, lr=0.1, tau=0.0, eps=1e-8).init_state((3,))
    np.testing.assert_allclose(ada.grad2diff(g), [-0.1, 0.1, -0.1], rtol=1e-5)
    np.testing.assert_allclose(ada.h, [1.0, 4.0, 0.25], rtol=1e-6)
    st = nn.Adam(gpu, lr=0.001).init_state((3,))
    d1 = np.asarray(st.grad2diff(g)).copy()
    np.testing.assert_allclose(d1, [-0.001, 0.001, -0.001], rtol=1e-4)   # first step: -lr * sign(g)
    np.testing.assert_allclose(st.m, 0.1 * np.asarray([1.0, -2.0, 0.5]), rtol=1e-6)
    np.testing.assert_allclose(st.v, 0.001 * np.asarray([1.0, 4.0, 0.25]), rtol=1e-5)
    assert abs(st.beta1t - 0.9) < 1e-12 and abs(st.beta2t - 0.999) < 1e-12


def test_adam_matches_float32_restatement(gpu, rs):
    """Every intermediate of optimizers.py:235-253 is rounded to float32 where the reference does."""
    g = rs.normal(size=257).astype(F)
    st = nn.Adam(gpu, lr=1e-3).init_state((257,))
    m = np.zeros(257, F)
    v = np.zeros(257, F)
    b1t = b2t = 1.0
    for _ in range(3):
        got = np.asarray(st.grad2diff(A(gpu, g))).copy()
        m = (m * F(0.9)).astype(F)
        m = (m + (g * F(1 - 0.9)).astype(F)).astype(F)
        v = (v * F(0.999)).astype(F)
        v = (v + (orc.scalar("pow", g, 2.0) * F(1 - 0.999)).astype(F)).astype(F)
        b1t *= 0.9
        b2t *= 0.999
        mhat = (m / F(1 - b1t)).astype(F)
        vhat = (v / F(1 - b2t)).astype(F)
        vhat = (np.sqrt(vhat).astype(F) + F(1e-8)).astype(F)
        want = ((mhat * F(-1e-3)).astype(F) / vhat).astype(F)
        np.testing.assert_array_equal(got, want)


def test_parameter_and_regularizers(gpu):
    p = Parameter(gpu, shape=(2, 2), opt=nn.SGD(0.5), initializer=nn.Constant(1.0), regularizer=nn.Ridge(0.1))
    assert p.is_trainable()
    p.add_grad(A(gpu, [[1, 2], [3, 4]]))
    p.regular_grad()                    # + 2 * 0.1 * value
    np.testing.assert_allclose(p.grad, [[1.2, 2.2], [3.2, 4.2]], rtol=1e-6)
    np.testing.assert_allclose(p.regular_loss(), [0.4], rtol=1e-6)
    p.update()
    np.testing.assert_allclose(p.value, 1 - 0.5 * np.asarray([[1.2, 2.2], [3.2, 4.2]]), rtol=1e-6)
    p.zero_grad()
    np.testing.assert_array_equal(np.asarray(p.grad), np.zeros((2, 2)))
    q = Parameter(gpu, shape=(2,), trainable=False)
    assert not q.is_trainable() and q.opt_state is None
    np.testing.assert_allclose(q.regular_loss(), [0.0])
    w = A(gpu, [-2.0, 0.0, 3.0])
    np.testing.assert_allclose(nn.Lasso(2.0).loss(w), [10.0])
    np.testing.assert_allclose(nn.Lasso(2.0).grad(w), [-2.0, 0.0, 2.0])
    np.testing.assert_allclose(nn.Ridge(0.5).loss(w), [6.5])
    np.testing.assert_allclose(nn.Ridge(0.5).grad(w), [-2.0, 0.0, 3.0])
    np.testing.assert_allclose(nn.Elastic(1.0, 1.0).loss(w), [18.0])
    np.testing.assert_allclose(nn.Elastic(1.0, 1.0).grad(w), [-5.0, 0.0, 7.0])


def test_sequence_trains(gpu, rs):
    """Dense-ReLU-Dense trained with SoftmaxCrossEntropyLoss (exact gradient softmax(x) - y) learns a
    separable toy problem; Softmax + CrossEntropyLoss keeps the reference's diagonal-only backward
    (layers.py:302-323), which is checked step-wise in the next test instead."""
    opt = nn.Adam(gpu, lr=1e-2)
    net = nn.Sequence([nn.Dense(gpu, 4, 32, w_opt=opt, b_opt=opt, w_init=nn.HeNormal(gpu, 4, seed=1)), nn.ReLU(),
                       nn.Dense(gpu, 32, 3, w_opt=opt, b_opt=opt, w_init=nn.HeNormal(gpu, 32, seed=2))],
                      nn.SoftmaxCrossEntropyLoss())
    x = rs.normal(size=(96, 4)).astype(F)
    labels = (x[:, 0] > 0).astype(np.uint32) + (x[:, 1] > 0.5).astype(np.uint32)
    X, Y = A(gpu, x), vk.U32Array(gpu, data=labels).to_onehot(3)
    _, first = net.train(X, Y)
    first = float(np.asarray(first).reshape(-1)[0])
    for _ in range(150):
        pred, loss = net.train(X, Y)
    last = float(np.asarray(loss).reshape(-1)[0])
    assert last < 0.5 * first
    p, l2 = net.predict(X, Y)
    assert np.asarray(p).shape == (96, 3) and np.asarray(net.predict(X)).shape == (96, 3)
    acc = (np.asarray(p).argmax(axis=1) == labels).mean()
    assert acc > 0.8


def test_one_training_step_matches_float64_model(gpu, rs):
    """Forward + backward + SGD update of Dense-ReLU-Dense-Softmax/CE against a float64 NumPy model."""
    W1, b1 = rs.normal(size=(16, 8)) * 0.5, rs.normal(size=16) * 0.1
    W2, b2 = rs.normal(size=(4, 16)) * 0.5, rs.normal(size=4) * 0.1
    x = rs.normal(size=(32, 8))
    y = np.eye(4)[rs.integers(0, 4, 32)]
    sgd = nn.SGD(0.1)
    mk = lambda v: (lambda g, s: vk.Array(g, data=v))
    d1 = nn.Dense(gpu, 8, 16, w_init=mk(W1), b_init=mk(b1), w_opt=sgd, b_opt=sgd)
    d2 = nn.Dense(gpu, 16, 4, w_init=mk(W2), b_init=mk(b2), w_opt=sgd, b_opt=sgd)
    net = nn.Sequence([d1, nn.ReLU(), d2, nn.Softmax()], nn.CrossEntropyLoss())
    pred, loss = net.train(A(gpu, x), A(gpu, y))
    h = np.maximum(x @ W1.T + b1, 0)
    p = softmax64(h @ W2.T + b2)
    np.testing.assert_allclose(pred, p, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(loss, -(y * np.log(p + 1e-8)).sum(axis=1).mean(), rtol=2e-5)
    dp = -y / (p + 1e-8) / 32
    dz2 = dp * p * (1 - p)                      # the reference's diagonal softmax backward
    np.testing.assert_allclose(d2.w.value, W2 - 0.1 * dz2.T @ h, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(d2.b.value, b2 - 0.1 * dz2.sum(axis=0), rtol=1e-4, atol=1e-5)
    dz1 = (dz2 @ W2) * (h > 0)
    np.testing.assert_allclose(d1.w.value, W1 - 0.1 * dz1.T @ x, rtol=1e-4, atol=1e-5)


def test_fused_kernels_match_the_op_by_op_compositions(gpu, rs, monkeypatch):
    """vkp_nn_adam / activation backward are bit-identical to the reference-ordered op chains; the
    fused softmax agrees with the float64 model and (up to the summation order) with the chain."""
    from vulkpy_b200.nn import optimizers as O
    g = rs.normal(size=1000).astype(F)
    y = rs.uniform(-1, 1, (37, 19)).astype(F)
    dy = rs.normal(size=(37, 19)).astype(F)
    x = rs.normal(size=(65, 40)).astype(F) * 3
    res = {}
    for unfused in (True, False):
        monkeypatch.setattr(O, "UNFUSED", unfused)
        st = nn.Adam(gpu, lr=1e-3).init_state((1000,))
        outs = [np.asarray(st.grad2diff(A(gpu, g * (k + 1)))).copy() for k in range(3)]
        relu, sig, sm = nn.ReLU(), nn.Sigmoid(), nn.Softmax()
        relu._y = sig._y = sm._y = A(gpu, y)
        res[unfused] = (outs, np.asarray(st.m).copy(), np.asarray(st.v).copy(),
                        np.asarray(relu.backward(A(gpu, dy))).copy(), np.asarray(sig.backward(A(gpu, dy))).copy(),
                        np.asarray(sm.backward(A(gpu, dy))).copy(), np.asarray(sm.forward(A(gpu, x))).copy(), x)
    a, b = res[True], res[False]
    for u, f in zip(a[0], b[0]):
        np.testing.assert_array_equal(u, f)
    for k, what in ((1, "adam m"), (2, "adam v"), (3, "relu backward"), (4, "sigmoid backward"), (5, "softmax backward")):
        np.testing.assert_array_equal(a[k], b[k], err_msg=what)
    np.testing.assert_allclose(b[6], softmax64(b[7].astype(np.float64)), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(a[6], b[6], rtol=1e-6, atol=1e-9)


def _mlp(gpu, d_in, hidden, classes):
    opt = lambda: nn.Adam(gpu, lr=1e-2)
    return nn.Sequence([nn.Dense(gpu, d_in, hidden, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, d_in, seed=3)),
                        nn.ReLU(),
                        nn.Dense(gpu, hidden, classes, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, hidden, seed=4)),
                        nn.Softmax()], nn.CrossEntropyLoss())


def _state(net):
    out = []
    for layer in (net.L[0], net.L[2]):
        for p in (layer.w, layer.b):
            out += [np.asarray(p.value).copy(), np.asarray(p.grad).copy(),
                    np.asarray(p.opt_state.m).copy(), np.asarray(p.opt_state.v).copy()]
    return out


@pytest.mark.parametrize("dims", [(33, 20, 24, 5), (256, 128, 256, 16)])
def test_sequence_step_one_launch_paths_equal_per_layer_calls(gpu, rs, dims):
    """Sequence.train (one-launch zero_grad, one-launch Adam + `value += diff`, no input gradient for
    the first layer) leaves bit-identical parameters, gradients, moments and loss to the same step
    driven layer by layer through the public methods the reference's Sequence calls
    (nn/models.py:37-78: layer.zero_grad(), layer.backward(dx) for EVERY layer, layer.update())."""
    B, d_in, hidden, classes = dims
    x = rs.normal(size=(B, d_in)).astype(F)
    y = np.eye(classes, dtype=F)[rs.integers(0, classes, B)]
    a, b = _mlp(gpu, d_in, hidden, classes), _mlp(gpu, d_in, hidden, classes)
    for _ in range(3):
        _, la = a.train(vk.Array(gpu, data=x), vk.Array(gpu, data=y))
        pred = b._forward(vk.Array(gpu, data=x))
        lb = b.loss(pred, vk.Array(gpu, data=y))
        for layer in b.L:
            layer.zero_grad()
        dx = b.loss.grad()
        for layer in reversed(b.L):
            dx = layer.backward(dx)
        for layer in b.L:
            layer.update()
        np.testing.assert_array_equal(np.asarray(la), np.asarray(lb))
    for k, (u, f) in enumerate(zip(_state(a), _state(b))):
        np.testing.assert_array_equal(u, f, err_msg=f"state {k}")


@pytest.mark.parametrize("dims", [(33, 20, 24), (512, 256, 128)])
def test_dense_backward_epilogue_accumulate_equals_add_grad(gpu, rs, monkeypatch, dims):
    """`grad += dy^T x` inside the GEMM epilogue == GEMM into a temporary followed by add_grad
    (nn/layers.py:126-141, nn/parameters.py:69-79), bit for bit, also on a non-zero gradient."""
    import vulkpy_b200.nn.optimizers as O
    B, d_in, d_out = dims
    x = rs.normal(size=(B, d_in)).astype(F)
    dy = rs.normal(size=(B, d_out)).astype(F)
    res = {}
    for unfused in (True, False):
        monkeypatch.setattr(O, "UNFUSED", unfused)
        layer = nn.Dense(gpu, d_in, d_out, w_init=nn.HeNormal(gpu, d_in, seed=9))
        layer(vk.Array(gpu, data=x))
        dxs = [np.asarray(layer.backward(vk.Array(gpu, data=dy))).copy() for _ in range(2)]   # second call accumulates
        res[unfused] = dxs + [np.asarray(layer.w.grad).copy(), np.asarray(layer.b.grad).copy()]
    for u, f in zip(res[True], res[False]):
        np.testing.assert_array_equal(u, f)
    want = 2 * (dy.astype(np.float64).T @ x.astype(np.float64))
    np.testing.assert_allclose(res[False][2], want, rtol=1e-4, atol=1e-4)
