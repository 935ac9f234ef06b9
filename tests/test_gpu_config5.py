"""ONE training step of BASELINE.json config 5 at its real shape -- Sequence[Dense(1024,1024), ReLU,
Dense(1024,16), Softmax] + CrossEntropyLoss + Adam, batch 8192 -- the shape that takes the split-K tcgen05
GEMM, the skinny Dense kernels and the one-launch Adam step together (reference: nn/models.py:55-78,
nn/layers.py:71-141, nn/optimizers.py:235-253).  Checked against a float64 NumPy model of the same
step and against the reference-ordered op-by-op path (VULKPY_NN_UNFUSED=1)."""
import numpy as np
import pytest

import vulkpy_b200 as vk
from vulkpy_b200 import nn

pytestmark = pytest.mark.gpu
F = np.float32
B, D, H, C = 8192, 1024, 1024, 16
LR, B1, B2, EPS = 1e-3, 0.9, 0.999, 1e-8


def make_net(gpu):
    opt = lambda: nn.Adam(gpu, lr=LR, beta1=B1, beta2=B2, eps=EPS)
    return nn.Sequence([nn.Dense(gpu, D, H, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, D, seed=1)), nn.ReLU(),
                        nn.Dense(gpu, H, C, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, H, seed=2)), nn.Softmax()],
                       nn.CrossEntropyLoss())


def params(net):
    d1, d2 = net.L[0], net.L[2]
    return [d1.w, d1.b, d2.w, d2.b]


@pytest.fixture(scope="module")
def batch():
    rs = np.random.default_rng(5)
    x = rs.normal(size=(B, D)).astype(F)
    y = np.eye(C, dtype=F)[rs.integers(0, C, B)]
    return x, y


def test_config5_step_against_float64_model(gpu, batch):
    x, y = batch
    net = make_net(gpu)
    W1, b1, W2, b2 = [np.asarray(p.value).astype(np.float64) for p in params(net)]
    pred, loss = net.train(vk.Array(gpu, data=x), vk.Array(gpu, data=y))
    # ---- float64 model of the reference's step (diagonal softmax Jacobian: nn/layers.py:302-323) ----
    x64 = x.astype(np.float64)
    z1 = x64 @ W1.T + b1
    # ReLU mask taken from the device: float32 rounding flips the sign of a handful of the 8.4 M
    # pre-activations that sit within ~1e-6 of zero, and each flip moves a gradient entry by one
    # whole product term (~1e-5); everywhere else the float64 mask must agree
    mask = np.asarray(net.L[1]._y) > 0
    flips = mask != (z1 > 0)
    assert flips.sum() < 100 and (np.abs(z1[flips]) < 1e-4).all(), (flips.sum(), np.abs(z1[flips]).max(initial=0))
    h = np.where(mask, z1, 0.0)
    z2 = h @ W2.T + b2
    e = np.exp(z2 - z2.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    L = (-(y * np.log(p + 1e-8)).sum(axis=1)).mean()
    dp = -y / (p + 1e-8) / B
    dz2 = dp * p * (1 - p)
    gW2, gb2 = dz2.T @ h, dz2.sum(axis=0)
    dz1 = (dz2 @ W2) * mask
    gW1, gb1 = dz1.T @ x64, dz1.sum(axis=0)
    np.testing.assert_allclose(np.asarray(pred), p, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(np.asarray(loss).reshape(-1)[0], L, rtol=2e-5)
    # gradients: fp32 accumulation over 8192 rows, 3xTF32 products (6e-6 sum|a||b| per GEMM)
    for got, want, name in zip([q.grad for q in params(net)], [gW1, gb1, gW2, gb2], ["dW1", "db1", "dW2", "db2"]):
        g = np.asarray(got).astype(np.float64)
        scale = np.abs(want).max()
        assert np.abs(g - want).max() <= 3e-4 * scale, (name, np.abs(g - want).max(), scale)
    # Adam (t = 1) in float64 on the DEVICE's own gradient: isolates the optimizer from the GEMM error
    for q, v0, name in zip(params(net), [W1, b1, W2, b2], ["W1", "b1", "W2", "b2"]):
        g = np.asarray(q.grad).astype(np.float64)
        m, v = (1 - B1) * g, (1 - B2) * g * g
        diff = -LR * (m / (1 - B1)) / (np.sqrt(v / (1 - B2)) + EPS)
        np.testing.assert_allclose(np.asarray(q.opt_state.m), m, rtol=2e-6, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(np.asarray(q.opt_state.v), v, rtol=2e-6, atol=1e-18, err_msg=name)
        np.testing.assert_allclose(np.asarray(q.value), v0 + diff, rtol=0, atol=2e-7 + 2e-7 * np.abs(v0).max(), err_msg=name)


def test_config5_step_fused_equals_reference_op_order(gpu, batch, monkeypatch):
    """The fused path (softmax kernel, activation backward, GEMM-epilogue accumulate, one-launch Adam) against the
    reference's op-by-op compositions on the same inputs: identical up to the softmax's summation order."""
    from vulkpy_b200.nn import optimizers as O
    x, y = batch
    out = {}
    for unfused in (True, False):
        monkeypatch.setattr(O, "UNFUSED", unfused)
        net = make_net(gpu)
        launches0 = gpu.gpu.launch_count()
        _, loss = net.train(vk.Array(gpu, data=x), vk.Array(gpu, data=y))
        gpu.wait()
        out[unfused] = ([np.asarray(q.grad).copy() for q in params(net)] + [np.asarray(q.opt_state.v).copy() for q in params(net)],
                        float(np.asarray(loss).reshape(-1)[0]), gpu.gpu.launch_count() - launches0)
    (ga, la, na), (gb, lb, nb) = out[True], out[False]
    assert abs(la - lb) <= 1e-6 * abs(la)
    for u, f in zip(ga, gb):
        np.testing.assert_allclose(f, u, rtol=1e-4, atol=1e-6 * np.abs(u).max())
    assert nb < na, (nb, na)          # the fused step launches fewer kernels than the op-by-op chain
