"""Lazy element-wise fusion on the device: `vkp_ew_chain` evaluates a recorded chain of
reference shaders in ONE launch, bit-identical to the op-by-op sequence (one rounding per reference op), and a
chain moves its 8 B per element once (reference chains: nn/layers.py:239-243, nn/losses.py:294-296,373-377)."""
import numpy as np
import pytest

import vulkpy_b200 as vk
from vulkpy_b200 import nn
from vulkpy_b200._backend import Timer

pytestmark = pytest.mark.gpu
F = np.float32
UNARY = ("abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh",
         "exp", "log", "exp2", "log2", "sqrt", "invsqrt")


def bits(a):
    return np.asarray(a).view(np.uint32)


@pytest.mark.parametrize("n", [1, 3, 4, 255, 1024, 2051, 1 << 16, (1 << 20) + 7])
def test_sigmoid_chain_is_bitwise_the_four_op_sequence(gpu, rs, n):
    x_h = rs.uniform(-12, 12, n).astype(F)
    x_h[: min(n, 3)] = [0.0, -120.0, 120.0][: min(n, 3)]          # exp underflow / overflow take the careful path
    x = vk.Array(gpu, data=x_h)
    y = 0.0 - x
    y.exp(inplace=True)
    y += 1.0
    eager = np.asarray(1.0 / y).copy()
    l0 = gpu.gpu.launch_count()
    with vk.fuse():
        y = 0.0 - x
        y.exp(inplace=True)
        y += 1.0
        z = 1.0 / y
    z.wait()
    assert gpu.gpu.launch_count() - l0 == 1
    np.testing.assert_array_equal(bits(z), bits(eager))


def test_every_operation_in_a_chain_matches_its_stand_alone_kernel(gpu, rs):
    n = 70001
    a_h = rs.uniform(0.6, 0.95, n).astype(F)          # inside every unary function's domain but acosh
    b_h = rs.uniform(-2, 2, n).astype(F)
    a, b = vk.Array(gpu, data=a_h), vk.Array(gpu, data=b_h)
    for name in UNARY:
        src = (a + 1.0) if name == "acosh" else a
        eager = np.asarray(getattr(src * 1.0, name)()).copy()
        with vk.fuse():
            lazy = getattr(src * 1.0, name)()
        np.testing.assert_array_equal(bits(lazy), bits(eager), err_msg=name)
    two = {"add": lambda p, q: p + q, "sub": lambda p, q: p - q, "mul": lambda p, q: p * q, "div": lambda p, q: p / q,
           "max": lambda p, q: p.max(q), "min": lambda p, q: p.min(q), "pow": lambda p, q: p ** q}
    for name, f in two.items():
        for other in (b, 1.7, 2.0, -0.5):
            eager = np.asarray(f(a * 1.0, other)).copy()
            with vk.fuse():
                lazy = f(a * 1.0, other)
            np.testing.assert_array_equal(bits(lazy), bits(eager), err_msg=f"{name} {other!r}")
    for name, f in {"rsub": lambda p: 1.7 - p, "rdiv": lambda p: 1.7 / p, "rpow": lambda p: 1.7 ** p, "radd": lambda p: 1.7 + p}.items():
        eager = np.asarray(f(b * 1.0)).copy()
        with vk.fuse():
            lazy = f(b * 1.0)
        np.testing.assert_array_equal(bits(lazy), bits(eager), err_msg=name)
    # in place, several inputs, a long mixed chain; negative bases / zeros reach pow's careful path
    c_h = rs.uniform(-1, 1, n).astype(F)
    c_h[:4] = [0.0, -0.0, -1.0, 1.0]

    def expr(p, q, r):
        t = p * q
        t += r
        t.abs(inplace=True)
        t **= 0.5
        t = (t - q).max(r).min(2.5) / (p + 3.0)
        t -= 0.125
        return (r ** 2.0) + t.tanh()
    eager = np.asarray(expr(a, b, vk.Array(gpu, data=c_h))).copy()
    l0 = gpu.gpu.launch_count()
    with vk.fuse():
        lazy = expr(a, b, vk.Array(gpu, data=c_h))
    lazy.wait()
    assert gpu.gpu.launch_count() - l0 <= 3
    np.testing.assert_array_equal(bits(lazy), bits(eager))


def test_chain_entry_point_with_saved_copy(gpu, rs):
    """vkp_ew_chain directly: Huber's 0.5 * min(|d|, d^2) (nn/losses.py:373-377) with the SAVE step."""
    n = 4099
    x_h, y_h = rs.normal(size=n).astype(F), rs.normal(size=n).astype(F)
    x, y = vk.Array(gpu, data=x_h), vk.Array(gpu, data=y_h)
    out = vk.Array(gpu, shape=(n,))
    SUB, MUL, MIN, POW, ABS, SAVE = 1, 2, 5, 6, 11 + 0, 31
    # acc = y - x; |acc|; tmp = acc; acc = acc ** 2; acc = min(acc, tmp); acc *= 0.5
    out.job = gpu.gpu.ew_chain([y.buffer, x.buffer], out.buffer, [SUB, ABS, SAVE, POW, MIN, MUL], [1, 0, 0, 0, 4, 0],
                               [0, 0, 0, 2.0, 0, 0.5])
    d = y - x
    d.abs(inplace=True)
    d.min(d ** 2.0, inplace=True)
    d *= 0.5
    np.testing.assert_array_equal(bits(out), bits(d))
    with pytest.raises(RuntimeError):
        gpu.gpu.ew_chain([y.buffer], out.buffer, [0], [2], [0.0])          # step reads an input that is not there


def test_nn_chains_fused_equal_unfused(gpu, rs, monkeypatch):
    from vulkpy_b200.nn import optimizers as O
    x_h = rs.normal(size=(130, 37)).astype(F) * 4
    y_h = rs.normal(size=(130, 37)).astype(F)
    res = {}
    for unfused in (True, False):
        monkeypatch.setattr(O, "UNFUSED", unfused)
        sig = nn.Sigmoid()
        l0 = gpu.gpu.launch_count()
        s = sig(vk.Array(gpu, data=x_h))
        s.wait()
        n_sig = gpu.gpu.launch_count() - l0
        mse, hub = nn.MSELoss(), nn.HuberLoss()
        lm = mse(vk.Array(gpu, data=x_h), vk.Array(gpu, data=y_h))
        lh = hub(vk.Array(gpu, data=x_h), vk.Array(gpu, data=y_h))
        ada = nn.AdaGrad(gpu, lr=0.1, tau=0.01).init_state((130, 37))
        d1 = np.asarray(ada.grad2diff(vk.Array(gpu, data=y_h))).copy()
        d2 = np.asarray(ada.grad2diff(vk.Array(gpu, data=x_h))).copy()
        res[unfused] = [np.asarray(s).copy(), np.asarray(lm).copy(), np.asarray(mse.grad()).copy(), np.asarray(lh).copy(),
                        np.asarray(hub.grad()).copy(), d1, d2, np.asarray(ada.h).copy()], n_sig
    (a, na), (b, nb) = res[True], res[False]
    # unfused: 4 ops; exp over 130*37 = 4810 elements is one whole-tile launch + one tail launch (ew_tab_kernel<FULL>)
    assert (na, nb) == (5, 1)
    for u, f in zip(a, b):
        np.testing.assert_array_equal(bits(f), bits(u))


def test_chain_moves_its_bytes_once(gpu):
    """2^26 elements: the fused sigmoid takes about one 8 B/element pass, the op-by-op form four."""
    n = 1 << 27
    x = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=3).random(shape=(n,))

    def chain():
        y = 0.0 - x
        y.exp(inplace=True)
        y += 1.0
        z = 1.0 / y
        z.buffer                    # enqueue (a recorded chain launches here); no host synchronisation
        return z

    def timed(f, reps=10):
        for _ in range(3):
            f()
        t0, t1 = Timer(gpu.gpu), Timer(gpu.gpu)
        t0.record()
        for _ in range(reps):
            f()
        t1.record()
        return t0.elapsed_ms(t1) / reps

    def fused():
        with vk.fuse():
            return chain()
    t_eager, t_fused = timed(chain), timed(fused)
    gbs = 8 * n / t_fused / 1e6
    assert t_fused < 0.8 * t_eager, (t_eager, t_fused)
    assert gbs > 2000, gbs                      # one HBM pass; exp + IEEE divide make the chain issue-bound, not HBM-bound
