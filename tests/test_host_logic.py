"""Host logic of the Python layer on the CPU (no GPU): shape rules, parameter blocks, result
allocation, job / keep-alive bookkeeping, reductions over axes, broadcasting, gather, the PRNG
plumbing and the vulkpy.nn compositions including their fused variants.  Every kernel is answered
by the CPU oracle through ``tests/fake_device.py``; parity of the CUDA kernels themselves is the job
of the ``-m gpu`` tests.  Cases follow the reference's own tests (test/test_vulkpy.py,
test/test_random.py, test/test_nn.py: shapes, error behaviour, known answers)."""
import numpy as np
import pytest

import fake_device
from oracle import vulkpy_oracle as orc

F = np.float32


@pytest.fixture
def vk():
    import vulkpy_b200 as vk
    return vk


@pytest.fixture
def gpu(monkeypatch, vk):
    fake_device.install(monkeypatch)
    return vk.GPU(0)


def A(vk, gpu, x):
    return vk.Array(gpu, data=x)


# ---- arithmetic, in-place forms, scalars ------------------------------------------------------------
def test_operators_and_job_bookkeeping(vk, gpu):
    a, b = A(vk, gpu, [1, 2, 3]), A(vk, gpu, [3, 3, 3])
    c = a + b
    assert c.job is not None and c._keep == [a, b] and c.shape == (3,)
    c.wait()
    assert c.job is None and c._keep == []
    np.testing.assert_array_equal(c, [4, 5, 6])
    np.testing.assert_array_equal(a - b, [-2, -1, 0])
    np.testing.assert_array_equal(a * b, [3, 6, 9])
    np.testing.assert_allclose(a / b, [1 / 3, 2 / 3, 1], rtol=1e-7)
    np.testing.assert_array_equal(a + 1.5, [2.5, 3.5, 4.5])
    np.testing.assert_array_equal(1.5 - a, [0.5, -0.5, -1.5])
    np.testing.assert_allclose(6.0 / a, [6, 3, 2], rtol=1e-7)
    np.testing.assert_allclose(a ** 2.0, [1, 4, 9], rtol=1e-7)
    np.testing.assert_allclose(2.0 ** a, [2, 4, 8], rtol=1e-7)
    a += b
    a *= 2.0
    np.testing.assert_array_equal(a, [8, 10, 12])
    assert repr(a) == "<Array(shape=(3,))>"
    with pytest.raises(ValueError):
        a + A(vk, gpu, [1, 2])
    with pytest.raises(ValueError):
        vk.Array(gpu)


def test_host_write_then_op_and_fill(vk, gpu):
    a = A(vk, gpu, [1, 2, 3])
    a[1] = 7.0
    np.testing.assert_array_equal(a * 2.0, [2, 14, 6])
    n0 = gpu.gpu.launch_count()
    a[:] = 0.5                      # whole-array scalar assignment is one device fill
    assert gpu.gpu.log[-1][0] == "fill" and gpu.gpu.launch_count() == n0 + 1
    np.testing.assert_array_equal(a, [0.5, 0.5, 0.5])
    z = vk.zeros(gpu, shape=(2, 3))
    assert z.shape == (2, 3) and not np.asarray(z).any()


def test_unary_members_match_oracle(vk, gpu):
    x = np.asarray([0.25, 0.5, 0.75], dtype=F)
    for name in fake_device.UN:
        src = x + F(1) if name == "acosh" else x
        a = A(vk, gpu, src)
        np.testing.assert_array_equal(np.asarray(getattr(a, name)()), orc.unary(name, src))
        getattr(a, name)(inplace=True)
        np.testing.assert_array_equal(np.asarray(a), orc.unary(name, src))
        assert gpu.gpu.log[-1] == ("i" + name, 1)


def test_max_min_clamp(vk, gpu):
    a = A(vk, gpu, [[1, 5, 3], [4, 2, 6]])
    b = A(vk, gpu, [[2, 2, 2], [5, 5, 5]])
    np.testing.assert_array_equal(a.max(b), [[2, 5, 3], [5, 5, 6]])
    np.testing.assert_array_equal(a.min(3.0), [[1, 3, 3], [3, 2, 3]])
    np.testing.assert_array_equal(a.clamp(2.0, 4.0), [[2, 4, 3], [4, 2, 4]])
    lo, hi = A(vk, gpu, [2, 3, 4]), A(vk, gpu, [[3], [5]])
    np.testing.assert_array_equal(a.clamp(lo, 5.0), np.clip(np.asarray(a), [2, 3, 4], 5))
    np.testing.assert_array_equal(a.clamp(0.0, hi), np.clip(np.asarray(a), 0, [[3], [5]]))
    np.testing.assert_array_equal(a.clamp(lo, hi), np.minimum(np.maximum(np.asarray(a), [2, 3, 4]), [[3], [5]]))
    # clamping a row vector against matrix bounds broadcasts the source too
    r = A(vk, gpu, [1, 5, 3])
    np.testing.assert_array_equal(r.clamp(b, 4.0), np.clip([1, 5, 3], np.asarray(b), 4))
    with pytest.raises(ValueError):
        r.clamp(b, 4.0, inplace=True)
    a.clamp(2.0, 4.0, inplace=True)
    np.testing.assert_array_equal(a, [[2, 4, 3], [4, 2, 4]])


# ---- broadcasting (doc/broadcasting.md; test_vulkpy.py:1149-1409) ------------------------------------
@pytest.mark.parametrize("sa,sb", [((2, 3), (3,)), ((2, 1), (1, 3)), ((4, 1, 3), (2, 1)), ((3,), (2, 3)), ((1,), (2, 2, 2)),
                                   ((2, 3, 4, 2), (4, 1))])
def test_broadcast_binary_shapes(vk, gpu, sa, sb):
    rs = np.random.default_rng(0)
    x, y = rs.uniform(1, 2, sa).astype(F), rs.uniform(1, 2, sb).astype(F)
    for op, f in (("__add__", np.add), ("__sub__", np.subtract), ("__mul__", np.multiply), ("__truediv__", np.divide),
                  ("__pow__", None)):
        got = getattr(A(vk, gpu, x), op)(A(vk, gpu, y))
        want = orc.broadcast_binary(op.strip("_").replace("truediv", "div"), x, y)
        assert got.shape == want.shape
        np.testing.assert_array_equal(np.asarray(got), want)
        assert gpu.gpu.log[-1][1] == 4            # A, B, C and the shape binding
    np.testing.assert_array_equal(np.asarray(A(vk, gpu, x).max(A(vk, gpu, y))), np.maximum(x, y))


def test_broadcast_inplace_and_errors(vk, gpu):
    x = np.arange(24, dtype=F).reshape(2, 3, 4)
    a = A(vk, gpu, x)
    a += A(vk, gpu, [1, 2, 3, 4])
    a *= A(vk, gpu, [[1], [2], [3]])
    np.testing.assert_array_equal(a, (x + [1, 2, 3, 4]) * [[1], [2], [3]])
    with pytest.raises(ValueError):
        b = A(vk, gpu, [1, 2, 3, 4])
        b += a                                     # the result shape would differ from the target
    with pytest.raises(ValueError):
        a + A(vk, gpu, [1, 2, 3])
    np.testing.assert_array_equal(A(vk, gpu, [1, 2]).broadcast_to((3, 2)), [[1, 2]] * 3)
    np.testing.assert_array_equal(A(vk, gpu, [[1], [2]]).broadcast_to((2, 2, 3)), np.broadcast_to([[1], [2]], (2, 2, 3)))
    with pytest.raises(ValueError):
        A(vk, gpu, [1, 2, 3]).broadcast_to((2,))


# ---- reductions (test_vulkpy.py:808-1121, 1444-1472) -------------------------------------------------------
@pytest.mark.parametrize("name,f", [("sum", np.sum), ("prod", np.prod), ("maximum", np.max), ("minimum", np.min),
                                    ("mean", np.mean)])
def test_reductions_axes_keepdims_rebroadcast(vk, gpu, name, f):
    x = np.random.default_rng(1).uniform(0.5, 1.5, (2, 3, 4, 2)).astype(F)
    a = A(vk, gpu, x)
    r = getattr(a, name)()
    assert r.shape == (1,)
    np.testing.assert_allclose(r, [f(x.astype(np.float64))], rtol=1e-6)
    assert getattr(a, name)(keepdims=True).shape == (1, 1, 1, 1)
    for axis in (0, 1, -1, (0, 2), [1, 3], (3, 1, 1)):      # duplicates are dropped like the reference does
        np_axis = axis if isinstance(axis, int) else tuple(sorted(set(axis)))
        want = f(x.astype(np.float64), axis=np_axis)
        got = getattr(a, name)(axis=axis)
        assert got.shape == want.shape
        np.testing.assert_allclose(got, want, rtol=1e-6)
        wk = f(x.astype(np.float64), axis=np_axis, keepdims=True)
        gk = getattr(a, name)(axis=axis, keepdims=True)
        assert gk.shape == wk.shape
    rb = getattr(a, name)(axis=2, rebroadcast=True)
    assert rb.shape == x.shape
    np.testing.assert_allclose(rb, np.broadcast_to(f(x.astype(np.float64), axis=2, keepdims=True), x.shape), rtol=1e-6)
    with pytest.raises(ValueError):
        getattr(a, name)(axis=(0, 1), rebroadcast=True)
    with pytest.raises(ValueError):
        getattr(a, name)(axis=4)


def test_reduction_launch_counts(vk, gpu):
    a = A(vk, gpu, np.ones((70, 5), F))
    n0 = gpu.gpu.launch_count()
    a.sum()
    assert gpu.gpu.launch_count() == n0 + 1      # one submission, not the reference's log64(n) host loop
    a.sum(axis=(0, 1))
    assert gpu.gpu.launch_count() == n0 + 3      # one pass per axis, descending (vkarray.py:1194-1222)
    assert [name for name, _ in gpu.gpu.log[-2:]] == ["sum_axis", "sum_axis"]
    a.mean(axis=0)
    assert [name for name, _ in gpu.gpu.log[-2:]] == ["sum_axis", "imul_scalar"]   # Q11: sum, then one scale


# ---- reshape / matmul / gather -----------------------------------------------------------------------
def test_reshape_in_place(vk, gpu):
    a = A(vk, gpu, np.arange(6, dtype=F))
    assert a.reshape((2, 3)) is None and a.shape == (2, 3)
    a.reshape((3, -1))
    assert a.shape == (3, 2)
    np.testing.assert_array_equal(a, np.arange(6).reshape(3, 2))
    with pytest.raises(ValueError):
        a.reshape((4, 2))


def test_matmul_shapes(vk, gpu):
    m = np.arange(6, dtype=F).reshape(2, 3)
    v3, v2 = np.asarray([1, 2, 3], F), np.asarray([1, 2], F)
    np.testing.assert_array_equal(A(vk, gpu, m) @ A(vk, gpu, m.T.copy()), m @ m.T)
    r = A(vk, gpu, m) @ A(vk, gpu, v3)
    assert r.shape == (2,)
    np.testing.assert_array_equal(r, m @ v3)
    r = A(vk, gpu, v2) @ A(vk, gpu, m)
    assert r.shape == (3,)
    np.testing.assert_array_equal(r, v2 @ m)
    r = A(vk, gpu, v3) @ A(vk, gpu, v3)
    assert r.shape == (1,)                        # () -> (1,) as in the reference (vkarray.py:585-605)
    np.testing.assert_array_equal(r, [14])
    with pytest.raises(ValueError):
        A(vk, gpu, m) @ A(vk, gpu, m)


def test_gather_onehot_argmax(vk, gpu):
    x = np.arange(24, dtype=F).reshape(2, 3, 4)
    a = A(vk, gpu, x)
    idx = vk.U32Array(gpu, data=[[5, 0], [23, 7]])
    g = a.gather(idx)
    assert g.shape == (2, 2)
    np.testing.assert_array_equal(g, x.reshape(-1)[[[5, 0], [23, 7]]])
    i1 = vk.U32Array(gpu, data=[2, 0])
    g1 = a.gather(i1, axis=1)
    assert g1.shape == (2, 2, 4)                  # index dimensions lead: [idx, prev, post]
    np.testing.assert_array_equal(g1, orc.gather_axis(x, [2, 0], 1))
    oh = vk.U32Array(gpu, data=[0, 2, 1]).to_onehot(3)
    np.testing.assert_array_equal(oh, np.eye(3)[[0, 2, 1]])
    np.testing.assert_array_equal(a.argmax(axis=2), np.argmax(x, axis=2))
    np.testing.assert_array_equal(a.argmin(), [0])
    with pytest.raises(ValueError):
        vk.U32Array(gpu)


# ---- random (random.py:12-24; test_random.py) ------------------------------------------------------------
def test_random_plumbing(vk, gpu):
    r = vk.random.Xoshiro128pp(gpu, seed=0)
    np.testing.assert_allclose(r.random(shape=(3,)), [0.42977667, 0.8235899, 0.90622926], rtol=1e-7)
    np.testing.assert_allclose(r.normal(shape=(3,)), [-2.3403292, 0.7247794, 0.7118352], rtol=2e-6)
    # odd n consumes n + 1 uniforms (random.py:105-124): the next draw continues after them
    o = orc.Xoshiro128pp(64, 0)
    o.random(3), o.random(4)
    np.testing.assert_array_equal(np.asarray(r.randint(shape=(5,))), o.randint(5))
    buf = vk.Array(gpu, shape=(2, 5))
    out = r.random(buffer=buf)
    assert out is buf and buf.shape == (2, 5)
    assert ((np.asarray(buf) >= 0) & (np.asarray(buf) < 1)).all()
    rr = r.randrange(shape=(50,), low=3, high=9)
    assert ((np.asarray(rr) >= 3) & (np.asarray(rr) < 9)).all()
    np.testing.assert_array_equal(np.asarray(r.randrange(shape=(4,), low=3, high=4)), [3] * 4)
    for bad in (dict(low=5, high=5), dict(low=-1, high=3), dict(low=0, high=(1 << 32) + 1)):
        with pytest.raises(ValueError):
            r.randrange(shape=(3,), **bad)
    with pytest.raises(ValueError):
        r.random()
    a, b = vk.random.Xoshiro128pp(gpu, seed=5), vk.random.Xoshiro128pp(gpu, seed=5)
    n1, n2 = a.normal(shape=(10,)), b.normal(shape=(10,), mean=5.0, stddev=3.0)
    np.testing.assert_allclose((np.asarray(n2) - 5) / np.asarray(n1), 3.0, rtol=1e-5)
    p = np.asarray(a.permutation(100))
    assert sorted(p.tolist()) == list(range(100))


# ---- vulkpy.nn (test_nn.py) ----------------------------------------------------------------------------------
def test_dense_known_answers_and_step_paths(vk, gpu, monkeypatch):
    from vulkpy_b200 import nn
    import vulkpy_b200.nn.optimizers as O
    d = nn.Dense(gpu, 2, 2, w_init=nn.Constant(0.0), b_init=nn.Constant(1.5))
    np.testing.assert_array_equal(d(A(vk, gpu, [[1, 2], [3, 4]])), [[1.5, 1.5], [1.5, 1.5]])
    with pytest.raises(ValueError):
        d(A(vk, gpu, [1, 2]))
    rs = np.random.default_rng(3)
    x = rs.normal(size=(6, 4)).astype(F)
    dy = rs.normal(size=(6, 3)).astype(F)
    res = {}
    for unfused in (True, False):
        monkeypatch.setattr(O, "UNFUSED", unfused)
        layer = nn.Dense(gpu, 4, 3, w_init=nn.HeNormal(gpu, 4, seed=1))
        layer(A(vk, gpu, x))
        dx = [np.asarray(layer.backward(A(vk, gpu, dy))).copy() for _ in range(2)]
        res[unfused] = dx + [np.asarray(layer.w.grad).copy(), np.asarray(layer.b.grad).copy()]
    for u, f in zip(res[True], res[False]):
        np.testing.assert_allclose(u, f, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(res[False][2], 2 * dy.T.astype(np.float64) @ x, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(res[False][3], 2 * dy.sum(axis=0), rtol=1e-5, atol=1e-5)


def _mlp(vk, gpu, nn, opt):
    return nn.Sequence([nn.Dense(gpu, 5, 7, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 5, seed=3)), nn.ReLU(),
                        nn.Dense(gpu, 7, 3, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 7, seed=4)), nn.Softmax()],
                       nn.CrossEntropyLoss())


@pytest.mark.parametrize("optname", ["adam", "sgd", "adagrad"])
def test_sequence_train_one_launch_paths(vk, gpu, optname):
    """Sequence.train's one-launch zero_grad / optimizer step and the skipped input gradient of the
    first layer leave the same state as driving the layers one by one (nn/models.py:37-78)."""
    from vulkpy_b200 import nn
    opt = {"adam": lambda: nn.Adam(gpu, lr=1e-2), "sgd": lambda: nn.SGD(0.1), "adagrad": lambda: nn.AdaGrad(gpu, lr=0.1)}[optname]
    rs = np.random.default_rng(7)
    x = rs.normal(size=(9, 5)).astype(F)
    y = np.eye(3, dtype=F)[rs.integers(0, 3, 9)]
    a, b = _mlp(vk, gpu, nn, opt), _mlp(vk, gpu, nn, opt)
    for _ in range(3):
        _, la = a.train(A(vk, gpu, x), A(vk, gpu, y))
        pred = b._forward(A(vk, gpu, x))
        lb = b.loss(pred, A(vk, gpu, y))
        for layer in b.L:
            layer.zero_grad()
        dx = b.loss.grad()
        for layer in reversed(b.L):
            dx = layer.backward(dx)
        for layer in b.L:
            layer.update()
        np.testing.assert_array_equal(np.asarray(la), np.asarray(lb))
    for la_, lb_ in ((a.L[0], b.L[0]), (a.L[2], b.L[2])):
        for pa, pb in ((la_.w, lb_.w), (la_.b, lb_.b)):
            np.testing.assert_array_equal(np.asarray(pa.value), np.asarray(pb.value))
            np.testing.assert_array_equal(np.asarray(pa.grad), np.asarray(pb.grad))
    names = [n for n, _ in gpu.gpu.log]
    assert "fill_many" not in names                  # zero_grad launches nothing: the first contribution overwrites
    assert "nn_softmax_ce_train" in names            # Softmax + CrossEntropy forward / backward tail: one launch
    if optname == "adam":
        assert "nn_adam_apply_many" in names         # all parameters stepped by one launch
    loss0 = float(np.asarray(a.train(A(vk, gpu, x), A(vk, gpu, y))[1]).reshape(-1)[0])
    for _ in range(30):
        loss = float(np.asarray(a.train(A(vk, gpu, x), A(vk, gpu, y))[1]).reshape(-1)[0])
    assert loss < loss0                              # it learns


def test_losses_known_answers(vk, gpu):
    from vulkpy_b200.nn import losses as L
    x, y = A(vk, gpu, [[0.5, 0.5]]), A(vk, gpu, [[1.0, 0.0]])
    ce = L.CrossEntropyLoss()
    np.testing.assert_allclose(ce(x, y), 0.6931472, rtol=1e-6)            # test_nn.py:216-223
    np.testing.assert_allclose(ce.grad(), [[-2.0, 0.0]], rtol=1e-6)
    mse = L.MSELoss(reduce="sum")
    np.testing.assert_allclose(mse(A(vk, gpu, [[1.0, -2.0]]), A(vk, gpu, [[0.0, 0.0]])), 5.0, rtol=1e-6)
    np.testing.assert_allclose(mse.grad(), [[2.0, -4.0]], rtol=1e-6)
    hub = L.HuberLoss(reduce="sum")
    # the reference's Huber is 0.5 * min(|d|^2, |d|) (nn/losses.py:355-359), not the textbook |d| - 0.5 branch
    np.testing.assert_allclose(hub(A(vk, gpu, [[0.5, 3.0]]), A(vk, gpu, [[0.0, 0.0]])), 0.125 + 1.5, rtol=1e-6)
    np.testing.assert_allclose(hub.grad(), [[0.5, 1.0]], rtol=1e-6)
    sce = L.SoftmaxCrossEntropyLoss()
    v = sce(A(vk, gpu, [[0.0, 0.0]]), y)
    np.testing.assert_allclose(v, 0.6931472, rtol=1e-6)
    np.testing.assert_allclose(sce.grad(), [[-0.5, 0.5]], rtol=1e-5, atol=1e-7)
    with pytest.raises(KeyError):                    # a dict lookup in the reference (nn/losses.py:41-46)
        L.MSELoss(reduce="median")


def test_regularizers_and_parameter(vk, gpu):
    from vulkpy_b200 import nn
    from vulkpy_b200.nn.parameters import Parameter
    w = A(vk, gpu, [[-1.0, 2.0], [0.5, -0.5]])
    np.testing.assert_allclose(nn.Lasso(0.5).loss(w), 2.0, rtol=1e-6)
    np.testing.assert_allclose(nn.Lasso(0.5).grad(w), [[-0.5, 0.5], [0.5, -0.5]], rtol=1e-6)
    np.testing.assert_allclose(nn.Ridge(2.0).loss(w), 2.0 * 5.5, rtol=1e-6)
    np.testing.assert_allclose(nn.Ridge(2.0).grad(w), 4.0 * np.asarray(w), rtol=1e-6)
    np.testing.assert_allclose(nn.Elastic(0.5, 2.0).loss(w), 2.0 + 11.0, rtol=1e-6)
    p = Parameter(gpu, shape=(2, 2), opt=nn.SGD(0.5), initializer=nn.Constant(1.0))
    assert p.is_trainable()
    p.add_grad(w)
    p.add_grad(w)
    p.update()
    np.testing.assert_allclose(p.value, 1.0 - 0.5 * 2 * np.asarray(w), rtol=1e-6)
    p.zero_grad()
    assert not np.asarray(p.grad).any()
    frozen = Parameter(gpu, shape=(2,), trainable=False)
    assert not frozen.is_trainable() and frozen.grad is None
    frozen.update()                                  # no-op


# ---- property tests (hypothesis): random shapes through the broadcasting / reduction / gather rules ---------
from hypothesis import given, settings, HealthCheck, strategies as st   # noqa: E402
from hypothesis.extra import numpy as hnp   # noqa: E402

_shapes2 = hnp.mutually_broadcastable_shapes(num_shapes=2, min_dims=0, max_dims=4, min_side=1, max_side=4)
_prop = settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


@_prop
@given(shapes=_shapes2, op=st.sampled_from(["add", "sub", "mul", "div", "max", "min"]), seed=st.integers(0, 1000))
def test_property_broadcast_matches_numpy(vk, gpu, shapes, op, seed):
    (sa, sb), out_shape = shapes.input_shapes, shapes.result_shape
    if not sa or not sb:          # 0-d arrays do not exist in the reference (data is always ravelled to >= 1-d)
        return
    rs = np.random.default_rng(seed)
    x, y = rs.uniform(1, 2, sa).astype(F), rs.uniform(1, 2, sb).astype(F)
    f = {"add": np.add, "sub": np.subtract, "mul": np.multiply, "div": np.divide, "max": np.maximum, "min": np.minimum}[op]
    a, b = A(vk, gpu, x), A(vk, gpu, y)
    got = {"add": lambda: a + b, "sub": lambda: a - b, "mul": lambda: a * b, "div": lambda: a / b,
           "max": lambda: a.max(b), "min": lambda: a.min(b)}[op]()
    assert tuple(got.shape) == tuple(out_shape)
    np.testing.assert_array_equal(np.asarray(got), f(x, y).astype(F))
    if tuple(sa) == tuple(out_shape) and op in ("add", "sub", "mul", "div"):     # in-place form keeps the target's shape
        {"add": a.__iadd__, "sub": a.__isub__, "mul": a.__imul__, "div": a.__itruediv__}[op](b)
        np.testing.assert_array_equal(np.asarray(a), f(x, y).astype(F))
    np.testing.assert_array_equal(np.asarray(b.broadcast_to(out_shape)), np.broadcast_to(y, out_shape))


@_prop
@given(shape=hnp.array_shapes(min_dims=1, max_dims=4, min_side=1, max_side=5), data=st.data())
def test_property_reductions_match_numpy(vk, gpu, shape, data):
    nd = len(shape)
    axes = data.draw(st.lists(st.integers(-nd, nd - 1), min_size=1, max_size=nd))
    keep = data.draw(st.booleans())
    name, f = data.draw(st.sampled_from([("sum", np.sum), ("maximum", np.max), ("minimum", np.min), ("mean", np.mean), ("prod", np.prod)]))
    x = np.random.default_rng(0).uniform(0.5, 1.5, shape).astype(F)
    np_axes = tuple(sorted({a % nd for a in axes}))
    want = f(x.astype(np.float64), axis=np_axes, keepdims=keep)
    got = getattr(A(vk, gpu, x), name)(axis=axes, keepdims=keep)
    assert tuple(got.shape) == tuple(want.shape)
    np.testing.assert_allclose(np.asarray(got), want, rtol=2e-6)
    ax = axes[0]
    rb = getattr(A(vk, gpu, x), name)(axis=ax, rebroadcast=True)
    np.testing.assert_allclose(np.asarray(rb), np.broadcast_to(f(x.astype(np.float64), axis=ax, keepdims=True), shape), rtol=2e-6)
    am = A(vk, gpu, x).argmax(axis=ax)
    np.testing.assert_array_equal(np.asarray(am).reshape(np.argmax(x, axis=ax).shape), np.argmax(x, axis=ax))


@_prop
@given(shape=hnp.array_shapes(min_dims=1, max_dims=4, min_side=1, max_side=4), data=st.data())
def test_property_gather_matches_numpy(vk, gpu, shape, data):
    nd = len(shape)
    axis = data.draw(st.integers(0, nd - 1))
    idx_shape = data.draw(hnp.array_shapes(min_dims=1, max_dims=2, min_side=1, max_side=3))
    x = np.arange(int(np.prod(shape)), dtype=F).reshape(shape)
    idx = np.random.default_rng(1).integers(0, shape[axis], idx_shape).astype(np.uint32)
    got = A(vk, gpu, x).gather(vk.U32Array(gpu, data=idx), axis=axis)
    want = np.moveaxis(np.take(x, idx.astype(np.int64), axis=axis), list(range(axis, axis + idx.ndim)), list(range(idx.ndim)))
    assert tuple(got.shape) == tuple(want.shape)       # index dimensions lead, then prev, then post
    np.testing.assert_array_equal(np.asarray(got), want)
    flat = np.random.default_rng(2).integers(0, x.size, idx_shape).astype(np.uint32)
    np.testing.assert_array_equal(np.asarray(A(vk, gpu, x).gather(vk.U32Array(gpu, data=flat))), x.reshape(-1)[flat])


# ---- sharded generator: every rank draws its rows of the single-GPU stream (dist.Group.random) ----------------
@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("kind", ["random", "randint", "normal"])
def test_sharded_generator_equals_single_stream(vk, gpu, world, kind):
    from vulkpy_b200 import dist
    shape, size, seed = (12, 8), 4, 3
    single = vk.random.Xoshiro128pp(gpu, size=size, seed=seed)
    want = np.asarray(getattr(single, kind)(shape=shape)).copy()
    parts, states = [], []
    for rank in range(world):
        g = dist.Group(None, rank, world, gpu=gpu)
        r = vk.random.Xoshiro128pp(gpu, size=size, seed=seed)      # identical seed on every rank
        sh = g.random(r, shape, kind)
        lo, hi = g.bounds(shape[0])
        assert tuple(sh.shape) == shape and tuple(sh.local.shape) == (hi - lo, shape[1])
        parts.append(np.asarray(sh.local).copy())
        states.append(r.rng.state())
    np.testing.assert_array_equal(np.concatenate(parts, axis=0), want)   # bit-identical to one GPU
    for s in states:                                                      # and every generator ends in the same state
        np.testing.assert_array_equal(s, single.rng.state())
    with pytest.raises(ValueError):
        dist.Group(None, 0, world, gpu=gpu).random(vk.random.Xoshiro128pp(gpu, size=64, seed=1), (world, 5), "random")


# ---- DataParallel over real nn.Sequence replicas (threads as ranks, in-process sum as the exchange) -----------
class _ThreadTransport:
    """All-reduce among threads of one process: the bucket entry point of the NCCL transport."""

    def __init__(self, shared, rank, world):
        self.shared, self.rank, self.world = shared, rank, world

    def allreduce_many(self, arrs, op, scale=1.0):
        assert op == "sum"
        sh = self.shared
        sh["slots"][self.rank] = [np.asarray(a).astype(F).copy() for a in arrs]
        sh["barrier"].wait()
        total = [sum(sh["slots"][r][i] for r in range(self.world)).astype(F) for i in range(len(arrs))]
        sh["barrier"].wait()
        for a, t in zip(arrs, total):
            a[...] = (t * F(scale)).astype(F)
        return arrs


@pytest.mark.parametrize("world", [2, 4])
def test_data_parallel_equals_full_batch_step(vk, gpu, world):
    """Gradient bucket all-reduce + 1/world scaling reproduces the single-process full-batch Adam
    steps (reduce="mean" losses: nn/losses.py:41-46) with the real layers and their fused paths."""
    import threading
    from vulkpy_b200 import dist, nn
    rs = np.random.default_rng(11)
    B = 8 * world
    x = rs.normal(size=(B, 5)).astype(F)
    y = np.eye(3, dtype=F)[rs.integers(0, 3, B)]
    opt = lambda: nn.Adam(gpu, lr=1e-2)
    single = _mlp(vk, gpu, nn, opt)
    for _ in range(3):
        _, want_loss = single.train(A(vk, gpu, x), A(vk, gpu, y))
    shared = {"slots": [None] * world, "barrier": threading.Barrier(world)}
    nets, losses, errors = [None] * world, [None] * world, []

    def rank_main(rank):
        try:
            g = dist.Group(_ThreadTransport(shared, rank, world), rank, world, gpu=gpu)
            nets[rank] = _mlp(vk, gpu, nn, opt)
            dp = dist.DataParallel(nets[rank], g)
            lo, hi = g.bounds(B)
            for _ in range(3):
                _, losses[rank] = dp.train(A(vk, gpu, x[lo:hi]), A(vk, gpu, y[lo:hi]))
        except Exception as e:   # pragma: no cover
            errors.append(e)
            shared["barrier"].abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    for net in nets:
        for ls, ld in ((single.L[0], net.L[0]), (single.L[2], net.L[2])):
            np.testing.assert_allclose(np.asarray(ld.w.value), np.asarray(ls.w.value), rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(np.asarray(ld.b.value), np.asarray(ls.b.value), rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(float(np.asarray(losses[0]).reshape(-1)[0]), float(np.asarray(want_loss).reshape(-1)[0]), rtol=1e-5)


# ---- lazy element-wise fusion (vk.fuse()): recording, hazards, launch counts ------------------------------------
def test_fuse_records_chains_and_matches_eager(vk, gpu):
    rs = np.random.default_rng(3)
    x_h = rs.uniform(-2, 2, (7, 9)).astype(F)
    b_h = rs.uniform(0.5, 2, (7, 9)).astype(F)

    def sigmoid_chain(x):                     # nn/layers.py:239-243 of the reference
        y = 0.0 - x
        y.exp(inplace=True)
        y += 1.0
        return 1.0 / y

    x, b = vk.Array(gpu, data=x_h), vk.Array(gpu, data=b_h)
    eager = np.asarray(sigmoid_chain(x)).copy()
    n0 = gpu.gpu.launch_count()
    with vk.fuse():
        y = sigmoid_chain(x)
        assert gpu.gpu.launch_count() == n0                    # nothing launched yet
        z = (y * b - 0.25).max(b) ** 2.0                       # continues the same chain, second input
    assert gpu.gpu.launch_count() == n0
    l0 = len(gpu.gpu.log)
    got_y, got_z = np.asarray(y).copy(), np.asarray(z).copy()
    assert gpu.gpu.log[l0:] == [("ew_chain", 2), ("ew_chain", 3)]      # 4 ops -> 1 launch; 8 ops, 2 inputs -> 1 launch
    np.testing.assert_array_equal(got_y, eager)
    want = np.asarray(((vk.Array(gpu, data=eager) * b - 0.25).max(b)) ** 2.0)
    np.testing.assert_array_equal(got_z, want)
    assert y.shape == (7, 9) and z.shape == (7, 9)


def test_fuse_hazards(vk, gpu):
    a_h = np.arange(12, dtype=F).reshape(3, 4) + 1
    a = vk.Array(gpu, data=a_h)
    with vk.fuse():
        c = a + 1.0                 # recorded: reads a
        a *= 2.0                    # overwrites a: c must be evaluated from the OLD a first
        d = a + 0.5                 # reads the NEW a (extends a's in-place chain)
        a -= 1.0                    # d must see 2a, not 2a - 1
        e = a.sqrt()
    np.testing.assert_array_equal(np.asarray(c), a_h + 1)
    np.testing.assert_array_equal(np.asarray(d), a_h * 2 + F(0.5))
    np.testing.assert_array_equal(np.asarray(a), a_h * 2 - 1)
    np.testing.assert_array_equal(np.asarray(e), np.sqrt(a_h * 2 - 1))
    # lazy operand on the right, later modified in place
    p, q = vk.Array(gpu, data=a_h), vk.Array(gpu, data=a_h)
    with vk.fuse():
        r = q * 3.0                 # lazy
        s = p + r                   # chain over p with the lazy r as second input
        r += 100.0                  # s must have been evaluated with the old r
    np.testing.assert_array_equal(np.asarray(s), a_h + a_h * 3)
    np.testing.assert_array_equal(np.asarray(r), a_h * 3 + 100)
    # a non-fusable consumer (reduction, broadcast) forces evaluation; host writes go to the evaluated array
    with vk.fuse():
        t = (p * 2.0).sum(axis=0)
        u = p * 2.0
        u[:] = 7.0
        w = p + vk.Array(gpu, data=a_h[0])          # broadcast: eager path
    np.testing.assert_array_equal(np.asarray(t), (a_h * 2).sum(axis=0))
    np.testing.assert_array_equal(np.asarray(u), np.full_like(a_h, 7))
    np.testing.assert_array_equal(np.asarray(w), a_h + a_h[0])


def test_fuse_long_chains_split_and_dropped_results_cost_nothing(vk, gpu):
    a_h = np.linspace(0.1, 1.0, 64, dtype=F)
    a = vk.Array(gpu, data=a_h)
    n0 = gpu.gpu.launch_count()
    with vk.fuse():
        y = a
        for k in range(40):         # 40 steps: 16 per launch
            y = y + 0.5
        _ = a * 3.0                 # never looked at: never launched
    got = np.asarray(y)
    want = a_h.copy()
    for k in range(40):
        want = want + F(0.5)
    np.testing.assert_array_equal(got, want)
    assert gpu.gpu.launch_count() - n0 == 3
    # five arrays in one expression: at most four inputs per chain
    arrs = [vk.Array(gpu, data=a_h * (k + 1)) for k in range(6)]
    with vk.fuse():
        s = arrs[0]
        for t in arrs[1:]:
            s = s + t
    np.testing.assert_array_equal(np.asarray(s), sum((a_h * (k + 1) for k in range(1, 6)), a_h * 1))


def test_fuse_continues_the_lazy_operand_chain(vk, gpu):
    a_h = np.linspace(0.5, 2.0, 30, dtype=F).reshape(5, 6)
    g_h = np.linspace(-1.0, 1.0, 30, dtype=F).reshape(5, 6)
    a, g = vk.Array(gpu, data=a_h), vk.Array(gpu, data=g_h)
    l0 = len(gpu.gpu.log)
    with vk.fuse():
        root = a.sqrt()
        root += 1e-3
        d = g / root              # concrete / lazy: root's chain continues with a reversed divide
        e = 2.0 - (g - root)      # concrete - lazy, then scalar - lazy
    got_d, got_e = np.asarray(d).copy(), np.asarray(e).copy()
    assert gpu.gpu.log[l0:] == [("ew_chain", 3), ("ew_chain", 3)]
    r = np.sqrt(a_h) + F(1e-3)
    np.testing.assert_array_equal(got_d, g_h / r)
    np.testing.assert_array_equal(got_e, F(2.0) - (g_h - r))
