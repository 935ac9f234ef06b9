"""GPU parity tests of the array hot path (run with -m gpu on the B200 box).

Structure follows the reference's acceptance suite (test/test_vulkpy.py: small known answers
compared with NumPy through np.testing.assert_allclose, same tolerances: rtol 1e-7 for
exp/log/exp2/log2/sqrt/invsqrt/pow, 1e-5 for the trigonometric/hyperbolic group, 1e-3 for
asin/acos) and adds oracle comparisons on random, ragged and empty inputs.  Every call goes
through the public API -> ctypes -> C ABI -> CUDA kernels."""
import numpy as np
import pytest

import vulkpy_b200 as vk
from oracle import vulkpy_oracle as orc

pytestmark = pytest.mark.gpu

F = np.float32


def A(gpu, data):
    return vk.Array(gpu, data=data)


# ------------------------------------------------------------------------- arithmetic
@pytest.mark.parametrize("op,a,b,want", [
    ("add", [5, 5, 5], [1, 1, 1], [6, 6, 6]),
    ("sub", [4, 4, 4], [2, 2, 2], [2, 2, 2]),
    ("mul", [2, 2, 2], [3, 3, 3], [6, 6, 6]),
    ("div", [8, 8, 8], [2, 2, 2], [4, 4, 4]),
])
def test_binary_known_answers(gpu, op, a, b, want):
    x, y = A(gpu, a), A(gpu, b)
    f = {"add": lambda: x + y, "sub": lambda: x - y, "mul": lambda: x * y, "div": lambda: x / y}[op]
    c = f()
    c.wait()
    np.testing.assert_allclose(c, np.asarray(want))
    g = {"add": x.__iadd__, "sub": x.__isub__, "mul": x.__imul__, "div": x.__itruediv__}[op]
    r = g(y)
    assert r is x
    np.testing.assert_allclose(x, np.asarray(want))


def test_scalar_and_reflected(gpu):
    a = A(gpu, [2, 4, 8])
    np.testing.assert_allclose(a + 1, [3, 5, 9])
    np.testing.assert_allclose(a - 1, [1, 3, 7])
    np.testing.assert_allclose(a * 3, [6, 12, 24])
    np.testing.assert_allclose(a / 2, [1, 2, 4])
    np.testing.assert_allclose(1 + a, [3, 5, 9])
    np.testing.assert_allclose(10 - a, [8, 6, 2])
    np.testing.assert_allclose(3 * a, [6, 12, 24])
    np.testing.assert_allclose(8 / a, [4, 2, 1])
    a += 1
    a *= 2
    a -= 2
    a /= 4
    np.testing.assert_allclose(a, [1, 2, 4])


def test_cascade_chain_and_unrelated(gpu):
    a, b, c, d = A(gpu, [5, 5, 5]), A(gpu, [1, 1, 1]), A(gpu, [4, 4, 4]), A(gpu, [2, 2, 2])
    np.testing.assert_allclose(a + b + c, [10, 10, 10])
    e, f = a + b, c + d
    np.testing.assert_allclose(e, f)
    g = a + e
    np.testing.assert_allclose(g, [11, 11, 11])


def test_host_write_paths(gpu):
    a = vk.Array(gpu, shape=(3,))
    a[:] = 10
    np.testing.assert_allclose(a, [10, 10, 10])
    b = vk.Array(gpu, shape=(3,))
    a[:] = 5
    b[:] = 7
    c = a + b
    c.wait()
    np.testing.assert_allclose(c, [12, 12, 12])
    a[1] = 9                      # element write through the host view
    np.testing.assert_allclose(a + b, [12, 16, 12])
    a[:] = np.asarray([1, 2, 3])  # array write through the host view
    np.testing.assert_allclose(a * 2, [2, 4, 6])
    a.array[0] = 100              # the live view (example/02-nn.py uses .array directly)
    np.testing.assert_allclose(a + 0, [100, 2, 3])
    assert a[2] == 3


def test_view_keeps_buffer_alive(gpu):
    v = np.asarray(A(gpu, [1, 2, 3]) + A(gpu, [1, 1, 1]))   # temporaries die here
    for _ in range(8):
        junk = A(gpu, [9, 9, 9]) * 2
        junk.wait()
    np.testing.assert_allclose(v, [2, 3, 4])


def test_shapes_and_errors(gpu):
    with pytest.raises(ValueError):
        A(gpu, [1, 1]) + A(gpu, [1, 1, 1])
    with pytest.raises(ValueError):
        vk.Array(gpu)
    with pytest.raises(ValueError):
        vk.U32Array(gpu)
    c = A(gpu, [[1, 1], [1, 1]]) + A(gpu, [[2, 2], [2, 2]])
    np.testing.assert_allclose(c, [[3, 3], [3, 3]])
    a = A(gpu, [1, 2, 3, 4])
    assert a.reshape((2, 2)) is None
    np.testing.assert_allclose(a, [[1, 2], [3, 4]])
    with pytest.raises(ValueError):
        A(gpu, [1, 2, 3]).reshape((2, 2))
    assert repr(a) == "<Array(shape=(2, 2))>"
    assert "1." in str(a)
    assert vk.GPU(0) == gpu and (gpu == 3) is False


def test_job_semantics(gpu):
    a, b = A(gpu, [1, 2, 3]), A(gpu, [1, 1, 1])
    c = a + b
    assert c.job is not None and c._keep == [a, b]
    c.job.wait(10_000_000_000)
    c.wait()
    assert c.job is None and c._keep == []
    gpu.wait()
    a.flush()
    gpu.flush([a, b])


# ------------------------------------------------------------------------- matmul
def test_matmul_forms(gpu):
    m = A(gpu, [[1, 2], [3, 4]])
    np.testing.assert_allclose(m @ m, [[7, 10], [15, 22]])
    np.testing.assert_allclose(m @ A(gpu, [1, 3]), [7, 15])
    np.testing.assert_allclose(A(gpu, [1, 2]) @ m, [7, 10])
    r = A(gpu, [1, 2]) @ A(gpu, [3, 4])
    assert r.shape == (1,)
    np.testing.assert_allclose(r, [11])
    with pytest.raises(ValueError):
        A(gpu, [[1, 2, 3], [4, 5, 6]]) @ A(gpu, [[1, 2, 3], [4, 5, 6]])


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (7, 13, 5), (64, 64, 64), (65, 17, 129), (200, 300, 100)])
def test_matmul_vs_oracle(gpu, rs, m, k, n):
    a = rs.uniform(-1, 1, (m, k)).astype(F)
    b = rs.uniform(-1, 1, (k, n)).astype(F)
    got = np.asarray(A(gpu, a) @ A(gpu, b))
    # reference: serial fp32 accumulation (matmul.comp:29-31); ours: fp32 FMA, different order
    np.testing.assert_allclose(got, orc.matmul(a, b), rtol=0, atol=2e-6 * k)


# ------------------------------------------------------------------------- max / min / unary
def test_max_min(gpu):
    a, b = A(gpu, [1, 5, 3]), A(gpu, [4, 2, 3])
    np.testing.assert_allclose(a.max(b), [4, 5, 3])
    np.testing.assert_allclose(a.min(b), [1, 2, 3])
    np.testing.assert_allclose(a.max(2.5), [2.5, 5, 3])
    np.testing.assert_allclose(a.min(2.5), [1, 2.5, 2.5])
    assert a.max(b, inplace=True) is a
    np.testing.assert_allclose(a, [4, 5, 3])
    a.min(4.5, inplace=True)
    np.testing.assert_allclose(a, [4, 4.5, 3])


UNARY_CASES = [
    # name, numpy function, inputs, rtol (the reference's own tolerances)
    ("abs", np.abs, [-1.5, 0, 2], 1e-7),
    ("sign", np.sign, [-3, 0, 2], 1e-7),
    ("sin", np.sin, [0.3, 1.2, -2.0], 1e-5), ("cos", np.cos, [0.3, 1.2, -2.0], 1e-5),
    ("tan", np.tan, [0.3, 1.2, -0.7], 1e-5),
    ("asin", np.arcsin, [0.1, -0.5, 0.9], 1e-3), ("acos", np.arccos, [0.1, -0.5, 0.9], 1e-3),
    ("atan", np.arctan, [0.1, -2.0, 30.0], 1e-5),
    ("sinh", np.sinh, [0.1, -1.5, 3.0], 1e-5), ("cosh", np.cosh, [0.1, -1.5, 3.0], 1e-5),
    ("tanh", np.tanh, [0.1, -1.5, 3.0], 1e-5),
    ("asinh", np.arcsinh, [0.1, -1.5, 300.0], 1e-5), ("acosh", np.arccosh, [1.1, 2.5, 300.0], 1e-5),
    ("atanh", np.arctanh, [0.1, -0.5, 0.9], 1e-5),
    ("exp", np.exp, [1, 2, 3], 1e-7), ("log", np.log, [1, 2, 3], 1e-7),
    ("exp2", np.exp2, [1, 2, 3], 1e-7), ("log2", np.log2, [1, 2, 3], 1e-7),
    ("sqrt", np.sqrt, [1, 2, 3], 1e-7), ("invsqrt", lambda x: 1 / np.sqrt(x), [1, 2, 3], 1e-7),
]


@pytest.mark.parametrize("name,f,x,rtol", UNARY_CASES)
def test_unary(gpu, name, f, x, rtol):
    a = A(gpu, x)
    want = f(np.asarray(x, dtype=np.float64))
    np.testing.assert_allclose(getattr(a, name)(), want, rtol=rtol)
    r = getattr(a, name)(inplace=True)
    assert r is a
    np.testing.assert_allclose(a, want, rtol=rtol)


@pytest.mark.parametrize("name", ["exp", "log", "exp2", "log2", "sqrt", "invsqrt", "sin", "cos", "tanh", "abs", "sign"])
@pytest.mark.parametrize("n", [1, 5, 1023, 4096 * 3 + 2])
def test_unary_vs_oracle_ragged(gpu, rs, name, n):
    lo, hi = (0.01, 50.0) if name in ("log", "log2", "sqrt", "invsqrt") else (-20.0, 20.0)
    x = rs.uniform(lo, hi, n).astype(F)
    got = np.asarray(getattr(A(gpu, x), name)())
    want = orc.unary(name, x)
    if name in ("abs", "sign", "sqrt"):
        np.testing.assert_array_equal(got, want)       # IEEE operations: bit exact
    elif name == "invsqrt":
        # 1 / sqrt(x) as two correctly rounded IEEE operations: <= 1 ulp from the CR value
        np.testing.assert_allclose(got, want, rtol=1.2e-7, atol=0)
        np.testing.assert_array_equal(got, (np.float32(1) / np.sqrt(x)).astype(F))
    elif name in ("exp", "log", "exp2", "log2"):
        # binary64 evaluation rounded once: <= 0.5001 ulp, i.e. at most 1 ulp from the CR value
        np.testing.assert_allclose(got, want, rtol=1.2e-7, atol=0)
        assert (got != want).mean() < 1e-3
    else:
        np.testing.assert_allclose(got, want, rtol=5e-7, atol=1e-7)   # libdevice: <= 2 ulp


def test_pow_family(gpu):
    a = A(gpu, [1, 2, 3])
    e = F([1.1, 2.2, 1.4])
    np.testing.assert_allclose(a ** A(gpu, e), np.power(np.float64([1, 2, 3]), e.astype(np.float64)), rtol=1e-7)
    np.testing.assert_allclose(a ** 2.7, np.power(np.float64([1, 2, 3]), np.float64(F(2.7))), rtol=1e-7)
    np.testing.assert_allclose(1.3 ** A(gpu, e), np.power(np.float64(F(1.3)), e.astype(np.float64)), rtol=1e-7)
    b = A(gpu, [1, 2, 3])
    b **= A(gpu, e)
    np.testing.assert_allclose(b, np.power(np.float64([1, 2, 3]), e.astype(np.float64)), rtol=1e-7)
    c = A(gpu, [1, 2, 3])
    c **= 2.7
    np.testing.assert_allclose(c, np.power(np.float64([1, 2, 3]), np.float64(F(2.7))), rtol=1e-7)
    # negative base with an integral exponent (nn/losses.py:294-296 squares negative values)
    np.testing.assert_allclose(A(gpu, [-3, -0.5, 2]) ** 2.0, [9, 0.25, 4], rtol=1e-7)


def test_pow_vs_oracle(gpu, rs):
    x = rs.uniform(0.5, 2.0, 10007).astype(F)
    y = rs.uniform(-2, 2, 10007).astype(F)
    got = np.asarray(A(gpu, x) ** A(gpu, y))
    np.testing.assert_allclose(got, orc.binary("pow", x, y), rtol=1.2e-7)


@pytest.mark.parametrize("s", [2.7, 0.5, -1.5, 3.0, 7.7, -7.75, 1 / 3, -0.25, 9.5])
def test_pow_scalar_binomial_kernel(gpu, rs, s):
    """a ** s on >= 2^15 elements takes the binomial-series kernel (ew_pows_kernel; |s| > 7.75 stays on the
    general one): correctly rounded against float64 like the oracle, over the whole positive range, with
    zeros / negatives / subnormals / inf / nan sprinkled in (their vectors re-run through pow_f), ragged
    sizes (partial tile + n % 4 tail) and in place (pow_scalar.comp:25, ipow_scalar.comp)."""
    for n in (1 << 15, (1 << 16) + 4099, (1 << 17) + 1):
        x = np.concatenate([rs.uniform(0.5, 2.0, n // 2), np.exp(rs.uniform(-80, 80, n - n // 2))]).astype(F)
        x[rs.integers(0, n, 64)] = rs.choice(F([0.0, -0.0, -1.5, -2.0, 1e-40, np.inf, np.nan, 1.0]), 64)
        want = orc.scalar("pow", x, s)
        got = np.asarray(A(gpu, x) ** s)
        np.testing.assert_allclose(got, want, rtol=1.2e-7, equal_nan=True)
        assert (got != want).mean() < 1e-3                    # both are <= 0.5001 ulp: they differ on near-ties only
        b = A(gpu, x)
        b **= s
        np.testing.assert_array_equal(np.asarray(b), got)


def test_pow_scalar_kernels_agree(gpu, rs):
    """below 2^15 elements the general table kernel runs; both kernels round the same values"""
    x = rs.uniform(0.1, 10.0, 1 << 15).astype(F)
    big = np.asarray(A(gpu, x) ** 2.7)
    small = np.concatenate([np.asarray(A(gpu, x[i:i + 4096]) ** 2.7) for i in range(0, x.size, 4096)])
    assert (big != small).mean() < 1e-3
    np.testing.assert_allclose(big, small, rtol=1.2e-7)


# ------------------------------------------------------------------------- clamp
def test_clamp_family(gpu):
    x = [1, 2, 3]
    lo, hi = A(gpu, [1.5, 1.5, 1.5]), A(gpu, [2.5, 2.5, 2.5])
    for inplace in (False, True):
        np.testing.assert_allclose(A(gpu, x).clamp(lo, hi, inplace=inplace), [1.5, 2, 2.5])
        np.testing.assert_allclose(A(gpu, x).clamp(1.5, hi, inplace=inplace), [1.5, 2, 2.5])
        np.testing.assert_allclose(A(gpu, x).clamp(lo, 2.5, inplace=inplace), [1.5, 2, 2.5])
        np.testing.assert_allclose(A(gpu, x).clamp(1.5, 2.5, inplace=inplace), [1.5, 2, 2.5])
    a = A(gpu, x)
    assert a.clamp(0, 2, inplace=True) is a


def test_clamp_broadcast(gpu):
    a = A(gpu, [[1, 2, 3], [4, 5, 6]])
    np.testing.assert_allclose(a.clamp(A(gpu, [2, 2, 2]), A(gpu, [[4], [5]])), [[2, 2, 3], [4, 5, 5]])
    np.testing.assert_allclose(A(gpu, [1, 2, 3]).clamp(A(gpu, [[0], [2.5]]), 10.0), [[1, 2, 3], [2.5, 2.5, 3]])
    np.testing.assert_allclose(a.clamp(1.5, A(gpu, [3, 4, 5]), inplace=True), [[1.5, 2, 3], [3, 4, 5]])
    with pytest.raises(ValueError):
        A(gpu, [1, 2, 3]).clamp(A(gpu, [[0], [1]]), 5.0, inplace=True)


# ------------------------------------------------------------------------- reductions
RED = [("sum", np.sum), ("prod", np.prod), ("maximum", np.max), ("minimum", np.min)]


@pytest.mark.parametrize("name,f", RED)
def test_reduce_full_and_axes(gpu, name, f):
    x = np.asarray([[1, 2, 3], [4, 5, 6]], dtype=np.float64) / 2
    a = A(gpu, x)
    r = getattr(a, name)()
    assert r.shape == (1,)
    np.testing.assert_allclose(r, [f(x)])
    assert getattr(a, name)(keepdims=True).shape == (1, 1)
    np.testing.assert_allclose(getattr(a, name)(axis=0), f(x, axis=0))
    np.testing.assert_allclose(getattr(a, name)(axis=1), f(x, axis=1))
    np.testing.assert_allclose(getattr(a, name)(axis=-1), f(x, axis=-1))
    k = getattr(a, name)(axis=0, keepdims=True)
    assert k.shape == (1, 3)
    np.testing.assert_allclose(k, f(x, axis=0, keepdims=True))
    np.testing.assert_allclose(getattr(a, name)(axis=(0, 1)), f(x, axis=(0, 1)))
    rb = getattr(a, name)(axis=1, rebroadcast=True)
    assert rb.shape == (2, 3)
    np.testing.assert_allclose(rb, np.broadcast_to(f(x, axis=1, keepdims=True), (2, 3)))
    with pytest.raises(ValueError):
        getattr(a, name)(axis=(0, 1), rebroadcast=True)


def test_reduce_crossing_one_workgroup(gpu):
    # test/test_vulkpy.py:822-827 uses 65 elements
    np.testing.assert_allclose(A(gpu, np.ones(65)).sum(), [65])
    np.testing.assert_allclose(A(gpu, np.arange(65)).maximum(), [64])
    np.testing.assert_allclose(A(gpu, np.arange(65)).minimum(), [0])
    np.testing.assert_allclose(A(gpu, np.full(65, 1.01)).prod(), [1.01 ** 65], rtol=1e-5)


def test_reduce_many_dims(gpu):
    x = np.ones((2, 3, 4, 2, 2, 4, 3))
    a = A(gpu, x)
    np.testing.assert_allclose(a.sum(axis=(1, 3, 5)), x.sum(axis=(1, 3, 5)))
    np.testing.assert_allclose(a.sum(axis=[6, 0]), x.sum(axis=(0, 6)))
    np.testing.assert_allclose(a.sum(), [x.sum()])
    assert a.sum(axis=(0, 2), keepdims=True).shape == (1, 3, 1, 2, 2, 4, 3)


def test_mean(gpu):
    x = np.asarray([[1, 2, 3], [4, 5, 9]], dtype=np.float64)
    a = A(gpu, x)
    np.testing.assert_allclose(a.mean(), [x.mean()], rtol=1e-6)
    np.testing.assert_allclose(a.mean(axis=0), x.mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(a.mean(axis=1, keepdims=True), x.mean(axis=1, keepdims=True), rtol=1e-6)
    np.testing.assert_allclose(a.mean(axis=1, rebroadcast=True),
                               np.broadcast_to(x.mean(axis=1, keepdims=True), x.shape), rtol=1e-6)


@pytest.mark.parametrize("shape,axis", [
    ((1000,), 0), ((37, 129), 0), ((37, 129), 1), ((5, 7, 9), 1), ((3, 2000), 1), ((2000, 3), 0),
    ((300, 5000), 0), ((300, 5000), 1), ((4, 100000), 1), ((16, 64, 32), 1), ((2, 70000, 4), 1), ((70000, 8), 0),
    # TMA-staged column kernel (post >= 128, post % 4 == 0, axis >= 512): ragged last box / last column tile,
    # several prev slabs, split axis with a second pass; and shapes just below its thresholds
    ((515, 128), 0), ((3, 600, 132), 1), ((2, 1000, 260), 1), ((512, 256), 0), ((20000, 1024), 0), ((5, 530, 4, 64), 1),
    ((70, 128), 0), ((511, 256), 0),
])
@pytest.mark.parametrize("name", ["sum", "maximum", "minimum"])
def test_axis_reduce_vs_oracle(gpu, rs, shape, axis, name):
    x = rs.uniform(0, 1, shape).astype(F)
    got = np.asarray(getattr(A(gpu, x), name)(axis=axis))
    want = orc.reduce_axis(name, x, axis)
    if name == "sum":
        # reference: serial fp32 order k=0..axis-1 (sum_axis.comp:27-31); ours: split + tree.
        # both are within n*eps/2 of the exact sum of non-negative terms; compare to the float64 sum
        exact = x.astype(np.float64).sum(axis=axis)
        np.testing.assert_allclose(got, exact, rtol=2e-6)
        np.testing.assert_allclose(got, want, rtol=shape[axis] * 1.2e-7)
    else:
        np.testing.assert_array_equal(got, want)
    rb = np.asarray(getattr(A(gpu, x), name)(axis=axis, rebroadcast=True))
    np.testing.assert_array_equal(rb, np.broadcast_to(np.expand_dims(got, axis), shape))


def test_axis_reduce_tma_path_signs_and_prod(gpu, rs):
    """Rows that TMA zero-fills past the end of the axis (or that belong to the next split) must not
    leak into max of negatives / min of positives / prod."""
    x = rs.uniform(-2.0, -1.0, (3, 547, 260)).astype(F)
    np.testing.assert_array_equal(np.asarray(A(gpu, x).maximum(axis=1)), x.max(axis=1))
    np.testing.assert_array_equal(np.asarray(A(gpu, -x).minimum(axis=1)), (-x).min(axis=1))
    y = rs.uniform(0.995, 1.005, (2, 700, 128)).astype(F)
    np.testing.assert_allclose(np.asarray(A(gpu, y).prod(axis=1)), y.astype(np.float64).prod(axis=1), rtol=5e-5)
    z = rs.uniform(-1, 1, (4000, 384)).astype(F)
    np.testing.assert_allclose(np.asarray(A(gpu, z).sum(axis=0)), z.astype(np.float64).sum(axis=0), rtol=0,
                               atol=4000 * 1.2e-7)
    np.testing.assert_allclose(np.asarray(A(gpu, z).mean(axis=0)), z.astype(np.float64).mean(axis=0), rtol=0, atol=2e-7)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 4096, 4097, 100003, 3_000_001])
def test_full_reduce_sizes(gpu, rs, n):
    x = rs.uniform(0.5, 1.5, n).astype(F)
    a = A(gpu, x)
    np.testing.assert_allclose(a.sum(), [orc.reduce_full_exact("sum", x)], rtol=2e-6)
    np.testing.assert_array_equal(np.asarray(a.maximum()), [x.max()])
    np.testing.assert_array_equal(np.asarray(a.minimum()), [x.min()])
    if n <= 4096:  # where the reference's own multi-pass result is defined
        np.testing.assert_allclose(a.sum(), orc.reduce_full_reference("sum", x), rtol=n * 1.2e-7)
    y = rs.uniform(0.999, 1.001, n).astype(F)
    # n-1 float32 multiplications, each within 2^-24: observed drift ~1e-9 * n (NumPy's own
    # float32 product drifts the same way); bound used here: n * 2^-26
    np.testing.assert_allclose(A(gpu, y).prod(), [orc.reduce_full_exact("prod", y)], rtol=max(1e-5, n * 1.5e-8))


def test_literal_strided_sum_shader(gpu, rs):
    """'sum' with sizeB > 1 keeps sum.comp's literal semantics: b[i] = sum a[i::sizeB]."""
    from vulkpy_b200._backend import MultiVector2Params, DataShape
    x = rs.uniform(0, 1, 1000).astype(F)
    a, b = A(gpu, x), vk.Array(gpu, shape=(16,))
    b.job = gpu._submit("sum", 64, 1, 1, [a, b], DataShape(64, 1, 1), MultiVector2Params(1000, 16))
    np.testing.assert_allclose(b, [x[i::16].astype(np.float64).sum() for i in range(16)], rtol=1e-5)


# ------------------------------------------------------------------------- broadcasting
def test_broadcast_to(gpu):
    a = A(gpu, [1, 2, 3])
    np.testing.assert_allclose(a.broadcast_to((2, 3)), [[1, 2, 3], [1, 2, 3]])
    np.testing.assert_allclose(A(gpu, [[1], [2]]).broadcast_to((2, 3)), [[1, 1, 1], [2, 2, 2]])
    np.testing.assert_allclose(A(gpu, [7]).broadcast_to((2, 2, 2)), np.full((2, 2, 2), 7))
    with pytest.raises(ValueError):
        a.broadcast_to((2, 2))


def test_broadcast_arithmetic(gpu):
    a = A(gpu, [[1, 2, 3], [4, 5, 6]])
    row, col = A(gpu, [10, 20, 30]), A(gpu, [[1], [2]])
    np.testing.assert_allclose(a + row, [[11, 22, 33], [14, 25, 36]])
    np.testing.assert_allclose(a - col, [[0, 1, 2], [2, 3, 4]])
    np.testing.assert_allclose(a * col, [[1, 2, 3], [8, 10, 12]])
    np.testing.assert_allclose(a / row, [[0.1, 0.1, 0.1], [0.4, 0.25, 0.2]], rtol=1e-6)
    np.testing.assert_allclose(row + col, [[11, 21, 31], [12, 22, 32]])     # both operands broadcast
    np.testing.assert_allclose(a.max(row), [[10, 20, 30], [10, 20, 30]])
    np.testing.assert_allclose(a.min(col), [[1, 1, 1], [2, 2, 2]])
    np.testing.assert_allclose(A(gpu, [[1, 2], [3, 4]]) ** A(gpu, [2, 3]), [[1, 8], [9, 64]])
    a += row
    np.testing.assert_allclose(a, [[11, 22, 33], [14, 25, 36]])
    a *= col
    np.testing.assert_allclose(a, [[11, 22, 33], [28, 50, 72]])
    a -= A(gpu, [1])
    np.testing.assert_allclose(a, [[10, 21, 32], [27, 49, 71]])
    b = A(gpu, [[1, 2], [3, 4]])
    b **= A(gpu, [2, 3])
    np.testing.assert_allclose(b, [[1, 8], [9, 64]])
    b.max(A(gpu, [[5], [10]]), inplace=True)
    np.testing.assert_allclose(b, [[5, 8], [10, 64]])
    with pytest.raises(ValueError):
        row += a


@pytest.mark.parametrize("sa,sb", [
    ((4, 8), (8,)), ((4, 8), (4, 1)), ((5, 7), (7,)), ((5, 7), (5, 1)), ((3, 1, 8), (1, 6, 1)),
    ((2, 3, 4, 5), (3, 1, 5)), ((6, 1, 4), (6, 5, 1)), ((1,), (3, 3)), ((129, 260), (260,)),
    ((129, 260), (129, 1)), ((2, 2, 2, 2, 2, 2, 2, 4), (2, 1, 2, 1, 2, 1, 2, 4)), ((300, 1), (1, 500)),
])
@pytest.mark.parametrize("op", ["add", "mul", "div", "max", "pow"])
def test_broadcast_vs_oracle(gpu, rs, sa, sb, op):
    a = rs.uniform(0.5, 2, sa).astype(F)
    b = rs.uniform(0.5, 2, sb).astype(F)
    f = {"add": lambda x, y: x + y, "mul": lambda x, y: x * y, "div": lambda x, y: x / y,
         "max": lambda x, y: x.max(y), "pow": lambda x, y: x ** y}[op]
    got = np.asarray(f(A(gpu, a), A(gpu, b)))
    want = orc.broadcast_binary(op, a, b)       # literal restatement of add_broadcast.comp's index walk
    if op == "pow":
        np.testing.assert_allclose(got, want, rtol=1.2e-7)
    else:
        np.testing.assert_array_equal(got, want)
    shape = np.broadcast_shapes(sa, sb)
    np.testing.assert_array_equal(np.asarray(A(gpu, b).broadcast_to(shape)), orc.broadcast_to(b, shape))
    if tuple(shape) == tuple(sa) and op != "pow":   # in-place form
        x = A(gpu, a)
        {"add": x.__iadd__, "mul": x.__imul__, "div": x.__itruediv__,
         "max": lambda y: x.max(y, inplace=True)}[op](A(gpu, b))
        np.testing.assert_array_equal(np.asarray(x), want)


# ------------------------------------------------------------------------- gather
def test_gather_known_answers(gpu):
    a = A(gpu, [[1, 2, 3], [4, 5, 6]])
    idx = vk.U32Array(gpu, data=[0, 5, 2])
    np.testing.assert_allclose(a.gather(idx), [1, 6, 3])
    np.testing.assert_allclose(a.gather(vk.U32Array(gpu, data=[1, 0, 1]), axis=0), [[4, 5, 6], [1, 2, 3], [4, 5, 6]])
    g = a.gather(vk.U32Array(gpu, data=[2, 0]), axis=1)
    assert g.shape == (2, 2)           # indices.shape + prev
    np.testing.assert_allclose(g, [[3, 6], [1, 4]])
    oh = vk.U32Array(gpu, data=[2, 0, 1]).to_onehot(3)
    np.testing.assert_allclose(oh, [[0, 0, 1], [1, 0, 0], [0, 1, 0]])
    i2 = vk.U32Array(gpu, data=[[0, 1], [1, 1]])
    assert a.gather(i2, axis=0).shape == (2, 2, 3)


@pytest.mark.parametrize("n_table,n_idx", [(10, 1), (1000, 7), (100000, 4096 * 5 + 3)])
def test_gather_vs_oracle(gpu, rs, n_table, n_idx):
    t = rs.normal(size=n_table).astype(F)
    idx = rs.integers(0, n_table, n_idx, dtype=np.uint32)
    got = np.asarray(A(gpu, t).gather(vk.U32Array(gpu, data=idx)))
    np.testing.assert_array_equal(got, orc.gather(t, idx))          # payload moved bit-exactly
    t3 = rs.normal(size=(5, 11, 13)).astype(F)
    i3 = rs.integers(0, 11, 37, dtype=np.uint32)
    np.testing.assert_array_equal(np.asarray(A(gpu, t3).gather(vk.U32Array(gpu, data=i3), axis=1)),
                                  orc.gather_axis(t3, i3, 1))


def test_u32array_roundtrip_is_bit_exact(gpu, rs):
    v = rs.integers(0, 2 ** 32, 1000, dtype=np.uint64).astype(np.uint32)
    u = vk.U32Array(gpu, data=v)
    np.testing.assert_array_equal(np.asarray(u), v)
    s = vk.Shape(gpu, data=[3, 4, 5])
    np.testing.assert_array_equal(np.asarray(s), [3, 4, 5])
    z = vk.U32Array(gpu, shape=(4,))
    z[:] = 7
    np.testing.assert_array_equal(np.asarray(z), [7, 7, 7, 7])


# ------------------------------------------------------------------------- misc
def test_empty_and_zeros(gpu):
    z = vk.zeros(gpu, (3, 4))
    np.testing.assert_array_equal(np.asarray(z), np.zeros((3, 4)))
    e = vk.Array(gpu, shape=(0,))
    assert np.asarray(e + e).shape == (0,)
    assert np.asarray(e * 2.0).shape == (0,)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 1023, 1024, 1025, 4096 * 7 + 1])
def test_binary_vs_oracle_ragged(gpu, rs, n):
    a = rs.normal(size=n).astype(F)
    b = rs.uniform(0.5, 2, n).astype(F)
    x, y = A(gpu, a), A(gpu, b)
    for op, f in [("add", lambda: x + y), ("sub", lambda: x - y), ("mul", lambda: x * y),
                  ("div", lambda: x / y), ("max", lambda: x.max(y)), ("min", lambda: x.min(y))]:
        np.testing.assert_array_equal(np.asarray(f()), orc.binary(op, a, b))
    np.testing.assert_array_equal(np.asarray(x * 2.5), orc.scalar("mul", a, 2.5))
    np.testing.assert_array_equal(np.asarray(2.5 - x), orc.scalar("sub", a, 2.5, reverse=True))
    np.testing.assert_array_equal(np.asarray(2.5 / y), orc.scalar("div", b, 2.5, reverse=True))
    np.testing.assert_array_equal(np.asarray(x.clamp(-0.5, 0.5)), orc.clamp(a, -0.5, 0.5))


def test_submit_by_spv_path_and_unknown(gpu):
    """The reference names kernels by .spv path (util.py:58-72); both forms resolve."""
    from vulkpy_b200._backend import VectorParams, DataShape
    a, b, c = A(gpu, [1, 2, 3]), A(gpu, [1, 1, 1]), vk.Array(gpu, shape=(3,))
    c.job = gpu._submit("/site-packages/vulkpy/shader/add.spv", 64, 1, 1, [a, b, c], DataShape(3, 1, 1), VectorParams(3))
    np.testing.assert_allclose(c, [2, 3, 4])
    with pytest.raises(RuntimeError):
        gpu._submit("nope.spv", 64, 1, 1, [a], DataShape(3, 1, 1), VectorParams(3))
    from vulkpy_b200.util import getShader
    assert getShader("add.spv") == "add"
