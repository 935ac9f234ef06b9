"""The oracle against the reference's own published vectors and known answers (no GPU).

Pins: vulkpy/random.py:12-24 (Xoshiro128pp(seed=0) docstring), the arithmetic
example of vulkpy/__init__.py:19-24, doc/broadcasting.md, and the known answers that
test/test_vulkpy.py and test/test_nn.py assert (restated here, not copied).  Also cross-checks
the NumPy oracle against the C restatement (oracle/cpu_ref.c) on larger inputs."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import vulkpy_oracle as orc
from oracle import cpu_ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_docstring_vectors_seed0():
    vec = json.load(open(os.path.join(GOLDEN, "reference_vectors.json")))
    r = orc.Xoshiro128pp(64, 0)
    # uniform: integer arithmetic + one exact float subtraction -> bit exact with the printed float32
    np.testing.assert_array_equal(r.random(3), np.asarray(vec["random_3"], dtype=np.float32))
    # normal went through the driver's log/sqrt/sin/cos in the reference: 2 ulp
    np.testing.assert_allclose(r.normal(3), np.asarray(vec["normal_3_after_random_3"], dtype=np.float32),
                               rtol=3e-7)


def test_lane0_state_and_first_draws():
    st = orc.seed_states(64, 0)
    assert [hex(int(v)) for v in st[0]] == ["0x7b1dcdaf", "0x4d197e6f", "0x38fcbe91", "0xaac80268"]
    draws = orc.Xoshiro128pp(64, 0).randint(4)
    assert [hex(int(v)) for v in draws] == ["0x6e05d941", "0xd2d6cb01", "0xe7fea429", "0x2cced82c"]
    # (0x6e05d941 >> 9) | 0x3f800000 is 1.42977667f: the docstring's first uniform
    assert orc.u32_to_unit_float(draws[:1])[0] == np.float32(0.42977667)


def test_jump_differs_from_canonical():
    """The reference writes the accumulators back after each JUMP word (_vkarray.cc:616-619)."""
    s = [1, 2, 3, 4]
    orc.jump_reference(s)
    acc, t = [0, 0, 0, 0], [1, 2, 3, 4]
    for j in orc.JUMP:  # Vigna's jump: a single write-back at the end
        for b in range(32):
            if j & (1 << b):
                acc = [x ^ y for x, y in zip(acc, t)]
            orc.next_scalar(t)
    assert s != acc


def test_stream_layout_and_persistence():
    a = orc.Xoshiro128pp(4, 9)
    first = a.randint(10)       # chunks 4 + 4 + 2
    b = orc.Xoshiro128pp(4, 9)
    lanes = [[], [], [], []]
    st = b.state
    for c in range(3):
        m = min(4, 10 - 4 * c)
        d = orc.next_lanes(st, m)
        for lane in range(m):
            lanes[lane].append(int(d[lane]))
    assert list(first[0::4]) == lanes[0] and list(first[1::4]) == lanes[1]
    assert list(first[2::4]) == lanes[2][:2] and list(first[3::4]) == lanes[3][:2]
    # n <= size advances only the first n lanes
    before = a.state.copy()
    a.randint(2)
    assert (a.state[2:] == before[2:]).all() and (a.state[:2] != before[:2]).any()


def test_golden_streams_regenerate():
    g = np.load(os.path.join(GOLDEN, "prng_streams.npz"))
    r = orc.Xoshiro128pp(64, 0)
    np.testing.assert_array_equal(r.randint(3), g["s64_seed0_call0_u32_3"])
    np.testing.assert_array_equal(r.random(17), g["s64_seed0_call1_f32_17"])


def test_c_prng_matches_numpy_oracle():
    L = cpu_ref.load()
    for size, seed, n in [(64, 0, 5000), (7, 3, 100), (256, 11, 256 * 40 + 3)]:
        st = np.zeros((size, 4), np.uint32)
        L.ref_xoshiro_seed(cpu_ref.ptr(st), size, C.c_uint64(seed))
        np.testing.assert_array_equal(st, orc.seed_states(size, seed))
        out = np.zeros(n, np.uint32)
        L.ref_xoshiro_fill(cpu_ref.ptr(st), size, cpu_ref.ptr(out), C.c_uint64(n), 0)
        o = orc.Xoshiro128pp(size, seed)
        np.testing.assert_array_equal(out, o.randint(n))
        np.testing.assert_array_equal(st, o.state)


def test_randrange_known_answer():
    # test/test_random.py:99-103: randrange(low=3, high=4) is always 3
    u = orc.Xoshiro128pp(64, 1).random(5)
    assert (orc.randrange_shader(u, 3, 3) == 3).all()
    # Q23: range > 2^23 only reaches multiples of range / 2^23
    v = orc.randrange_shader(orc.Xoshiro128pp(64, 1).random(64), 0, (1 << 26) - 1)
    assert (v % 8 == 0).all()


def test_arithmetic_docstring_example():
    # vulkpy/__init__.py:19-24: [1,2,3] + [3,3,3] = [4,5,6]
    np.testing.assert_array_equal(orc.binary("add", [1, 2, 3], [3, 3, 3]), [4, 5, 6])


@pytest.mark.parametrize("op,a,b,want", [
    ("sub", [4, 4, 4], [2, 2, 2], [2, 2, 2]),
    ("mul", [2, 2, 2], [3, 3, 3], [6, 6, 6]),
    ("div", [8, 8, 8], [2, 2, 2], [4, 4, 4]),
    ("max", [1, 5, 3], [4, 2, 3], [4, 5, 3]),
    ("min", [1, 5, 3], [4, 2, 3], [1, 2, 3]),
])
def test_binary_known_answers(op, a, b, want):
    np.testing.assert_array_equal(orc.binary(op, a, b), want)


def test_tight_tolerance_points_appendix_a():
    """The reference asserts these at rtol=1e-7 against float64 (test/test_vulkpy.py:551-702)."""
    x = np.array([1, 2, 3], dtype=np.float32)
    np.testing.assert_allclose(orc.unary("exp", x), np.exp([1., 2., 3.]), rtol=1e-7)
    np.testing.assert_allclose(orc.unary("log", x), np.log([1., 2., 3.]), rtol=1e-7)
    np.testing.assert_array_equal(orc.unary("exp2", x), [2, 4, 8])
    np.testing.assert_allclose(orc.unary("invsqrt", x), 1 / np.sqrt([1., 2., 3.]), rtol=1e-7)
    np.testing.assert_allclose(orc.binary("pow", x, [1.1, 2.2, 1.4]),
                               np.power(x.astype(np.float64), np.float32([1.1, 2.2, 1.4]).astype(np.float64)),
                               rtol=1e-7)


def test_broadcast_matches_numpy_and_doc():
    # doc/broadcasting.md: (2,3) with (3,), (2,1) with (1,3), scalar-like
    rs = np.random.default_rng(0)
    for sa, sb in [((2, 3), (3,)), ((2, 1), (1, 3)), ((4, 1, 5), (3, 1)), ((1,), (2, 2)), ((2, 3, 4), (2, 3, 4))]:
        a = rs.normal(size=sa).astype(np.float32)
        b = rs.normal(size=sb).astype(np.float32)
        np.testing.assert_array_equal(orc.broadcast_binary("add", a, b), a + b)
        np.testing.assert_array_equal(orc.broadcast_to(b, np.broadcast_shapes(sa, sb)),
                                      np.broadcast_to(b, np.broadcast_shapes(sa, sb)))


def test_c_broadcast_matches_numpy_oracle():
    L = cpu_ref.load()
    rs = np.random.default_rng(1)
    a = rs.normal(size=(5, 1, 7)).astype(np.float32)
    b = rs.normal(size=(3, 1)).astype(np.float32)
    shape = np.broadcast_shapes(a.shape, b.shape)
    sh = np.ones(9, np.uint32)
    sh[0:3] = a.shape
    sh[4:6] = b.shape
    sh[6:9] = shape
    c = np.zeros(shape, np.float32)
    L.ref_broadcast_binary(2, cpu_ref.ptr(a), cpu_ref.ptr(b), cpu_ref.ptr(c), cpu_ref.ptr(sh), a.size, b.size,
                           c.size, 3)
    np.testing.assert_array_equal(c, orc.broadcast_binary("mul", a, b))


def test_reductions():
    a = np.arange(1, 25, dtype=np.float32).reshape(2, 3, 4)
    np.testing.assert_array_equal(orc.reduce_axis("sum", a, 1), a.sum(axis=1))
    np.testing.assert_array_equal(orc.reduce_axis("maximum", a, 2), a.max(axis=2))
    np.testing.assert_array_equal(orc.reduce_axis("prod", a[:, :, :2], 0), a[:, :, :2].prod(axis=0))
    rb = orc.reduce_axis("sum", a, 1, rebroadcast=True)
    assert rb.shape == a.shape and (rb[:, 0, :] == a.sum(axis=1)).all()
    # test/test_vulkpy.py:822-827 style: 65 elements crosses one workgroup
    np.testing.assert_array_equal(orc.reduce_full_reference("sum", np.ones(65)), [65])
    np.testing.assert_array_equal(orc.reduce_full_reference("maximum", np.arange(65)), [64])
    L = cpu_ref.load()
    x = np.random.default_rng(2).uniform(0, 1, (33, 65, 17)).astype(np.float32)
    out = np.zeros((33, 17), np.float32)
    L.ref_reduce_axis(0, cpu_ref.ptr(x), cpu_ref.ptr(out), 33, 65, 17, 0)
    np.testing.assert_array_equal(out, orc.reduce_axis("sum", x, 1))


def test_gather_and_onehot():
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    np.testing.assert_array_equal(orc.gather(a, [0, 5, 11]), [0, 5, 11])
    np.testing.assert_array_equal(orc.gather_axis(a, [2, 0], 0), a[[2, 0]])
    got = orc.gather_axis(a, [3, 1], 1)           # result is [idx, prev]
    np.testing.assert_array_equal(got, a[:, [3, 1]].T)
    np.testing.assert_array_equal(orc.gather_axis(np.identity(3, dtype=np.float32), [1, 0, 2, 1], 0),
                                  np.identity(3)[[1, 0, 2, 1]])


def test_matmul_and_affine():
    # test/test_vulkpy.py:225-255
    np.testing.assert_array_equal(orc.matmul([[1, 2], [3, 4]], [[1, 2], [3, 4]]), [[7, 10], [15, 22]])
    np.testing.assert_array_equal(orc.matmul([[1, 2], [3, 4]], [1, 3]), [7, 15])
    np.testing.assert_array_equal(orc.matmul([1, 2], [[1, 2], [3, 4]]), [7, 10])
    rs = np.random.default_rng(3)
    w, b, x = rs.normal(size=(5, 7)), rs.normal(size=5), rs.normal(size=(4, 7))
    np.testing.assert_allclose(orc.batch_affine(w, b, x), x @ w.T + b, rtol=1e-5, atol=1e-6)
    L = cpu_ref.load()
    a32, b32 = rs.normal(size=(19, 33)).astype(np.float32), rs.normal(size=(33, 21)).astype(np.float32)
    c = np.zeros((19, 21), np.float32)
    L.ref_matmul(cpu_ref.ptr(a32), cpu_ref.ptr(b32), cpu_ref.ptr(c), 19, 33, 21)
    np.testing.assert_array_equal(c, orc.matmul(a32, b32))


def test_cross_entropy_known_answer():
    # test/test_nn.py:216-223: -log(0.5) = 0.6931472
    np.testing.assert_allclose(orc.cross_entropy([[0.5, 0.5]], [[1.0, 0.0]]).sum(), 0.6931472, rtol=1e-6)
    np.testing.assert_allclose(orc.cross_entropy_backward([[0.5, 0.5]], [[1.0, 0.0]]), [[-2.0, 0.0]], rtol=1e-6)


def test_arg_and_permutation_oracle():
    x = np.array([[1, 5, 5], [np.nan, 2, -1]], np.float32)
    np.testing.assert_array_equal(orc.argmax(x, 1), [1, 0])
    np.testing.assert_array_equal(orc.argmin(x, 1), [0, 0])
    np.testing.assert_array_equal(orc.argmax(x), [3])
    assert orc.argmax(x, 0).dtype == np.uint32
    np.testing.assert_array_equal(orc.permutation_from_keys([5, 1, 5, 0]), [3, 1, 0, 2])
