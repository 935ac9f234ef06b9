"""argmax / argmin / permutation / shuffle on the device against NumPy (what the reference's own
example falls back to: example/02-nn.py:82,96).  Index work: bit-exact."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vulkpy_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vk():
    import vulkpy_b200 as vk
    return vk


@pytest.fixture(scope="module")
def gpu(vk):
    return vk.GPU(0)


SHAPES = [((7,), None), ((7,), 0), ((1,), None), ((4099,), None), ((1 << 20,), 0), ((3, 5), None), ((3, 5), 0),
          ((3, 5), 1), ((3, 5), -1), ((1000, 16), 1), ((1000, 16), 0), ((33, 1000), 1), ((33, 4097), 1),
          ((5, 70001), 1), ((70001, 5), 0), ((6, 7, 8), 1), ((6, 7, 8), 0), ((6, 7, 8), 2), ((2, 3000, 130), 1),
          ((300, 2, 2), 0), ((70000, 3), 1)]


@pytest.mark.parametrize("shape,axis", SHAPES)
@pytest.mark.parametrize("kind", ["random", "ties", "nan"])
def test_arg_against_numpy(vk, gpu, shape, axis, kind):
    rng = np.random.default_rng(abs(hash((shape, axis, kind))) % (1 << 31))
    if kind == "ties":
        x = rng.integers(0, 4, shape).astype(np.float32)
    else:
        x = rng.standard_normal(shape).astype(np.float32)
    if kind == "nan":
        flat = x.reshape(-1)
        flat[rng.integers(0, flat.size, max(1, flat.size // 50))] = np.nan
    a = vk.Array(gpu, data=x)
    for name in ("argmax", "argmin"):
        got = np.asarray(getattr(a, name)(axis=axis))
        want = getattr(O, name)(x, axis)
        assert got.dtype == np.uint32 and got.shape == want.shape
        np.testing.assert_array_equal(got, want)


def test_arg_full_size(vk, gpu):
    """BASELINE size: 16384^2, all three layouts; plus a planted extreme at the very end."""
    R = 16384
    r = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=5)
    a = r.random(shape=(R, R))
    x = a.to_host()
    np.testing.assert_array_equal(a.argmax(axis=1).to_host(), np.argmax(x, axis=1).astype(np.uint32))
    np.testing.assert_array_equal(a.argmin(axis=0).to_host(), np.argmin(x, axis=0).astype(np.uint32))
    np.testing.assert_array_equal(a.argmax().to_host(), [np.argmax(x)])
    x[-1, -1] = 2.0
    b = vk.Array(gpu, data=x)
    assert int(np.asarray(b.argmax())[0]) == R * R - 1


def test_arg_errors(vk, gpu):
    a = vk.Array(gpu, data=np.zeros((2, 3), np.float32))
    with pytest.raises(ValueError):
        a.argmax(axis=2)


@pytest.mark.parametrize("n", [0, 1, 2, 63, 64, 1000, 1 << 16, (1 << 22) + 5])
def test_permutation(vk, gpu, n):
    r1 = vk.random.Xoshiro128pp(gpu, seed=11)
    r2 = vk.random.Xoshiro128pp(gpu, seed=11)
    p = r1.permutation(n)
    keys = np.asarray(r2.randint(shape=(n,))) if n else np.zeros(0, np.uint32)
    got = np.asarray(p) if n else np.zeros(0, np.uint32)
    assert p.shape == (n,)
    np.testing.assert_array_equal(got, O.permutation_from_keys(keys))
    np.testing.assert_array_equal(np.sort(got), np.arange(n, dtype=np.uint32))
    # the generator advanced by exactly n draws
    np.testing.assert_array_equal(np.asarray(r1.randint(shape=(8,))), np.asarray(r2.randint(shape=(8,))))


def test_shuffle_rows(vk, gpu):
    x = np.arange(1000 * 3, dtype=np.float32).reshape(1000, 3)
    a = vk.Array(gpu, data=x)
    r1 = vk.random.Xoshiro128pp(gpu, seed=3)
    r2 = vk.random.Xoshiro128pp(gpu, seed=3)
    s = np.asarray(a.shuffle(r1))
    perm = np.asarray(r2.permutation(1000))
    np.testing.assert_array_equal(s, x[perm])
    assert not np.array_equal(s, x)
