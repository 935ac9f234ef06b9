"""One process driving two GPUs (`GPU(0)` then `GPU(1)`: vkarray.py:89-101 of the reference only picks a device):
per-device kernel attributes (the shared-memory opt-ins of gemm_skinny_n and reduce_cols_tma) must be set on
BOTH devices.  Skipped on a single-GPU box."""
import numpy as np
import pytest

import vulkpy_b200 as vk
from vulkpy_b200 import nn
from vulkpy_b200._backend import device_count

pytestmark = pytest.mark.gpu
F = np.float32


def test_opt_in_kernels_on_second_device_after_first():
    if device_count() < 2:       # asked at run time: on a box without a driver the query itself raises
        pytest.skip("needs two GPUs")
    rs = np.random.default_rng(0)
    x_h = rs.uniform(0, 1, (1024, 384)).astype(F)            # TMA-staged column reduction (96 KiB shared memory)
    a_h = rs.normal(size=(2048, 1024)).astype(F)             # skinny-N Dense forward (64 KiB shared memory)
    for idx in (0, 1):
        gpu = vk.GPU(idx)
        x = vk.Array(gpu, data=x_h)
        np.testing.assert_allclose(np.asarray(x.sum(axis=0)), x_h.astype(np.float64).sum(axis=0), rtol=2e-6)
        d = nn.Dense(gpu, 1024, 16, w_init=nn.HeNormal(gpu, 1024, seed=2))
        y = np.asarray(d(vk.Array(gpu, data=a_h)))
        want = a_h.astype(np.float64) @ np.asarray(d.w.value).astype(np.float64).T + np.asarray(d.b.value)
        np.testing.assert_allclose(y, want, rtol=1e-4, atol=1e-4)


def test_tall_matvec_beyond_65535_row_tiles(gpu):
    """`Array(2^23, 32) @ Array(32,)`: 131072 row tiles of the SIMT kernel (grid.x, not grid.y: ADVICE r1)."""
    rs = np.random.default_rng(1)
    a_h = rs.uniform(-1, 1, (1 << 23, 32)).astype(F)
    v_h = rs.uniform(-1, 1, 32).astype(F)
    got = np.asarray(vk.Array(gpu, data=a_h) @ vk.Array(gpu, data=v_h))
    np.testing.assert_allclose(got, a_h.astype(np.float64) @ v_h, rtol=0, atol=2e-5)
