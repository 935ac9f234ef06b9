"""Small shapes that exercise the mbarrier / TMA / tcgen05 / flag-protocol kernels, for
`compute-sanitizer --tool memcheck|racecheck|synccheck python scripts/sanitize_targets.py <target>`
(the reference's analogue is util.enable_debug's validation layers, vulkpy/util.py:19-55).
Targets: gemm (MODE_CONVERT, split-K; set VKP_TC_PRESPLIT=1 / VKP_TC_REWRITE_HI=1 for the other modes),
reduce (TMA-staged column reduction), ew (element-wise / broadcast / gather / PRNG), pull (2 ranks, torchrun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk

target = sys.argv[1] if len(sys.argv) > 1 else "gemm"
rs = np.random.default_rng(0)
F = np.float32


def check_mm(gpu, M, N, K, flags=2):
    a_h, b_h = rs.uniform(-1, 1, (M, K)).astype(F), rs.uniform(-1, 1, (K, N)).astype(F)
    a, b = vk.Array(gpu, data=a_h), vk.Array(gpu, data=b_h)
    c = vk.Array(gpu, shape=(M, N))
    c.job = gpu.gpu.gemm(False, False, M, N, K, a.buffer, b.buffer, c.buffer, None, flags)
    want = a_h.astype(np.float64) @ b_h
    mag = np.abs(a_h).astype(np.float64) @ np.abs(b_h)
    err = float((np.abs(np.asarray(c) - want) / mag).max())
    assert err < 6e-6, err
    return err


if target == "gemm":
    gpu = vk.GPU(0)
    print("tc 256x256x256", check_mm(gpu, 256, 256, 256))
    print("tc 128x384x96 (ragged k-blocks)", check_mm(gpu, 128, 384, 96))
    print("tc split-K 128x128x4096", check_mm(gpu, 128, 128, 4096))
    print("tc 2 tiles/CTA 1280x2048x64", check_mm(gpu, 1280 * 2, 2048, 64))
elif target == "reduce":
    gpu = vk.GPU(0)
    x_h = rs.uniform(0, 1, (1024, 384)).astype(F)
    x = vk.Array(gpu, data=x_h)
    np.testing.assert_allclose(np.asarray(x.sum(axis=0)), x_h.astype(np.float64).sum(axis=0), rtol=2e-6)
    np.testing.assert_array_equal(np.asarray(x.maximum(axis=0)), x_h.max(axis=0))
    x3_h = rs.uniform(0, 1, (3, 700, 256)).astype(F)
    x3 = vk.Array(gpu, data=x3_h)
    np.testing.assert_allclose(np.asarray(x3.sum(axis=1)), x3_h.astype(np.float64).sum(axis=1), rtol=2e-6)
    print("reduce_cols_tma ok")
elif target == "ew":
    gpu = vk.GPU(0)
    a_h, b_h = rs.uniform(0.5, 2, (257, 131)).astype(F), rs.uniform(-2, 2, (257, 131)).astype(F)
    a, b = vk.Array(gpu, data=a_h), vk.Array(gpu, data=b_h)
    big = vk.Array(gpu, data=rs.uniform(0.5, 2, (1 << 15) + 4103).astype(F))
    (big ** 2.7).wait()   # binomial-series kernel (shared-memory tables)
    for r in (a + b, a ** b, a ** 2.7, a.log(), b.exp(), a + vk.Array(gpu, data=b_h[0]), a.sum(axis=1), a.sum(),
              a.gather(vk.U32Array(gpu, data=rs.integers(0, a_h.size, 999, dtype=np.uint32)))):
        r.wait()
    g = vk.random.Xoshiro128pp(gpu, seed=3)
    g.random(shape=(100000,)).wait(); g.normal(shape=(100001,)).wait()
    print("ew ok")
elif target == "pull":
    from vulkpy_b200 import dist
    g = dist.Group.from_env()
    Mg, Kg, Ng = 256 * g.world, 64 * g.world, 384
    A_f, B_f = rs.uniform(-1, 1, (Mg, Kg)).astype(F), rs.uniform(-1, 1, (Kg, Ng)).astype(F)
    for rep in range(3):
        C = g.shard(A_f) @ g.shard(B_f)
        want = A_f.astype(np.float64) @ B_f
        mag = np.abs(A_f).astype(np.float64) @ np.abs(B_f)
        err = float((np.abs(C.to_numpy() - want) / mag).max())
        assert err < 6e-6, err
        B_f = B_f + F(0.25)
    print("pull ok, fused used:", not g.t._fused_broken)
    g.t.close()
