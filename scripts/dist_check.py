"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
Every sharded result is compared with the single-GPU / NumPy answer; also times the sharded C2 add,
the all-reduced full sum and the all-gather matmul (device events, max over ranks)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200 import dist, nn
from vulkpy_b200._backend import Timer

g = dist.Group.from_env()
gpu, rank, world = g.gpu, g.rank, g.world
F = np.float32
rs = np.random.default_rng(0)
out = {"world": world}

# ---- parity on moderate sizes ------------------------------------------------------------------
R, Ccols = 64 * world, 96
a_f = rs.uniform(0.5, 2, (R, Ccols)).astype(F)
b_f = rs.uniform(0.5, 2, (R, Ccols)).astype(F)
a, b = g.shard(a_f), g.shard(b_f)
np.testing.assert_array_equal((a + b).to_numpy(), a_f + b_f)
np.testing.assert_array_equal((a * 2.5 - b).to_numpy(), (a_f * F(2.5)) - b_f)
np.testing.assert_allclose(np.asarray(a.sum()), [a_f.astype(np.float64).sum()], rtol=2e-6)
np.testing.assert_array_equal(np.asarray(a.maximum()), [a_f.max()])
np.testing.assert_allclose(np.asarray(a.sum(axis=0)), a_f.astype(np.float64).sum(axis=0), rtol=2e-6)
np.testing.assert_allclose(a.sum(axis=1).to_numpy(), a_f.astype(np.float64).sum(axis=1), rtol=2e-6)
np.testing.assert_allclose(np.asarray(a.mean()), [a_f.astype(np.float64).mean()], rtol=2e-6)
np.testing.assert_allclose(a.maximum(axis=0, rebroadcast=True).to_numpy(),
                           np.broadcast_to(a_f.max(axis=0, keepdims=True), a_f.shape))
K = 32 * world
A_f, B_f = rs.uniform(-1, 1, (R, K)).astype(F), rs.uniform(-1, 1, (K, 80)).astype(F)
C = g.shard(A_f) @ g.shard(B_f)
np.testing.assert_allclose(C.to_numpy(), A_f.astype(np.float64) @ B_f, atol=1e-4)

# fused pull + GEMM path (vkp_comm_matmul_allgather): local rows >= 128, K/world a multiple of 32
Mg, Kg, Ng = 256 * world, 64 * world, 384
A_f, B_f = rs.uniform(-1, 1, (Mg, Kg)).astype(F), rs.uniform(-1, 1, (Kg, Ng)).astype(F)
for rep in range(3):      # three calls: both staging copies and the epoch flags get reused
    C = g.shard(A_f) @ g.shard(B_f)
    want = A_f.astype(np.float64) @ B_f
    mag = np.abs(A_f).astype(np.float64) @ np.abs(B_f)
    err = float((np.abs(C.to_numpy() - want) / mag).max())
    assert err < 6e-6, err
    B_f = B_f + F(0.25)    # new contents every call: a stale staging copy would show
out["fused_matmul_used"] = not g.t._fused_broken
out["fused_matmul_err"] = err

# ---- sharded PRNG == single-GPU stream -------------------------------------------------------------
shape = (16 * world, 256)
ref = np.asarray(vk.random.Xoshiro128pp(gpu, seed=11).random(shape=shape))
r_sh = vk.random.Xoshiro128pp(gpu, seed=11)
sh = g.random(r_sh, shape, "random")
np.testing.assert_array_equal(sh.to_numpy(), ref)
r_ref = vk.random.Xoshiro128pp(gpu, seed=11); r_ref.random(shape=shape)
np.testing.assert_array_equal(r_sh.rng.state(), r_ref.rng.state())       # same final state on every rank
refn = np.asarray(vk.random.Xoshiro128pp(gpu, seed=12).normal(shape=shape))
np.testing.assert_array_equal(g.random(vk.random.Xoshiro128pp(gpu, seed=12), shape, "normal").to_numpy(), refn)

# ---- data-parallel nn step == single-process full batch ----------------------------------------------
def make_net():
    sgd = nn.SGD(0.05)
    return nn.Sequence([nn.Dense(gpu, 16, 32, w_opt=sgd, b_opt=sgd, w_init=nn.HeNormal(gpu, 16, seed=1)), nn.ReLU(),
                        nn.Dense(gpu, 32, 4, w_opt=sgd, b_opt=sgd, w_init=nn.HeNormal(gpu, 32, seed=2))],
                       nn.SoftmaxCrossEntropyLoss())
Bg = 8 * world
x_f = rs.normal(size=(Bg, 16)).astype(F)
y_f = np.eye(4, dtype=F)[rs.integers(0, 4, Bg)]
single = make_net()
single.train(vk.Array(gpu, data=x_f), vk.Array(gpu, data=y_f))
lo, hi = g.bounds(Bg)
net = make_net()
dp = dist.DataParallel(net, g)
_, loss = dp.train(vk.Array(gpu, data=x_f[lo:hi]), vk.Array(gpu, data=y_f[lo:hi]))
for ls, ld in zip(single.L, net.L):
    if hasattr(ls, "w"):
        np.testing.assert_allclose(np.asarray(ld.w.value), np.asarray(ls.w.value), rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(np.asarray(ld.b.value), np.asarray(ls.b.value), rtol=2e-5, atol=1e-6)
out["parity"] = "ok"

# ---- timings (weak scaling: 2^28 elements per GPU) ------------------------------------------------------
def timed(fn, reps=10):
    for _ in range(3):
        r = fn(); del r
    gpu.wait()
    t0, t1 = Timer(gpu.gpu), Timer(gpu.gpu)
    t0.record()
    for _ in range(reps):
        r = fn(); del r
    t1.record()
    ms = t0.elapsed_ms(t1) / reps
    part = vk.Array(gpu, data=[ms]); g.t.allreduce(part, "maximum")
    return float(np.asarray(part)[0])

rows = 16384
big = (rows * world, 16384)
rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=5)
x = g.random(rng, big, "random")
y = g.random(rng, big, "random")
n_total = big[0] * big[1]
ms = timed(lambda: x + y); out["a+b"] = {"ms": round(ms, 4), "agg_gbs": round(12 * n_total / ms / 1e6, 1)}
ms = timed(lambda: x.sum()); out["sum(None)+allreduce"] = {"ms": round(ms, 4), "agg_gbs": round(4 * n_total / ms / 1e6, 1)}
ms = timed(lambda: x.sum(axis=0)); out["sum(axis=0)+allreduce"] = {"ms": round(ms, 4), "agg_gbs": round(4 * n_total / ms / 1e6, 1)}
ms = timed(lambda: x.sum(axis=1)); out["sum(axis=1)"] = {"ms": round(ms, 4), "agg_gbs": round(4 * n_total / ms / 1e6, 1)}
del x, y
M = 8192
A = g.random(rng, (M, M), "random"); Bm = g.random(rng, (M, M), "random")
ms = timed(lambda: A @ Bm, reps=5)
out["matmul 8192^3 row-sharded, peers' B pulled inside one GEMM"] = {"ms": round(ms, 3), "agg_tflops": round(2 * M ** 3 / ms / 1e9, 1)}
g.t._fused_broken = True          # the NCCL all-gather + GEMM path, for comparison
ms = timed(lambda: A @ Bm, reps=5)
out["matmul 8192^3 row-sharded, ncclAllGather(B) then GEMM"] = {"ms": round(ms, 3), "agg_tflops": round(2 * M ** 3 / ms / 1e9, 1)}
g.t._fused_broken = False
del A, Bm

# weak-scaling matmul: 8192 rows of A per GPU (M = 8192 * world), B 8192 x 8192 sharded by rows
A = g.random(rng, (M * world, M), "random"); Bm = g.random(rng, (M, M), "random")
ms = timed(lambda: A @ Bm, reps=3)
out["matmul (8192*world) x 8192 x 8192 weak scaling, fused"] = {"ms": round(ms, 3), "agg_tflops": round(2 * world * M ** 3 / ms / 1e9, 1)}
del A, Bm

# data-parallel MLP step (config 5): Dense(1024,1024)-ReLU-Dense(1024,16)-Softmax, Adam, 8192 rows per GPU
def make_mlp():
    opt = lambda: nn.Adam(gpu, lr=1e-3)
    return nn.Sequence([nn.Dense(gpu, 1024, 1024, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 1024, seed=1)), nn.ReLU(),
                        nn.Dense(gpu, 1024, 16, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 1024, seed=2)), nn.Softmax()],
                       nn.CrossEntropyLoss())
Bl = 8192
xb = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=100 + rank).normal(shape=(Bl, 1024))
yb = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=200 + rank).randrange(shape=(Bl,), low=0, high=16).to_onehot(16)
mlp = make_mlp()
dpm = dist.DataParallel(mlp, g)
ms = timed(lambda: dpm.train(xb, yb)[1], reps=10)
out["DP MLP step (8192 rows/GPU)"] = {"ms": round(ms, 4), "agg_rows_per_s": round(Bl * world / ms * 1e3, 1)}
single_mlp = make_mlp()
ms1 = timed(lambda: single_mlp.train(xb, yb)[1], reps=10)
out["MLP step without exchange"] = {"ms": round(ms1, 4)}
if rank == 0:
    print(json.dumps(out))
g.t.close()
