"""Row-sharded matmul timing at the 8-GPU range geometry (32 MiB ranges, 4 KB rows), runnable on 2 GPUs:
    python -m torch.distributed.run --nproc-per-node 2 ... scripts/pull_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200 import dist
from vulkpy_b200._backend import Timer

g = dist.Group.from_env()
gpu, rank, world = g.gpu, g.rank, g.world
rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=5)
out = {"world": world}
for (M, N, K) in ((1024 * world, 8192, 1024 * world), (4096 * world, 8192, 8192)):
    A = g.random(rng, (M, K), "random"); Bm = g.random(rng, (K, N), "random")
    for _ in range(3):
        c = A @ Bm; del c
    gpu.wait()
    t0, t1 = Timer(gpu.gpu), Timer(gpu.gpu)
    t0.record()
    for _ in range(5):
        c = A @ Bm; del c
    t1.record()
    ms = t0.elapsed_ms(t1) / 5
    per_rank_bytes = (world - 1) * (K // world) * N * 4
    out[f"{M}x{N}x{K}"] = {"ms": round(ms, 4), "pulled_MB_per_rank": round(per_rank_bytes / 1e6, 1),
                           "tflops_agg": round(2 * M * N * K / ms / 1e9, 1)}
    del A, Bm
if rank == 0:
    print(json.dumps(out))
g.t.close()
