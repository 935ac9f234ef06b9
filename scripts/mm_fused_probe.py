"""Row-sharded fixed-size 8192^3 matmul: time of the fused call under the switches given in the environment
(torchrun; prints on rank 0).  VKP_COMM_NO_PULL=1 gives the same staging + GEMM with nothing crossing NVLink."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as td
import vulkpy_b200 as vk
from vulkpy_b200 import dist
from vulkpy_b200._backend import Timer

lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
td.init_process_group("nccl", device_id=torch.device("cuda", lr))
gpu = vk.GPU(lr)
g = dist.Group.from_env()
M = 8192
rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=5)
A = g.random(rng, (M, M), "random")
B = g.random(rng, (M, M), "random")
for _ in range(5):
    C = A @ B
gpu.wait(); td.barrier(); torch.cuda.synchronize()
t0, t1 = Timer(gpu.gpu), Timer(gpu.gpu)
t0.record()
for _ in range(40):
    C = A @ B
t1.record()
ms = t0.elapsed_ms(t1) / 40
t = torch.tensor([ms], device="cuda", dtype=torch.float64)
td.all_reduce(t, op=td.ReduceOp.MAX)
if g.rank == 0:
    tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("VKP_"))
    print(f"fused row-sharded 8192^3 on {g.world} GPUs [{tag or 'default'}]: {t.item():.4f} ms", flush=True)
g.t.close()
