"""64-lane (and 2^20-lane) uniform generator throughput against the request size: separates the fixed cost
(jump-ahead / start states) from effects that grow with the footprint (concurrent segment streams far apart)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200._backend import Timer

gpu = vk.GPU(0)
dev = gpu.gpu
for size in (64, 1 << 20):
    for log2n in (24, 26, 28, 30):
        n = 1 << log2n
        buf = vk.U32Array(gpu, shape=(n,))
        g = vk.random.Xoshiro128pp(gpu, size=size, seed=7)
        for _ in range(3):
            g.randint(buffer=buf)
        gpu.wait()
        reps = max(5, (1 << 32) // n // 4)
        t0, t1 = Timer(dev), Timer(dev)
        t0.record()
        for _ in range(reps):
            g.randint(buffer=buf)
        t1.record()
        ms = t0.elapsed_ms(t1) / reps
        print(f"size={size:8d} n=2^{log2n}  {ms*1e3:9.1f} us  {4.0*n/ms/1e6:8.1f} GB/s", flush=True)
        del buf
