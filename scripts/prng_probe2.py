"""64-lane generator, 2^28 samples: per-call device time back to back vs isolated, fresh vs reused output buffer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200._backend import Timer

gpu = vk.GPU(0)
dev = gpu.gpu
n = 1 << 28
for size in (64, 1 << 20):
    bufs = [vk.U32Array(gpu, shape=(n,)) for _ in range(4)]
    g = vk.random.Xoshiro128pp(gpu, size=size, seed=7)
    for b in bufs:
        g.randint(buffer=b)
    gpu.wait()
    for mode in ("back-to-back same buffer", "back-to-back rotating buffers", "isolated (wait between)"):
        ts = [Timer(dev) for _ in range(13)]
        ts[0].record()
        for i in range(12):
            g.randint(buffer=bufs[0] if "same" in mode or "isolated" in mode else bufs[i % 4])
            if "isolated" in mode:
                gpu.wait()
            ts[i + 1].record()
        gpu.wait()
        d = [ts[i].elapsed_ms(ts[i + 1]) * 1e3 for i in range(12)]
        print(f"size={size:8d} {mode:32s} us per call:", " ".join(f"{x:6.1f}" for x in d), flush=True)
    del bufs
