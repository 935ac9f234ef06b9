"""Small driver for ncu captures (round 2): one launch of each kernel the verdict asked evidence for.
  python scripts/r02_probe.py [names...]   names: pow log prng reduce gather (default: all)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk

which = set(sys.argv[1:]) or {"pow", "log", "prng", "reduce", "gather"}
gpu = vk.GPU(0)
R = 16384
rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=1234)
a = rng.random(shape=(R, R)); a *= 1.5; a += 0.5
b = rng.random(shape=(R, R)); b *= 4.0; b -= 2.0
gpu.wait()
for rep in range(2):
    if "pow" in which:
        (a ** b).wait(); (a ** 2.7).wait(); (1.3 ** b).wait()
    if "log" in which:
        a.log().wait(); a.log2().wait(); b.exp2().wait(); b.exp().wait(); b.asinh().wait()
    if "reduce" in which:
        a.sum(axis=0).wait(); a.maximum(axis=0).wait(); a.sum(axis=1).wait(); a.sum().wait()
    if "prng" in which:
        buf = vk.Array(gpu, shape=(1 << 28,)); ubuf = vk.U32Array(gpu, shape=(1 << 28,))
        for size in (64, 1 << 20):
            g = vk.random.Xoshiro128pp(gpu, size=size, seed=7)
            g.random(buffer=buf).wait(); g.randint(buffer=ubuf).wait(); g.normal(buffer=buf).wait()
        del buf, ubuf
    if "gather" in which:
        G, NI = 8192, 1 << 26
        table = rng.random(shape=(G, G))
        idx_h = np.random.default_rng(99).integers(0, G * G, NI, dtype=np.uint32)
        idx = vk.U32Array(gpu, data=idx_h)
        table.gather(idx).wait()
        idx_s = vk.U32Array(gpu, data=np.sort(idx_h))
        table.gather(idx_s).wait()
        del table, idx, idx_s
gpu.wait()
print("probe done")
