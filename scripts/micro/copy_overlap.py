"""Do the two copy engines overlap?  Times H2D alone, D2H alone and both in flight together."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import vulkpy_b200 as vk

gpu = vk.GPU(0)
n = 1 << 28
h_in = vk.pinned_empty((n,)); h_in[...] = 1.0
h_out = vk.pinned_empty((n,))
src = vk.Array(gpu, shape=(n,)); src += 1.0
gpu.wait()

def t(label, fn):
    gpu.wait()
    t0 = time.perf_counter(); r = fn(); t1 = time.perf_counter(); gpu.wait(); t2 = time.perf_counter()
    print(f"{label:34s} enqueue {1e3*(t1-t0):8.2f} ms   total {1e3*(t2-t0):8.2f} ms", flush=True)
    return r

for rep in range(2):
    a = t("H2D 1 GiB (from_host)", lambda: vk.Array.from_host(gpu, h_in))
    t("D2H 1 GiB (to_host wait=False)", lambda: src.to_host(h_out, wait=False))
    def both():
        x = vk.Array.from_host(gpu, h_in)
        src.to_host(h_out, wait=False)
        return x
    b = t("H2D + D2H together", both)
    def sync_up():
        return vk.Array(gpu, data=h_in)
    c = t("H2D 1 GiB (Array(data=pinned))", sync_up)
    t("D2H 1 GiB (to_host sync)", lambda: src.to_host(h_out))
    del a, b, c
