"""Element-wise arithmetic, broadcasting and an axis reduction on 4096 x 4096 float32 arrays
(BASELINE.json configs[0]: the reference's example/00-arithmetic.py scaled up), timed on the device.

    python examples/00_arithmetic.py [--n 4096]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk   # `import vulkpy as vk` works too with the repository on sys.path

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
args = ap.parse_args()
n = args.n

gpu = vk.GPU()
rng = np.random.default_rng(0)
a_h = rng.uniform(0, 1, (n, n)).astype(np.float32)
b_h = rng.uniform(0, 1, (n, n)).astype(np.float32)
a = vk.Array(gpu, data=a_h)
b = vk.Array(gpu, data=b_h)

for name, fn, ref in (("a + b", lambda: a + b, lambda: a_h + b_h),
                      ("a * b", lambda: a * b, lambda: a_h * b_h),
                      ("a.sum(axis=0)", lambda: a.sum(axis=0), lambda: a_h.sum(axis=0, dtype=np.float64))):
    c = fn()
    c.wait()                     # the job model of the reference: results are futures until waited for
    t0 = time.perf_counter()
    for _ in range(10):
        c = fn()
    c.wait()
    dt = (time.perf_counter() - t0) / 10
    np.testing.assert_allclose(np.asarray(c), ref(), rtol=2e-6)
    print(f"{name:16s} {dt * 1e3:8.3f} ms per call  (matches NumPy)")

row = vk.Array(gpu, data=b_h[0])
print("a + row ->", (a + row).shape, " a @ b[:, :8] ->", (a @ vk.Array(gpu, data=b_h[:, :8].copy())).shape)
a += 1.0
a *= b
print("in place: a = (a + 1) * b, a[0, :3] =", a[0, :3], "expected", ((a_h + 1) * b_h)[0, :3])
