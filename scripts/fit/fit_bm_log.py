"""Fit the float32 polynomial behind vkpm::bm_log (vkp_math.cuh): log1p(f) = f + f^2 Q(f) on
f in [-1/3, 1/3], Q of degree 8, then report the worst error of the whole float32 evaluation over
ALL 2^23 inputs Box-Muller can produce (om = k 2^-23, k = 1..2^23)."""
import numpy as np
from numpy.polynomial import chebyshev as C

F = np.float32
lo, hi = -1.0 / 3, 1.0 / 3
x = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * (hi - lo) / 2 + (hi + lo) / 2
target = (np.log1p(x) - x) / (x * x)
w = np.ones_like(x)
for it in range(30):                       # Lawson-style reweighting towards minimax of the abs error of log1p
    A = np.vander(x, 9, increasing=True)
    wt = w * x * x                         # error in log1p = f^2 * error in Q
    coef, *_ = np.linalg.lstsq(A * wt[:, None], target * wt, rcond=None)
    err = np.abs((A @ coef - target) * x * x)
    w = w * (1 + 4 * err / err.max())
    w /= w.max()
print("max abs err of log1p approx (float64 coefs):", err.max())
c32 = coef.astype(F)
print("coefs:", ", ".join(f"{float(c):.9g}f" for c in c32))


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


k = np.arange(1, (1 << 23) + 1, dtype=np.int64)
om = (k.astype(np.float64) * 2.0 ** -23).astype(F)
bits = om.view(np.uint32).astype(np.int64)
ib = bits - 0x3f2aaaab
e = ib >> 23
m = (bits - (e << 23)).astype(np.uint32).view(F)
f = m - F(1)
q = np.full_like(f, c32[8])
for c in c32[7::-1]:
    q = fma(q, f, np.full_like(f, c))
t = (f * q).astype(F)
l1p = fma(f, t, f)
res = fma(e.astype(F), np.full_like(f, F(0.693147180559945)), l1p)
ref = np.log(om.astype(np.float64))
ulp = np.abs(res.astype(np.float64) - ref) / np.maximum(np.spacing(np.abs(ref).astype(F)).astype(np.float64), 1e-300)
ulp[ref == 0] = np.abs(res[ref == 0])
print("max ulp error over all 2^23 inputs:", ulp.max(), "at k =", k[ulp.argmax()])
