"""Exhaustive accuracy of the table-driven exp / exp2 / log / log2 (vkp_math.cuh *_fast: the kernels' code compiled for
the host, special inputs routed to the careful routines exactly as the kernels do) over EVERY float32 whose result is a
normal float, against float64.  CPU only.   python scripts/math_exhaustive.py >> profiles/r02_pows_exhaustive.txt"""
import ctypes as C, os, subprocess, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "vkp_math.cuh"
#define F(NAME, CALL) extern "C" void NAME(unsigned first, long n, float* z){ vkpm::HostTables t; \
  for(long i=0;i<n;i++){ const float x = vkpm::bits2f(first + (unsigned)i); z[i] = CALL; } }
F(e_exp, vkpm::exp_fast(x, t)) F(e_exp2, vkpm::exp2_fast(x, t)) F(e_log, vkpm::log_fast(x, t)) F(e_log2, vkpm::log2_fast(x, t))
'''
d = tempfile.mkdtemp()
open(os.path.join(d, "h.cpp"), "w").write(SRC)
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "vulkpy_b200", "csrc"),
                "-o", os.path.join(d, "h.so"), os.path.join(d, "h.cpp")], check=True)
lib = C.CDLL(os.path.join(d, "h.so"))
CH = 1 << 24
RANGES = {"exp": [(0x00000000, 0x7f800000), (0x80000000, 0xff800000)], "exp2": [(0x00000000, 0x7f800000), (0x80000000, 0xff800000)],
          "log": [(0x00800000, 0x7f800000)], "log2": [(0x00800000, 0x7f800000)]}
REF = {"exp": np.exp, "exp2": np.exp2, "log": np.log, "log2": np.log2}
for name, ranges in RANGES.items():
    fn = getattr(lib, "e_" + name)
    fn.argtypes = [C.c_uint, C.c_long, C.c_void_p]
    t0 = time.time()
    worst, nbad, ntot = 0.0, 0, 0
    z = np.empty(CH, np.float32)
    for lo, hi in ranges:
        for first in range(lo, hi, CH):
            n = min(CH, hi - first)
            fn(first, n, z.ctypes.data)
            x = np.arange(first, first + n, dtype=np.uint32).view(np.float32).astype(np.float64)
            with np.errstate(all="ignore"):
                ex = REF[name](x)
                a = np.abs(ex)
                ok = (a > 1.1754944e-38) & (a < 3.4028234e38)
                r32 = ex[ok].astype(np.float32)
                u = np.abs(z[:n][ok].astype(np.float64) - ex[ok]) / np.spacing(np.abs(r32)).astype(np.float64)
            worst = max(worst, float(u.max()) if u.size else 0.0)
            nbad += int((u > 0.5001).sum())
            ntot += int(ok.sum())
    print(f"{name}: {ntot} inputs with a normal non-zero result: max error {worst:.6f} ulp, {nbad} above 0.5001 ulp;  {time.time() - t0:.0f} s", flush=True)
