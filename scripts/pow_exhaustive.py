"""Exhaustive accuracy of the GENERAL pow path (vkp_math.cuh pow_fast = pow_core with the kernels' fallback) for fixed
exponents over EVERY positive normal float32 base, against float64.  CPU only."""
import ctypes as C, os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "vkp_math.cuh"
extern "C" void p_pow_bits(unsigned first, long n, float y, float* z){ vkpm::HostTables t;
  for(long i=0;i<n;i++) z[i] = vkpm::pow_fast(vkpm::bits2f(first + (unsigned)i), y, t); }
'''
d = tempfile.mkdtemp()
open(os.path.join(d, "h.cpp"), "w").write(SRC)
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "vulkpy_b200", "csrc"),
                "-o", os.path.join(d, "h.so"), os.path.join(d, "h.cpp")], check=True)
lib = C.CDLL(os.path.join(d, "h.so"))
lib.p_pow_bits.argtypes = [C.c_uint, C.c_long, C.c_float, C.c_void_p]
CH = 1 << 24
for yv in [float(a) for a in sys.argv[1:]] or [2.7]:
    y = np.float32(yv)
    t0 = time.time()
    worst, nbad, ntot = 0.0, 0, 0
    z = np.empty(CH, np.float32)
    for first in range(0x00800000, 0x7f800000, CH):
        n = min(CH, 0x7f800000 - first)
        lib.p_pow_bits(first, n, y, z.ctypes.data)
        x = np.arange(first, first + n, dtype=np.uint32).view(np.float32).astype(np.float64)
        with np.errstate(all="ignore"):
            ex = np.power(x, np.float64(y))
            ok = (ex > 1.1754944e-38) & (ex < 3.4028234e38)
            r32 = ex[ok].astype(np.float32)
            u = np.abs(z[:n][ok].astype(np.float64) - ex[ok]) / np.spacing(np.abs(r32)).astype(np.float64)
        worst = max(worst, float(u.max()) if u.size else 0.0)
        nbad += int((u > 0.5001).sum())
        ntot += int(ok.sum())
    print(f"pow(x, {float(y)!r}) general path: {ntot} bases with a normal result: max error {worst:.6f} ulp, {nbad} above 0.5001 ulp;  {time.time() - t0:.0f} s", flush=True)
