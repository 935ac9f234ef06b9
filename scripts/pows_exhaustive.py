"""Exhaustive accuracy of the binomial-series a**s path (vkp_math.cuh pows_core, the same code the kernel runs,
compiled for the host) over EVERY positive normal float32, against float64 pow.  CPU only.
  python scripts/pows_exhaustive.py 2.7 [more exponents] > profiles/r02_pows_exhaustive.txt"""
import ctypes as C, os, subprocess, sys, tempfile, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "vkp_math.cuh"
extern "C" int p_plan(float s){ vkpm::PowsCoef c; return vkpm::pows_plan(s, c); }
extern "C" void p_pows_bits(unsigned first, long n, float s, float* z){
  vkpm::PowsCoef c; const int D = vkpm::pows_plan(s, c);
  vkpm::PowsHostTables t(s, c);
  for(long i=0;i<n;i++){ const unsigned u = first + (unsigned)i;
    z[i] = D == 6 ? vkpm::pows_core<6>(u, t) : D == 8 ? vkpm::pows_core<8>(u, t) : vkpm::pows_core<10>(u, t); }
}
'''
d = tempfile.mkdtemp()
open(os.path.join(d, "h.cpp"), "w").write(SRC)
subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "vulkpy_b200", "csrc"),
                "-o", os.path.join(d, "h.so"), os.path.join(d, "h.cpp")], check=True)
lib = C.CDLL(os.path.join(d, "h.so"))
lib.p_plan.argtypes = [C.c_float]
lib.p_pows_bits.argtypes = [C.c_uint, C.c_long, C.c_float, C.c_void_p]
CH = 1 << 24
for sv in [float(a) for a in sys.argv[1:]] or [2.7]:
    s = np.float32(sv)
    t0 = time.time()
    worst, nbad, ntot, nflush_bad = 0.0, 0, 0, 0
    z = np.empty(CH, np.float32)
    for first in range(0x00800000, 0x7f800000, CH):
        n = min(CH, 0x7f800000 - first)
        lib.p_pows_bits(first, n, s, z.ctypes.data)
        x = np.arange(first, first + n, dtype=np.uint32).view(np.float32).astype(np.float64)
        with np.errstate(all="ignore"):
            ex = np.power(x, np.float64(s))
            ok = (ex > 1.1754944e-38) & (ex < 3.4028234e38)
            r32 = ex[ok].astype(np.float32)
            u = np.abs(z[:n][ok].astype(np.float64) - ex[ok]) / np.spacing(np.abs(r32)).astype(np.float64)
        worst = max(worst, float(u.max()) if u.size else 0.0)
        nbad += int((u > 0.5001).sum())
        ntot += int(ok.sum())
        nflush_bad += int((z[:n][ex < 1.0e-38] != 0).sum()) + int((~np.isinf(z[:n][ex > 3.5e38])).sum())
    print(f"s = {float(s)!r}: degree {lib.p_plan(s)}, {ntot} inputs with a normal result (of {0x7f800000 - 0x00800000} positive normal "
          f"floats): max error {worst:.6f} ulp, {nbad} above 0.5001 ulp; {nflush_bad} wrong flushes / overflows;  {time.time() - t0:.0f} s",
          flush=True)
