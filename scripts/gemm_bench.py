"""8192^3 fp32 matmul timing (tcgen05 3xTF32 vs SIMT) with accuracy against float64 on a sample."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200._backend import Timer

gpu = vk.GPU(0)
dev = gpu.gpu
out = {}
for m in (2048, 4096, 8192):
    rng = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=1)
    a = rng.random(shape=(m, m)); a -= 0.5
    b = rng.random(shape=(m, m)); b -= 0.5
    for name, flags in (("tc", 2), ("simt", 1)):
        if name == "simt" and (m > 4096 or os.environ.get("GEMM_BENCH_NO_SIMT")):
            continue
        c = vk.Array(gpu, shape=(m, m))
        for _ in range(2):
            c.job = dev.gemm(False, False, m, m, m, a.buffer, b.buffer, c.buffer, None, flags)
        gpu.wait()
        t0, t1 = Timer(dev), Timer(dev)
        reps = 5
        t0.record()
        for _ in range(reps):
            c.job = dev.gemm(False, False, m, m, m, a.buffer, b.buffer, c.buffer, None, flags)
        t1.record()
        ms = t0.elapsed_ms(t1) / reps
        res = {"ms": round(ms, 3), "tflops_fp32": round(2 * m ** 3 / ms / 1e9, 2)}
        if name == "tc":
            res["tf32_pipe_tflops_issued"] = round(3 * 2 * m ** 3 / ms / 1e9, 1)
            t0.record()
            for _ in range(reps):
                c.job = dev.gemm(False, True, m, m, m, a.buffer, b.buffer, c.buffer, None, flags)
            t1.record()
            ms2 = t0.elapsed_ms(t1) / reps
            out[f"tc_tn_{m}"] = {"ms": round(ms2, 3), "tflops_fp32": round(2 * m ** 3 / ms2 / 1e9, 2),
                                 "tf32_pipe_tflops_issued": round(3 * 2 * m ** 3 / ms2 / 1e9, 1)}
            c.job = dev.gemm(False, False, m, m, m, a.buffer, b.buffer, c.buffer, None, flags)
            ah, bh = np.asarray(a)[:64].astype(np.float64), np.asarray(b).astype(np.float64)
            want = ah @ bh
            mag = np.abs(ah) @ np.abs(bh)
            got = np.asarray(c)[:64]
            res["max_err_over_sum_abs"] = float((np.abs(got - want) / mag).max())
            res["max_rel_err_vs_|C|rms"] = float(np.abs(got - want).max() / np.sqrt((want ** 2).mean()))
        out[f"{name}_{m}"] = res
print(json.dumps(out, indent=1))
