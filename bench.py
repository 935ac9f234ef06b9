#!/usr/bin/env python
"""
bench.py -- the vulkpy array hot path on B200: BASELINE.json configs[1]
("elementwise + broadcast arithmetic and sin/exp/pow/clamp on 2^28-element float32 arrays").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n 28]

One "step" = one pass of the 12-op list below over two resident 16384x16384 float32 arrays
(a in [0.5,2), b in [-2,2)) plus a row and a column vector.  metric = algorithmic bytes moved
per second, aggregated over all ranks (weak scaling: every rank owns its own 2^28-element
shard, no data-path collective -- SURVEY.md 8(e)).  Timing: CUDA events recorded on the stream
the kernels are launched on (vkp_timer_*), barrier + device sync on both sides, max over ranks.

Prints ONE JSON line.  Besides the contract keys it carries
  roofline         dominant kernel of the step (+ `worst`: the op furthest below the copy peak)
  roofline_matmul  the other half of BASELINE.json's metric: 8192^2 fp32 `@`, burst and sustained
  configs          C3 (axis reductions, 16384^2), C4 (gather, 2^26 indices), C5 (Xoshiro128pp 2^30
                   samples at 64 and 2^20 lanes; one MLP train step) -- each with its own clock record
  sharded          (N > 1) the exchange rows of SURVEY 8(e) on the PRODUCT's communicator
                   (vulkpy_b200.dist: peer-mailbox all-reduce, fused row-sharded tcgen05 matmul,
                   data-parallel MLP step, sharded generator), each parity-checked first
--impl reference times the CPU restatement of the reference's shaders (oracle/cpu_ref.c, OpenMP over
all host cores; the reference itself needs Vulkan and cannot run in this image) on the same workload.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "array_op_hbm_gbs"
UNIT = "GB/s"

# (name, algorithmic bytes per element) -- BASELINE.md section 4
OPS = [
    ("a+b", 12), ("a*b", 12), ("c+=b", 12), ("a*2.5", 8), ("a+row", 8), ("a*col", 8),
    ("sin(b)", 8), ("exp(b)", 8), ("a**b", 12), ("a**2.7", 8), ("clamp_ss", 8), ("clamp_vv", 16),
]
BYTES_PER_ELEM = sum(b for _, b in OPS)
# op -> kernel-name fragment in the committed ncu capture of this step (profiles/*ncu_elementwise*.csv)
NCU_KERNEL = {"a+b": "ew_kernel<2, FAdd>", "a*b": "ew_kernel<2, FMul>", "c+=b": "ew_kernel<2, FAdd>",
              "a*2.5": "ew_kernel<1, FScalar<FMul", "a+row": "bcast_vec_kernel<BAdd", "a*col": "bcast_vec_kernel<BMul",
              "sin(b)": "USin", "exp(b)": "TExp", "a**b": "<2, TPow", "a**2.7": "ew_pows_kernel", "clamp_ss": "CClampSS",
              "clamp_vv": "CClampVV"}


def ncu_traffic(op, n_elems):
    """DRAM bytes per launch of the op's kernel (dram__bytes_read.sum + dram__bytes_write.sum of the
    committed `ncu --set full` capture, which ran this same step on 2^28 elements), or None."""
    import csv, glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_elementwise*.csv")))
    if not files or n_elems != 1 << 28:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        h = rows[0]
        kn, rd, wr = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[rows[1][rd]]
        for r in rows[2:]:
            if NCU_KERNEL[op] in r[kn]:
                return int((float(r[rd]) + float(r[wr])) * unit), os.path.relpath(files[-1], ROOT)
    except Exception:
        pass
    return None, None


def op_list(a, b, row, col):
    """The step: returns the result of the last op (so that an end-to-end caller can read it)."""
    c = a + b
    yield "a+b", c
    d = a * b
    yield "a*b", d
    c += b
    yield "c+=b", c
    yield "a*2.5", a * 2.5
    yield "a+row", a + row
    yield "a*col", a * col
    yield "sin(b)", b.sin()
    yield "exp(b)", b.exp()
    yield "a**b", a ** b
    yield "a**2.7", a ** 2.7
    yield "clamp_ss", a.clamp(0.75, 1.5)
    yield "clamp_vv", c.clamp(b, a)


def run_step(a, b, row, col):
    last = None
    for _, last in op_list(a, b, row, col):
        pass
    return last


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md).  One process
    samples for the whole run; `window(t0, t1)` summarises the samples that arrived in a region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.samples = []       # (arrival time, sm, max, power, [reasons])

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                self.samples.append((time.time(), float(f[1]), float(f[2]), float(f[3]),
                                     [n for n, v in zip(self.NAMES, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue

    def stop(self):
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def window(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        rows = [s for s in self.samples if t0 <= s[0] <= t1 + 0.03]
        if not rows:     # region shorter than the sampling period: nearest sample
            rows = sorted(self.samples, key=lambda s: abs(s[0] - 0.5 * (t0 + t1)))[:1]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample"]}
        return {"sm_mhz": statistics.median(s[1] for s in rows), "sm_max_mhz": max(s[2] for s in rows),
                "power_w_max": max(s[3] for s in rows), "samples": len(rows),
                "reasons": sorted({r for s in rows for r in s[4]})}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        return {"hbm": float(j["hbm_gbs"]), "bf16": float(j["bf16_tflops"]),
                "bf16_sustained": float(j.get("bf16_tflops_sustained", j["bf16_tflops"])),
                "src": "MEASURED_PEAKS.json (driver-measured copy bandwidth / cuBLAS bf16 on this pool)"}
    except (OSError, KeyError, ValueError):
        return {"hbm": 6650.0, "bf16": 1590.0, "bf16_sustained": 1400.0,
                "src": "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}


# ------------------------------------------------------------------------------------- CPU arm
def cpu_reference_pass(L, ptr, a, b, row, col, c, out, shapes):
    """The same 12-op list through the C restatement of the shaders (oracle/cpu_ref.c)."""
    import ctypes as C
    n = a.size
    rows, cols = a.shape
    L.ref_binary(0, ptr(a), ptr(b), ptr(c), n)                       # a+b
    L.ref_binary(2, ptr(a), ptr(b), ptr(out), n)                     # a*b
    L.ref_binary(0, ptr(c), ptr(b), ptr(c), n)                       # c+=b
    L.ref_scalar(2, 0, ptr(a), C.c_float(2.5), ptr(out), n)          # a*2.5
    L.ref_broadcast_binary(0, ptr(a), ptr(row), ptr(out), ptr(shapes["row"]), n, cols, n, 2)
    L.ref_broadcast_binary(2, ptr(a), ptr(col), ptr(out), ptr(shapes["col"]), n, rows, n, 2)
    L.ref_unary(2, ptr(b), ptr(out), n)                              # sin
    L.ref_unary(14, ptr(b), ptr(out), n)                             # exp
    L.ref_binary(6, ptr(a), ptr(b), ptr(out), n)                     # a**b
    L.ref_scalar(6, 0, ptr(a), C.c_float(2.7), ptr(out), n)          # a**2.7
    L.ref_clamp_ss(ptr(a), C.c_float(0.75), C.c_float(1.5), ptr(out), n)
    L.ref_clamp_vv(ptr(c), ptr(b), ptr(a), ptr(out), n)


def cpu_arm(log2n_sample: int, steps: int, warmup: int):
    from oracle import cpu_ref
    L = cpu_ref.load()
    # all host threads (torchrun exports OMP_NUM_THREADS=1; the baseline is "every core the box has")
    L.ref_set_num_threads(len(os.sched_getaffinity(0)))
    rows = 1 << (log2n_sample // 2)
    cols = (1 << log2n_sample) // rows
    rs = np.random.default_rng(1234)
    a = rs.random((rows, cols), dtype=np.float32) * np.float32(1.5) + np.float32(0.5)
    b = rs.random((rows, cols), dtype=np.float32) * np.float32(4.0) - np.float32(2.0)
    row, col = b[0].copy(), b[:, 0].copy()
    c, out = np.empty_like(a), np.empty_like(a)
    shapes = {"row": np.array([rows, cols, 1, cols, rows, cols], np.uint32),
              "col": np.array([rows, cols, rows, 1, rows, cols], np.uint32)}
    for _ in range(warmup):
        cpu_reference_pass(L, cpu_ref.ptr, a, b, row, col, c, out, shapes)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(L, cpu_ref.ptr, a, b, row, col, c, out, shapes)
    dt = (time.perf_counter() - t0) / steps
    gbs = BYTES_PER_ELEM * a.size / dt / 1e9
    return gbs, dt * 1e3, L.ref_num_threads(), f"same 12-op list on {rows}x{cols} float32 (2^{log2n_sample} elements)"


def reference_main(args, rank, world):
    if rank != 0:
        return
    rows = 1 << (args.cpu_log2n // 2)
    cols = (1 << args.cpu_log2n) // rows
    gbs, ms, cores, sample = cpu_arm(args.cpu_log2n, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(rows, cols, sample=sample),
        "cpu_baseline": {"value": round(gbs, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference shaders (oracle/cpu_ref.c, OpenMP); the "
                                 "reference's own SPIR-V needs Vulkan+lavapipe, absent from this image"},
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(rows, cols, **extra):
    cfg = {"workload": f"C2: {len(OPS)} elementwise/broadcast/transcendental ops on "
                       f"{rows}x{cols} float32 per GPU ({BYTES_PER_ELEM} algorithmic B/elem)",
           "ops": [o for o, _ in OPS], "l2": "inputs (1 GiB each) are larger than the 126 MB L2",
           "memory": "cudaMalloc pool (buffers move to managed memory only when the host views them)"}
    cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------- GPU arm
def bind_near_gpu(local_rank):
    """Multi-rank runs: keep this rank's threads (and so its page-locked staging memory, which is
    placed on the allocating thread's NUMA node) on the CPUs next to its GPU's PCIe root."""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        cpus = set()
        for part in open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"bus": bus, "cpus": len(cpus), "numa_node": open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()}
    except Exception as e:
        return {"error": str(e)[:100]}
    return None


class Bench:
    """Shared helpers of the GPU arm: device events on the launch stream, max over ranks."""

    def __init__(self, vk, gpu, rank, world, local_rank, tdist, sampler, peaks):
        from vulkpy_b200._backend import Timer
        self.vk, self.gpu, self.dev = vk, gpu, gpu.gpu
        self.rank, self.world, self.local_rank, self.td = rank, world, local_rank, tdist
        self.Timer, self.sampler, self.peaks = Timer, sampler, peaks

    def barrier(self):
        self.gpu.wait()
        if self.td is not None:
            import torch
            self.td.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.td is None:
            return x
        import torch
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        self.td.all_reduce(t, op=self.td.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x: float) -> float:
        return -self.max_over_ranks(-x)

    def timed(self, fn, min_ms: float = 60.0, reps_cap: int = 400, warm: int = 3, collective: bool = False):
        """Median-free: `inner` back-to-back calls between one event pair (device time per call),
        inner chosen so that the region lasts >= min_ms.  collective=True: barrier first, max over ranks."""
        for _ in range(warm):
            r = fn()
            del r
        self.gpu.wait()
        t0, t1 = self.Timer(self.dev), self.Timer(self.dev)
        t0.record()
        r = fn()
        t1.record()
        est = max(t0.elapsed_ms(t1), 1e-3)
        del r
        inner = int(min(reps_cap, max(5, math.ceil(min_ms / est))))
        if collective:
            inner = int(self.max_over_ranks(float(inner)))
            self.barrier()
        l0 = self.dev.launch_count()
        t0.record()
        for _ in range(inner):
            r = fn()
        t1.record()
        ms = t0.elapsed_ms(t1) / inner
        launches = (self.dev.launch_count() - l0) // inner
        del r
        if collective:
            ms = self.max_over_ranks(ms)
        return ms, launches, inner


def hbm_row(B, name, fn, nbytes, note=None):
    ms, launches, inner = B.timed(fn)
    gbs = nbytes / ms / 1e6
    r = {"ms": round(ms, 4), "gbs": round(gbs, 1), "frac": round(gbs / B.peaks["hbm"], 4), "alg_bytes": int(nbytes),
         "launches": launches, "calls_timed": inner}
    if note:
        r["note"] = note
    return name, r


def config_c3(B, a, R):
    """C3: axis reductions on 16384^2 float32 (vkarray.py:1194-1276): 4*N_in + 4*N_out algorithmic bytes."""
    t0 = time.time()
    rows = {}
    B4 = 4 * R * R
    for op in ("sum", "maximum", "mean"):
        for axis, nout in ((0, R), (1, R), (None, 1)):
            k, v = hbm_row(B, f"{op}(axis={axis})", (lambda o, ax: (lambda: getattr(a, o)(axis=ax)))(op, axis), B4 + 4 * nout)
            rows[k] = v
    k, v = hbm_row(B, "sum(axis=1,rebroadcast)", lambda: a.sum(axis=1, rebroadcast=True), 2 * B4)
    rows[k] = v
    k, v = hbm_row(B, "maximum(axis=0,rebroadcast)", lambda: a.maximum(axis=0, rebroadcast=True), 2 * B4)
    rows[k] = v
    B.gpu.wait()
    return {"workload": f"sum/maximum/mean over axis 0 / 1 / None and rebroadcast on {R}x{R} float32 (input 1 GiB > L2)",
            "rows": rows, "clocks": B.sampler.window(t0, time.time()),
            "worst": min(rows, key=lambda k: rows[k]["frac"]), "bound": "hbm", "peak": B.peaks["hbm"]}


def config_c4(B, rng):
    """C4 gather half: 2^26 uint32 indices into an 8192^2 table (vkarray.py:1490-1521), 12 B per index."""
    vk, gpu = B.vk, B.gpu
    t0 = time.time()
    G, NI = 8192, 1 << 26
    table = rng.random(shape=(G, G))
    idx_h = np.random.default_rng(99).integers(0, G * G, NI, dtype=np.uint32)
    idx = vk.U32Array(gpu, data=idx_h)
    idx_sorted = vk.U32Array(gpu, data=np.sort(idx_h))
    del idx_h
    rows = {}
    k, v = hbm_row(B, "gather 2^26 random idx", lambda: table.gather(idx), 12 * NI,
                   note="DRAM-bound on fetch granularity: ncu counts 6.26 GB read from DRAM per launch = 93.4 B per 4-byte "
                        "payload (195.8 M sectors for 67.1 M elements; L2 hit 10 %: the 256 MiB table is 2x the L2), "
                        "profiles/r02_ncu_rows.md; cudaLimitMaxL2FetchGranularity 32/64/128 changes nothing "
                        "(profiles/r02_gather_gran.txt)")
    # DRAM bytes the launch really moves: measured 93.4 B per random element read + 4 B index read + 4 B write
    v["dram_gbs_ncu_traffic"] = round((93.4 + 8) * NI / v["ms"] / 1e6, 1)
    v["dram_frac_ncu_traffic"] = round(v["dram_gbs_ncu_traffic"] / B.peaks["hbm"], 4)
    rows[k] = v
    k, v = hbm_row(B, "gather 2^26 sorted idx", lambda: table.gather(idx_sorted), 12 * NI)
    rows[k] = v
    B.gpu.wait()
    return {"workload": "table 8192^2 float32 (256 MiB), 2^26 uint32 indices from np.random.default_rng(99), and a sorted copy",
            "rows": rows, "clocks": B.sampler.window(t0, time.time()), "bound": "hbm (random: DRAM fetch granularity; sorted: streaming)", "peak": B.peaks["hbm"]}


def config_c5(B):
    """C5: Xoshiro128pp random / randint / normal, 2^30 samples, default 64 lanes and 2^20 lanes
    (_vkarray.cc:697-717, random.py:60-124), 4 B per sample; one MLP train step (nn/models.py:55-78)."""
    vk, gpu = B.vk, B.gpu
    from vulkpy_b200 import nn
    t0 = time.time()
    P = 1 << 30
    rows = {}
    buf = vk.Array(gpu, shape=(P,))
    ubuf = vk.U32Array(gpu, shape=(P,))
    for size in (64, 1 << 20):
        g = vk.random.Xoshiro128pp(gpu, size=size, seed=7)
        for kind, target in (("random", buf), ("randint", ubuf), ("normal", buf)):
            k, v = hbm_row(B, f"{kind} 2^30 (size={size})",
                           (lambda gg, kk, tt: (lambda: getattr(gg, kk)(buffer=tt)))(g, kind, target), 4 * P)
            rows[k] = v
    del buf, ubuf
    prng_clocks = B.sampler.window(t0, time.time())
    t1 = time.time()
    Bsz, D, H, C = 8192, 1024, 1024, 16
    opt = nn.Adam(gpu, lr=1e-3)
    net = nn.Sequence([nn.Dense(gpu, D, H, w_opt=opt, b_opt=opt, w_init=nn.HeNormal(gpu, D, seed=1)), nn.ReLU(),
                       nn.Dense(gpu, H, C, w_opt=opt, b_opt=opt, w_init=nn.HeNormal(gpu, H, seed=2)), nn.Softmax()],
                      nn.CrossEntropyLoss())
    x = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=3).normal(shape=(Bsz, D))
    y = vk.random.Xoshiro128pp(gpu, seed=4).randrange(shape=(Bsz,), low=0, high=C).to_onehot(C)
    gpu.wait()
    ms, launches, inner = B.timed(lambda: net.train(x, y)[1], min_ms=150.0)
    flops = 6 * Bsz * (D * H + H * C)
    w0 = time.perf_counter()
    for _ in range(20):
        net.train(x, y)
    gpu.wait()
    wall = (time.perf_counter() - w0) / 20 * 1e3
    tf32_peak = B.peaks["bf16_sustained"] / 2
    mlp = {"workload": f"Sequence[Dense({D},{H}), ReLU, Dense({H},{C}), Softmax] + CrossEntropyLoss + Adam, batch {Bsz}",
           "ms": round(ms, 4), "launches": launches, "calls_timed": inner, "rows_per_s": round(Bsz / ms * 1e3, 1),
           "tflops_fp32": round(flops / ms / 1e9, 2), "wall_ms_per_step": round(wall, 3),
           "frac_of_tf32_pipe": round(3 * flops / ms / 1e9 / tf32_peak, 4),
           "peak": tf32_peak, "peak_note": "TF32 dense = bf16_tflops_sustained / 2; 3 tensor-core MMAs per useful fp32 product",
           "clocks": B.sampler.window(t1, time.time())}
    return {"workload": "Xoshiro128pp random / randint / normal, 2^30 samples (4 GiB written), size=64 (reference default) and 2^20",
            "rows": rows, "clocks": prng_clocks, "worst": min(rows, key=lambda k: rows[k]["frac"]),
            "bound": "hbm (writes); normal: instruction issue", "peak": B.peaks["hbm"], "mlp_step": mlp}


def matmul_block(B, rng, seconds: float):
    """8192^2 fp32 `@` (vkarray.py:585-605, shader/matmul.comp): burst (3 calls) and sustained."""
    vk, gpu, dev, Timer = B.vk, B.gpu, B.dev, B.Timer
    m = 8192
    ma = rng.random(shape=(m, m))
    mb = rng.random(shape=(m, m))
    for _ in range(2):
        mc = ma @ mb
    gpu.wait()
    flops = 2 * m ** 3
    t0, t1 = Timer(dev), Timer(dev)
    w0 = time.time()
    t0.record()
    for _ in range(3):
        mc = ma @ mb
    t1.record()
    burst_ms = t0.elapsed_ms(t1) / 3
    w1 = time.time()
    n = max(4, int(seconds * 1e3 / burst_ms))
    l0 = dev.launch_count()
    t0.record()
    for _ in range(n):
        mc = ma @ mb
    t1.record()
    sus_ms = t0.elapsed_ms(t1) / n
    launches = (dev.launch_count() - l0) // n
    w2 = time.time()
    del ma, mb, mc
    pk_b, pk_s = B.peaks["bf16"] / 2, B.peaks["bf16_sustained"] / 2
    tb, ts = flops / burst_ms / 1e9, flops / sus_ms / 1e9
    return {"bound": "tensor", "kernel": "gemm_tc_kernel<256,16,PRESPLIT> (tcgen05 kind::tf32, 3xTF32 split) + transpose/split pre-pass",
            "shape": [m, m, m], "unit": "TFLOP/s", "launches_per_call": launches,
            "burst": {"ms": round(burst_ms, 3), "achieved": round(tb, 2), "issued": round(3 * tb, 1), "peak": round(pk_b, 1),
                      "frac": round(3 * tb / pk_b, 4), "calls": 3, "clocks": B.sampler.window(w0, w1)},
            "sustained": {"ms": round(sus_ms, 3), "achieved": round(ts, 2), "issued": round(3 * ts, 1), "peak": round(pk_s, 1),
                          "frac": round(3 * ts / pk_s, 4), "calls": n, "seconds": round(n * sus_ms / 1e3, 2),
                          "clocks": B.sampler.window(w1, w2)},
            "frac_of_nominal_tf32_1100": round(3 * ts / 1100.0, 4),
            "peak_source": B.peaks["src"] + "; TF32 dense peak taken as bf16 / 2; `achieved` = useful fp32 FLOP/s, "
                           "`issued` = 3x that (three TF32 MMAs per product), frac = issued / peak"}


# ------------------------------------------------------------------------------------- sharded rows (N > 1)
def nvlink_bytes(index):
    """Sum of the NVLink data counters of GPU `index` in bytes (tx, rx), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True,
                             timeout=10).stdout
        tx = rx = 0
        for ln in out.splitlines():
            if "Data Tx" in ln:
                tx += int(ln.split(":")[-1].strip().split()[0])
            elif "Data Rx" in ln:
                rx += int(ln.split(":")[-1].strip().split()[0])
        return tx * 1024, rx * 1024
    except Exception:
        return None


def sharded_block(B):
    """SURVEY 8(e) exchange rows through vulkpy_b200.dist on the product's own communicator.  Every
    row is first asserted against the single-GPU / NumPy answer (a mismatch raises: non-zero exit),
    then timed on the device (barrier first, max over ranks)."""
    vk, gpu, rank, world = B.vk, B.gpu, B.rank, B.world
    from vulkpy_b200 import dist, nn
    F = np.float32
    g = dist.Group.from_env()
    out = {"world": world, "communicator": "vulkpy_b200.dist.NcclTransport (vkp_comm_*: NCCL via dlopen + CUDA-IPC peer mailbox)"}
    rs = np.random.default_rng(0)

    # ---- parity on moderate sizes, against NumPy -------------------------------------------------------
    R, Cc = 64 * world, 96
    a_f = rs.uniform(0.5, 2, (R, Cc)).astype(F)
    b_f = rs.uniform(0.5, 2, (R, Cc)).astype(F)
    a, b = g.shard(a_f), g.shard(b_f)
    p_f = rs.uniform(0.97, 1.03, (R, Cc)).astype(F)          # products over 64 * world rows stay inside float32
    pa = g.shard(p_f)
    for mode in (1, 0):                        # peer mailbox, then plain NCCL
        peer = g.t.peer_mode(mode)
        np.testing.assert_array_equal((a + b).to_numpy(), a_f + b_f)
        np.testing.assert_allclose(np.asarray(a.sum()), [a_f.astype(np.float64).sum()], rtol=2e-6)
        np.testing.assert_array_equal(np.asarray(a.maximum()), [a_f.max()])
        np.testing.assert_array_equal(np.asarray(a.minimum(axis=0)), a_f.min(axis=0))
        np.testing.assert_allclose(np.asarray(a.sum(axis=0)), a_f.astype(np.float64).sum(axis=0), rtol=2e-6)
        np.testing.assert_allclose(np.asarray(pa.prod(axis=0)), p_f.astype(np.float64).prod(axis=0), rtol=2e-5)
        np.testing.assert_allclose(a.sum(axis=1).to_numpy(), a_f.astype(np.float64).sum(axis=1), rtol=2e-6)
        np.testing.assert_allclose(np.asarray(a.mean()), [a_f.astype(np.float64).mean()], rtol=2e-6)
        np.testing.assert_allclose(a.maximum(axis=0, rebroadcast=True).to_numpy(),
                                   np.broadcast_to(a_f.max(axis=0, keepdims=True), a_f.shape))
        if mode == 1:
            out["peer_mailbox_active"] = peer
    g.t.peer_mode(1)
    Mg, Kg, Ng = 256 * world, 64 * world, 384
    A_f, B_f = rs.uniform(-1, 1, (Mg, Kg)).astype(F), rs.uniform(-1, 1, (Kg, Ng)).astype(F)
    for rep in range(3):      # three calls: both staging copies and the epoch flags get reused
        Cm = g.shard(A_f) @ g.shard(B_f)
        want = A_f.astype(np.float64) @ B_f
        mag = np.abs(A_f).astype(np.float64) @ np.abs(B_f)
        err = float((np.abs(Cm.to_numpy() - want) / mag).max())
        assert err < 6e-6, f"fused row-sharded matmul: error {err}"
        B_f = B_f + F(0.25)
    out["fused_matmul_used"] = not g.t._fused_broken
    out["fused_matmul_err_vs_float64"] = err
    shape = (16 * world, 256)
    ref = np.asarray(vk.random.Xoshiro128pp(gpu, seed=11).random(shape=shape))
    np.testing.assert_array_equal(g.random(vk.random.Xoshiro128pp(gpu, seed=11), shape, "random").to_numpy(), ref)
    refn = np.asarray(vk.random.Xoshiro128pp(gpu, seed=12).normal(shape=shape))
    np.testing.assert_array_equal(g.random(vk.random.Xoshiro128pp(gpu, seed=12), shape, "normal").to_numpy(), refn)

    def make_small():
        sgd = nn.SGD(0.05)
        return nn.Sequence([nn.Dense(gpu, 16, 32, w_opt=sgd, b_opt=sgd, w_init=nn.HeNormal(gpu, 16, seed=1)), nn.ReLU(),
                            nn.Dense(gpu, 32, 4, w_opt=sgd, b_opt=sgd, w_init=nn.HeNormal(gpu, 32, seed=2))],
                           nn.SoftmaxCrossEntropyLoss())
    Bg = 8 * world
    x_f = rs.normal(size=(Bg, 16)).astype(F)
    y_f = np.eye(4, dtype=F)[rs.integers(0, 4, Bg)]
    single = make_small()
    single.train(vk.Array(gpu, data=x_f), vk.Array(gpu, data=y_f))
    lo, hi = g.bounds(Bg)
    net = make_small()
    dist.DataParallel(net, g).train(vk.Array(gpu, data=x_f[lo:hi]), vk.Array(gpu, data=y_f[lo:hi]))
    for ls, ld in zip(single.L, net.L):
        if hasattr(ls, "w"):
            np.testing.assert_allclose(np.asarray(ld.w.value), np.asarray(ls.w.value), rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(np.asarray(ld.b.value), np.asarray(ls.b.value), rtol=2e-5, atol=1e-6)
    out["parity_moderate_sizes"] = "ok (element-wise, sum/max/min/prod/mean over None / 0 / 1, rebroadcast, fused matmul x3, PRNG bit-exact, DP step)"

    # ---- full size: properties + timings (weak scaling: 2^28 elements per GPU) ---------------------------
    rows_t = {}
    rowsN = 16384
    big = (rowsN * world, 16384)
    rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=5)
    x = g.random(rng, big, "random")
    y = g.random(rng, big, "random")
    n_total = big[0] * big[1]

    def gathered(v):               # every rank's host value through torch.distributed (independent path)
        objs = [None] * world
        B.td.all_gather_object(objs, v)
        return objs

    # property: sharded sum == float64 sum of the ranks' local sums; axis-0 likewise
    s_sh = float(np.asarray(x.sum())[0])
    s_loc = gathered(float(np.asarray(x.local.sum())[0]))
    assert abs(s_sh - sum(s_loc)) <= 2e-6 * abs(s_sh), (s_sh, s_loc)
    v_sh = np.asarray(x.sum(axis=0)).astype(np.float64)
    v_loc = np.sum(gathered(np.asarray(x.local.sum(axis=0)).astype(np.float64)), axis=0)
    np.testing.assert_allclose(v_sh, v_loc, rtol=2e-6)
    m_sh = np.asarray(x.maximum(axis=0))
    m_loc = np.max(gathered(np.asarray(x.local.maximum(axis=0))), axis=0)
    np.testing.assert_array_equal(m_sh, m_loc)
    chk = gathered(v_sh.tobytes())
    assert all(c == chk[0] for c in chk), "all-reduce results differ between ranks"
    out["parity_full_size"] = "ok (2^28 elements per GPU: sharded sum / sum(axis=0) / maximum(axis=0) == combination of the ranks' local results; results bit-identical on all ranks)"

    def row(name, fn, single_fn, unit_work, unit, weak=True, note=None, **kw):
        ms, launches, inner = B.timed(fn, collective=True, **kw)
        ms1 = ms1_max = None
        if single_fn:
            # the single-GPU baseline runs on every rank at once without a barrier; the FASTEST rank is the
            # baseline (the strictest choice: with 8 processes enqueueing from one host the slowest rank of a
            # launch-bound row can be 1.6x off), the slowest is reported beside it
            t1 = B.timed(single_fn, **kw)[0]
            ms1, ms1_max = B.min_over_ranks(t1), B.max_over_ranks(t1)
        r = {"ms": round(ms, 4), "launches": launches, "agg": round(unit_work / ms, 1), "unit": unit}
        if ms1:
            r["ms_one_gpu_same_local_work" if weak else "ms_one_gpu_whole_problem"] = round(ms1, 4)
            r["ms_one_gpu_slowest_rank"] = round(ms1_max, 4)
            r["x_over_one_gpu"] = round((world * ms1 if weak else ms1) / ms, 2)
        if note:
            r["note"] = note
        rows_t[name] = r
        return r

    GB = 1e6   # bytes / ms -> GB/s
    row("a+b (16384^2 per GPU, no exchange)", lambda: x + y, None, 12 * n_total / GB, "GB/s aggregate")
    row("sum() + all-reduce of 1 float", lambda: x.sum(), lambda: x.local.sum(), 4 * n_total / GB, "GB/s aggregate")
    row("sum(axis=0) + all-reduce of 16384 floats", lambda: x.sum(axis=0), lambda: x.local.sum(axis=0), 4 * n_total / GB, "GB/s aggregate")
    row("maximum(axis=0) + all-reduce", lambda: x.maximum(axis=0), lambda: x.local.maximum(axis=0), 4 * n_total / GB, "GB/s aggregate")
    row("sum(axis=1) (no exchange)", lambda: x.sum(axis=1), lambda: x.local.sum(axis=1), 4 * n_total / GB, "GB/s aggregate")
    g.t.peer_mode(0)
    row("sum(axis=0), exchange through ncclAllReduce (A/B)", lambda: x.sum(axis=0), None, 4 * n_total / GB, "GB/s aggregate")
    row("sum(), exchange through ncclAllReduce (A/B)", lambda: x.sum(), None, 4 * n_total / GB, "GB/s aggregate")
    g.t.peer_mode(1)
    del x, y

    # sharded generator: this rank's rows of ONE 2^30*world-sample stream (jump-ahead), bit-exact
    P = 1 << 30
    prng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=21)
    shp = (world * 1024, P // 1024)
    sh = g.random(prng, shp, "random")
    lo, hi = g.bounds(shp[0])
    probe = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=21)
    probe.rng.advance(lo * shp[1])
    head = np.asarray(probe.random(shape=(1 << 20,)))
    first = vk.U32Array(gpu, data=np.arange(1 << 20, dtype=np.uint32))
    np.testing.assert_array_equal(np.asarray(sh.local.gather(first)), head)
    del first
    buf = sh.local
    del sh

    def sharded_draw():
        r = g.random(prng, shp, "random")
        return r
    local_rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=22)
    row("Xoshiro128pp.random, 2^30 samples per GPU of one global stream", sharded_draw,
        lambda: local_rng.random(buffer=buf), 4.0 * P * world / GB, "GB/s aggregate",
        note="each rank jumps its lanes over the lower ranks' chunks (GF(2) matrices), draws, jumps over the rest")
    del buf

    # row-sharded matmul: fixed size and weak scaling
    M = 8192
    A = g.random(rng, (M, M), "random")
    Bm = g.random(rng, (M, M), "random")
    Cf = A @ Bm
    g.t._fused_broken, keep = True, g.t._fused_broken
    Cg = A @ Bm                                   # ncclAllGather(B) + the same GEMM: independent exchange path
    g.t._fused_broken = keep
    d = np.abs(Cf.local.to_host() - Cg.local.to_host()).max()
    assert d <= 0.03, f"fused vs all-gather matmul differ by {d}"     # 2 x 6e-6 x sum|a||b| (~2048)
    del Cf, Cg
    A1 = rng.random(shape=(M, M))
    B1 = rng.random(shape=(M, M))
    nv0 = nvlink_bytes(B.local_rank)
    r = row("matmul 8192^3 row-sharded (fixed size), peers' B pulled inside one tcgen05 GEMM", lambda: A @ Bm,
            lambda: A1 @ B1, 2 * M ** 3 / 1e9, "TFLOP/s aggregate", weak=False, min_ms=40.0)
    nv1 = nvlink_bytes(B.local_rank)
    pull_bytes = 4 * M * M * (world - 1) / world
    r["limiter"] = (f"NVLink pull rate: every rank pulls {pull_bytes / 1e6:.0f} MB of B through the kernel's four spare warps per CTA "
                    f"while it multiplies; {pull_bytes / r['ms'] / 1e6:.0f} GB/s sustained over the call.  The same staging + GEMM with "
                    "nothing crossing NVLink (VKP_COMM_NO_PULL=1, timing only) takes 0.51 ms on 8 GPUs against 0.73-0.77 ms with the "
                    "pulls (profiles/r02_mm_fused_probe_n8.txt); 8 -> 16 loads in flight per pulling thread bought 4 %, the 128x224 "
                    "tile (2 waves of 224 instead of 2 of 256 columns for 1024 x 8192 outputs) 5 % without the pulls and nothing "
                    f"with them.  Ideal = one GPU's {r.get('ms_one_gpu_whole_problem', 0):.3f} ms / {world}")
    if nv0 and nv1:
        r["nvlink_rx_bytes_per_call_counter"] = None     # the counters move with every rank's calls; reported raw below
        r["nvlink_counter_delta_bytes"] = {"tx": nv1[0] - nv0[0], "rx": nv1[1] - nv0[1]}
    r["nvlink_pull_bytes_per_call"] = int(4 * M * M * (world - 1) / world)
    r["nvlink_pull_gbs_lower_bound"] = round(4 * M * M * (world - 1) / world / r["ms"] / 1e6, 1)
    g.t._fused_broken = True
    row("matmul 8192^3 row-sharded, ncclAllGather(B) then GEMM (A/B)", lambda: A @ Bm, None, 2 * M ** 3 / 1e9,
        "TFLOP/s aggregate", weak=False, min_ms=40.0)
    g.t._fused_broken = False
    del A
    Aw = g.random(rng, (M * world, M), "random")
    row("matmul (8192*world) x 8192 x 8192 weak scaling, fused", lambda: Aw @ Bm, lambda: A1 @ B1,
        2.0 * world * M ** 3 / 1e9, "TFLOP/s aggregate", weak=True, min_ms=40.0)
    del Aw, Bm, A1, B1

    # data-parallel MLP step (config 5): 8192 rows per GPU, global batch 8192 * world
    def make_mlp():
        opt = lambda: nn.Adam(gpu, lr=1e-3)
        return nn.Sequence([nn.Dense(gpu, 1024, 1024, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 1024, seed=1)), nn.ReLU(),
                            nn.Dense(gpu, 1024, 16, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 1024, seed=2)), nn.Softmax()],
                           nn.CrossEntropyLoss())
    Bl = 8192
    xb = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=100 + rank).normal(shape=(Bl, 1024))
    yb = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=200 + rank).randrange(shape=(Bl,), low=0, high=16).to_onehot(16)
    mlp = make_mlp()
    dpm = dist.DataParallel(mlp, g)
    single_mlp = make_mlp()
    row("data-parallel MLP step, 8192 rows per GPU (gradient bucket through the peer mailbox)",
        lambda: dpm.train(xb, yb)[1], lambda: single_mlp.train(xb, yb)[1], Bl * world * 1e3, "rows/s aggregate", min_ms=100.0)
    # the exchange alone, back to back: what the gradient bucket costs on the device when no rank is late
    grads = [p.grad for p in dpm.parameters()]
    bucket_bytes = sum(4 * gr.buffer.size() for gr in grads)
    r = row("gradient bucket all-reduce alone (peer mailbox, two-shot), back to back",
            lambda: g.t.allreduce_many(grads, "sum", 1.0)[0], None, bucket_bytes / GB, "GB/s of bucket per rank", min_ms=20.0)
    r["bucket_bytes"] = bucket_bytes
    dp_row = rows_t["data-parallel MLP step, 8192 rows per GPU (gradient bucket through the peer mailbox)"]
    dp_row["limiter"] = (f"step = local step {dp_row.get('ms_one_gpu_same_local_work')} ms + exchange; the exchange alone takes "
                         f"{r['ms']} ms back to back, the rest of the difference is rank skew: the host enqueues "
                         f"{dp_row['launches']} launches per step in about the time the device runs them, so a late "
                         "rank stalls all of them at the flag wait")
    # property: replicas stay identical (same reduced gradients, same optimizer step on every rank)
    sig = gathered(np.asarray(mlp.L[0].w.value.sum()).tobytes() + np.asarray(mlp.L[2].w.value.sum()).tobytes())
    assert all(s == sig[0] for s in sig), "data-parallel replicas diverged"
    g.t.peer_mode(0)
    row("data-parallel MLP step, gradient bucket through grouped ncclAllReduce (A/B)", lambda: dpm.train(xb, yb)[1],
        None, Bl * world * 1e3, "rows/s aggregate", min_ms=100.0)
    g.t.peer_mode(1)
    out["rows"] = rows_t
    out["target"] = ">= 7x on 8 GPUs for the comm-free and one-exchange rows (north_star); x_over_one_gpu is measured in this run"
    B.gpu.wait()
    g.t.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="elements per GPU = 2^log2n")
    ap.add_argument("--cpu-log2n", type=int, default=28, help="CPU arm: 2^cpu_log2n elements (same workload as the GPU arm)")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--matmul-seconds", type=float, default=2.0, help="length of the sustained matmul loop")
    ap.add_argument("--no-matmul", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3/C4/C5 rows")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the exchange rows")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_main(args, rank, world)
        return

    args.warmup = max(args.warmup, 3)
    dist = None
    binding = None
    if world > 1:
        if "NCCL_DEBUG" not in os.environ:               # the rank count of BOTH communicators (torch's barrier
            os.environ["NCCL_DEBUG"] = "INFO"            # group and the product's) shows in NCCL's own log;
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")        # stdout stays the one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        binding = bind_near_gpu(local_rank)
    import vulkpy_b200 as vk
    gpu = vk.GPU(local_rank)
    dev = gpu.gpu
    from vulkpy_b200._backend import Timer
    peaks = measured_peaks()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    B = Bench(vk, gpu, rank, world, local_rank, dist, sampler, peaks)

    rows = 1 << (args.log2n // 2)
    cols = (1 << args.log2n) // rows
    n = rows * cols
    rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=1234 + rank)
    a = rng.random(shape=(rows, cols))
    a *= 1.5
    a += 0.5                                     # [0.5, 2)
    b = rng.random(shape=(rows, cols))
    b *= 4.0
    b -= 2.0                                     # [-2, 2)
    row = rng.random(shape=(cols,))
    col = rng.random(shape=(rows, 1))
    gpu.wait()
    barrier = B.barrier

    for _ in range(args.warmup):
        run_step(a, b, row, col)
    barrier()

    w0 = time.time()
    t0, t1 = Timer(dev), Timer(dev)
    launches0 = dev.launch_count()
    t0.record()
    for _ in range(args.steps):
        run_step(a, b, row, col)
    t1.record()
    ms_total = t0.elapsed_ms(t1)
    launches = dev.launch_count() - launches0
    barrier()
    clocks = sampler.window(w0, time.time()) if rank == 0 else None

    ms_step = B.max_over_ranks(ms_total / args.steps)
    if dist is not None:
        import torch
        lt = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = world * BYTES_PER_ELEM * n / (ms_step * 1e-3) / 1e9

    # ---- per-op device times (outside the headline region) -> roofline of the dominant kernel
    per_op = {}
    reps = 5
    for _ in range(reps):
        ta = Timer(dev)
        ta.record()
        prev = ta
        for name, _res in op_list(a, b, row, col):
            tb = Timer(dev)
            tb.record()
            per_op.setdefault(name, []).append((prev, tb))
            prev = tb
        gpu.wait()
    op_ms = {k: statistics.median(p.elapsed_ms(q) for p, q in v) for k, v in per_op.items()}
    op_gbs = {name: bpe * n / (op_ms[name] * 1e-3) / 1e9 for name, bpe in OPS}
    # dominant kernel = largest share of the step over all launches of that kernel (a+b and c+=b are
    # the same ew_kernel<2,FAdd>): stable across runs, unlike "longest single launch"
    share = {}
    for name, _ in OPS:
        share[NCU_KERNEL[name]] = share.get(NCU_KERNEL[name], 0.0) + op_ms[name]
    dom_kernel = max(share, key=share.get)
    dominant = [nm for nm, _ in OPS if NCU_KERNEL[nm] == dom_kernel][0]
    worst = min(op_gbs, key=op_gbs.get)
    peak, peak_src = peaks["hbm"], peaks["src"]
    traffic, traffic_src = ncu_traffic(dominant, n)
    wtraffic, _ = ncu_traffic(worst, n)
    total_ms = sum(op_ms.values())
    roofline = {"bound": "hbm", "kernel": dominant, "kernel_symbol": dom_kernel, "achieved": round(op_gbs[dominant], 1),
                "peak": peak, "unit": "GB/s", "frac": round(op_gbs[dominant] / peak, 4), "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes": dict(OPS)[dominant] * n,
                "peak_source": peak_src, "share_of_step": round(share[dom_kernel] / total_ms, 4),
                "worst": {"kernel": worst, "kernel_symbol": NCU_KERNEL[worst], "achieved": round(op_gbs[worst], 1),
                          "frac": round(op_gbs[worst] / peak, 4), "share_of_step": round(op_ms[worst] / total_ms, 4),
                          "algorithmic_bytes": dict(OPS)[worst] * n, "traffic": wtraffic,
                          "bound": "instruction issue + FP64 pipe (binary64 log2 / exp2 for the 0.5001-ulp contract), not HBM"},
                "per_op_gbs": {k: round(v, 1) for k, v in op_gbs.items()},
                "per_op_frac": {k: round(v / peak, 4) for k, v in op_gbs.items()}}

    # ---- end to end: host buffers in, result out, through the public API
    a_h = vk.pinned_empty((rows, cols))
    b_h = vk.pinned_empty((rows, cols))
    out_h = vk.pinned_empty((rows, cols))
    a.to_host(a_h)
    b.to_host(b_h)
    row_h, col_h = np.asarray(row).copy(), np.asarray(col).copy()

    out2_h = vk.pinned_empty((rows, cols))
    outs = (out_h, out2_h)

    def e2e_upload():
        # the two large operands ride the host-to-device copy engine; the two vectors are pageable
        return (vk.Array.from_host(gpu, a_h), vk.Array.from_host(gpu, b_h),
                vk.Array(gpu, data=row_h), vk.Array(gpu, data=col_h))

    def e2e_run(steps):
        """Every step uploads its four inputs from host memory and downloads its result.  The loop is
        software-pipelined: the uploads of step i+1 are enqueued before the host waits for the
        result of step i, so H2D(i+1), the kernels of step i and D2H(i) overlap (PCIe is full duplex)."""
        nxt = e2e_upload()
        for i in range(steps):
            cur, nxt = nxt, None
            res = run_step(*cur)
            res.to_host(outs[i % 2], wait=False)
            if i + 1 < steps:
                nxt = e2e_upload()
            res.wait()

    e2e_run(2)
    barrier()
    w0 = time.perf_counter()
    e2e_run(args.e2e_steps)
    gpu.wait()
    e2e_ms = B.max_over_ranks((time.perf_counter() - w0) / args.e2e_steps * 1e3)
    e2e = {"value": round(world * BYTES_PER_ELEM * n / (e2e_ms * 1e-3) / 1e9, 2), "unit": UNIT,
           "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes + row_h.nbytes + col_h.nbytes),
           "d2h_bytes_per_step": int(out_h.nbytes), "ms_per_step": round(e2e_ms, 2),
           "path": "Array.from_host(pinned) x2 + Array(data=) x2 -> 12 ops -> Array.to_host(pinned, wait=False); "
                   "steps software-pipelined over the two copy engines"}
    if True:
        # the host bound beside it: the bare pinned transfers of one step (2 GiB in, 1 GiB out, both copy
        # engines at once) on all ranks together, no kernels
        def copies_only(steps):
            for i in range(steps):
                ua, ub = vk.Array.from_host(gpu, a_h), vk.Array.from_host(gpu, b_h)
                a.to_host(outs[i % 2], wait=False)
                ua.wait(); ub.wait(); a.wait()
        copies_only(1)
        barrier()
        w0 = time.perf_counter()
        copies_only(4)
        gpu.wait()
        cp_ms = B.max_over_ranks((time.perf_counter() - w0) / 4 * 1e3)
        e2e["host_copy_bound"] = {"ms_per_step": round(cp_ms, 2), "note": "bare pinned H2D 2 GiB + D2H 1 GiB per rank, all ranks at once, no kernels: "
                                  "what the box's PCIe / host memory gives", "h2d_gbs_per_rank": round((a_h.nbytes + b_h.nbytes) / cp_ms / 1e6, 1),
                                  "e2e_over_bound": round(cp_ms / e2e_ms, 3), "cpu_binding": binding}
    del a_h, b_h, out_h, out2_h, outs

    # ---- the other half of BASELINE.json's metric: 8192^2 fp32 matmul
    matmul = None
    if not args.no_matmul:
        try:
            matmul = matmul_block(B, rng, args.matmul_seconds if world == 1 else 0.5)
        except Exception as e:  # the bench line must still be printed
            matmul = {"error": str(e)[:200]}

    # ---- C3 / C4 / C5 (single-GPU rows; N > 1 runs the exchange rows instead)
    configs = None
    if world == 1 and not args.no_configs and args.log2n == 28:
        configs = {}
        for key, fn in (("C3", lambda: config_c3(B, a, rows)), ("C4", lambda: config_c4(B, rng)), ("C5", lambda: config_c5(B))):
            try:
                configs[key] = fn()
            except Exception as e:
                configs[key] = {"error": str(e)[:300]}
    del a, b, row, col

    sharded = None
    if world > 1 and not args.no_sharded:
        sharded = sharded_block(B)      # parity failures raise: non-zero exit, no bench line

    if rank == 0:
        sampler.stop()
        cpu_baseline = None
        if world == 1:
            dev.trim()
            cpu_gbs, cpu_ms, cores, sample = cpu_arm(args.cpu_log2n, 3, 1)
            cpu_baseline = {"value": round(cpu_gbs, 3), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": sample + ", 3 passes after 1 warm-up", "ms_per_step": round(cpu_ms, 1)}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(rows, cols),
            "frac_of_peak": round(value / world / peak, 4), "frac_of_nominal_8TBs": round(value / world / 8000.0, 4),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "roofline_matmul": matmul, "configs": configs, "sharded": sharded,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
