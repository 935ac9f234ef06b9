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

Prints ONE JSON line.  --impl reference times the CPU restatement of the reference's shaders
(oracle/cpu_ref.c, OpenMP over all host cores; the reference itself needs Vulkan and cannot run
in this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
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
              "sin(b)": "USin", "exp(b)": "TExp", "a**b": "TPow>", "a**2.7": "TPowScalar", "clamp_ss": "CClampSS",
              "clamp_vv": "CClampVV"}


def ncu_traffic(op, n_elems):
    """DRAM bytes per launch of the op's kernel (dram__bytes_read.sum + dram__bytes_write.sum of the
    committed `ncu --set full` capture, which ran this same step on 2^28 elements), or None."""
    import csv, glob
    root = os.path.dirname(os.path.abspath(__file__))
    files = sorted(glob.glob(os.path.join(root, "profiles", "*ncu_elementwise*.csv")))
    if not files or n_elems != 1 << 28:
        return None, None
    try:
        rows = list(csv.reader(open(files[-1])))
        h = rows[0]
        kn, rd, wr = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[rows[1][rd]]
        for r in rows[2:]:
            if NCU_KERNEL[op] in r[kn]:
                return int((float(r[rd]) + float(r[wr])) * unit), os.path.relpath(files[-1], root)
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
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


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
    a = rs.uniform(0.5, 2.0, (rows, cols)).astype(np.float32)
    b = rs.uniform(-2.0, 2.0, (rows, cols)).astype(np.float32)
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
    gbs, ms, cores, sample = cpu_arm(args.cpu_log2n, max(1, args.steps), max(1, min(args.warmup, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: 12 elementwise/broadcast/transcendental ops, float32", "sample": sample},
        "cpu_baseline": {"value": round(gbs, 3), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "C restatement of the reference shaders (oracle/cpu_ref.c, OpenMP); the "
                                 "reference's own SPIR-V needs Vulkan+lavapipe, absent from this image"},
        "e2e": {"value": round(gbs, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


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
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=28, help="elements per GPU = 2^log2n")
    ap.add_argument("--cpu-log2n", type=int, default=26, help="CPU sample size = 2^cpu_log2n elements")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-matmul", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        reference_main(args, rank, world)
        return

    args.warmup = max(args.warmup, 3)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        bind_near_gpu(local_rank)
    import vulkpy_b200 as vk
    gpu = vk.GPU(local_rank)
    dev = gpu.gpu
    from vulkpy_b200._backend import Timer

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

    def barrier():
        gpu.wait()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        run_step(a, b, row, col)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0, t1 = Timer(dev), Timer(dev)
    launches0 = dev.launch_count()
    t0.record()
    for _ in range(args.steps):
        run_step(a, b, row, col)
    t1.record()
    ms_total = t0.elapsed_ms(t1)
    launches = dev.launch_count() - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    if dist is not None:
        import torch
        t = torch.tensor([ms_step], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
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
    dominant = max(op_ms, key=op_ms.get)
    peak, peak_src = measured_peak()
    traffic, traffic_src = ncu_traffic(dominant, n)
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": round(op_gbs[dominant], 1), "peak": peak,
                "unit": "GB/s", "frac": round(op_gbs[dominant] / peak, 4), "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes": dict(OPS)[dominant] * n,
                "peak_source": peak_src, "share_of_step": round(op_ms[dominant] / sum(op_ms.values()), 4),
                "per_op_gbs": {k: round(v, 1) for k, v in op_gbs.items()},
                "per_op_frac": {k: round(v / peak, 4) for k, v in op_gbs.items()}}

    # ---- end to end: host buffers in, result out, through the public API
    e2e = None
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
    e2e_ms = (time.perf_counter() - w0) / args.e2e_steps * 1e3
    if dist is not None:
        import torch
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e = {"value": round(world * BYTES_PER_ELEM * n / (e2e_ms * 1e-3) / 1e9, 2), "unit": UNIT,
           "h2d_bytes_per_step": int(a_h.nbytes + b_h.nbytes + row_h.nbytes + col_h.nbytes),
           "d2h_bytes_per_step": int(out_h.nbytes), "ms_per_step": round(e2e_ms, 2),
           "path": "Array.from_host(pinned) x2 + Array(data=) x2 -> 12 ops -> Array.to_host(pinned, wait=False); "
                   "steps software-pipelined over the two copy engines"}
    del a_h, b_h, out_h, out2_h, outs

    # ---- the other half of BASELINE.json's metric: 8192^2 fp32 matmul (reported, not in the step)
    matmul = None
    if not args.no_matmul:
        try:
            m = 8192
            ma = rng.random(shape=(m, m))
            mb = rng.random(shape=(m, m))
            for _ in range(2):
                mc = ma @ mb
            gpu.wait()
            tm0, tm1 = Timer(dev), Timer(dev)
            tm0.record()
            for _ in range(3):
                mc = ma @ mb
            tm1.record()
            mm_ms = tm0.elapsed_ms(tm1) / 3
            matmul = {"shape": [m, m, m], "ms": round(mm_ms, 3), "tflops_fp32": round(2 * m ** 3 / (mm_ms * 1e-3) / 1e12, 2)}
            del ma, mb, mc
        except Exception as e:  # the bench line must still be printed
            matmul = {"error": str(e)[:200]}

    if rank == 0:
        cpu_gbs, cpu_ms, cores, sample = (None, None, None, None)
        cpu_baseline = None
        if world == 1:
            cpu_gbs, cpu_ms, cores, sample = cpu_arm(args.cpu_log2n, 3, 1)
            cpu_baseline = {"value": round(cpu_gbs, 3), "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": sample + ", 3 passes after 1 warm-up"}
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: {len(OPS)} elementwise/broadcast/transcendental ops on "
                                   f"{rows}x{cols} float32 per GPU ({BYTES_PER_ELEM} algorithmic B/elem)",
                       "ops": [o for o, _ in OPS], "l2": "inputs (1 GiB each) are larger than the 126 MB L2",
                       "memory": "cudaMalloc pool (buffers move to managed memory only when the host views them)"},
            "frac_of_peak": round(value / world / peak, 4),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "cpu_baseline": cpu_baseline, "matmul_8192": matmul,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
