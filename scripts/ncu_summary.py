"""Summarise ncu outputs into small tracked files under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python scripts/ncu_summary.py full gpurun_out/prof_full.ncu-rep profiles/r01_ncu_full.csv
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_xu.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg",
    "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gs, bs = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        key = (r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", ""), r[gs], r[bs])
        agg.setdefault(key, []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)\n\n")
        f.write("| share | launches | avg us | grid | block | kernel |\n|---:|---:|---:|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {sum(v) / tot * 100:.2f}% | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {k[1]} | {k[2]} | `{k[0][:110]}` |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([m for m, _ in idx])
        w.writerow([units[i] for _, i in idx])
        for r in data:
            w.writerow([r[i].replace("<unnamed>::", "")[:120] if m == "Kernel Name" else r[i] for m, i in idx])
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
