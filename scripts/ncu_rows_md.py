"""Long-format `ncu --metrics ... --csv --log-file X.csv` -> one markdown table (one row per kernel/grid, last launch).

  python scripts/ncu_rows_md.py gpurun_out/r02_ncu_rows_v2.csv profiles/r02_ncu_rows_v2.md "title" [regex: kernels to list launch by launch]
"""
import collections
import csv
import sys

SHORT = {
    "gpu__time_duration.sum": "us", "dram__bytes_read.sum": "DRAM rd MB", "dram__bytes_write.sum": "DRAM wr MB",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram %", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm %",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps act %", "smsp__inst_executed.sum": "warp inst M",
    "launch__registers_per_thread": "regs", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue act %",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64 pipe %",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor pipe %",
    "sm__inst_executed_pipe_fp64.sum": "fp64 inst M", "sm__inst_executed_pipe_fma.sum": "fma inst M",
    "sm__inst_executed_pipe_alu.sum": "alu inst M", "sm__inst_executed_pipe_xu.sum": "xu inst M",
    "sm__inst_executed_pipe_lsu.sum": "lsu inst M", "lts__t_sector_hit_rate.pct": "L2 hit %",
    "dram__sectors_read.sum": "DRAM rd sectors M", "lts__t_sectors_srcunit_tex_op_read.sum": "L2 rd sectors (tex) M",
    "lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum": "L2 rd hit sectors M",
    "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum": "L2 rd miss sectors M",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "L1 ld sectors M",
}
SCALE = {"DRAM rd MB": 1e-6, "DRAM wr MB": 1e-6, "warp inst M": 1e-6, "fp64 inst M": 1e-6, "fma inst M": 1e-6,
         "alu inst M": 1e-6, "xu inst M": 1e-6, "lsu inst M": 1e-6, "DRAM rd sectors M": 1e-6,
         "L2 rd sectors (tex) M": 1e-6, "L2 rd hit sectors M": 1e-6, "L2 rd miss sectors M": 1e-6, "L1 ld sectors M": 1e-6}


def main(src, dst, title, every=None):
    import re
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    c = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        name = r[c["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        key = (name, r[c["Grid Size"]])
        if every and re.search(every, name):
            key = (name + " #" + r[c["ID"]], r[c["Grid Size"]])      # list every launch of these kernels
        d = per.setdefault(key, {})
        if d.get("_id") != r[c["ID"]]:
            d.clear()
            d["_id"] = r[c["ID"]]
            d["_n"] = d.get("_n", 0) + 1
        v = float(r[c["Metric Value"]].replace(",", "") or 0)
        m, u = r[c["Metric Name"]], r[c["Metric Unit"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        elif u in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d[SHORT.get(m, m)] = v
    cols = []
    for d in per.values():
        for k in d:
            if not k.startswith("_") and k not in cols:
                cols.append(k)
    extra = "DRAM rd MB" in cols and "us" in cols
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics ... --clock-control none` (source: `{src}`); last launch of each kernel/grid; "
                "durations are serialised and cold-clock: compare ratios and shares, not absolutes\n\n")
        f.write("| kernel | grid | " + " | ".join(cols) + (" | DRAM GB/s |" if extra else " |") + "\n")
        f.write("|---|---|" + "---:|" * (len(cols) + (1 if extra else 0)) + "\n")
        for (name, grid), d in per.items():
            cells = []
            for k in cols:
                v = d.get(k)
                cells.append("" if v is None else (f"{v * SCALE.get(k, 1):.1f}" if k != "regs" else f"{int(v)}"))
            if extra:
                gbs = (d.get("DRAM rd MB", 0) + d.get("DRAM wr MB", 0)) / (d["us"] * 1e-6) / 1e9 if d.get("us") else 0
                cells.append(f"{gbs:.0f}")
            f.write(f"| `{name[:90]}` | {grid} | " + " | ".join(cells) + " |\n")
    print(open(dst).read()[:5000])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "ncu rows", sys.argv[4] if len(sys.argv) > 4 else None)
