"""Executed-instruction histogram of one kernel from an ncu report's source page.

  python scripts/ncu_source.py gpurun_out/prof_full.ncu-rep TPowScalar [--top 25] [--lines]
Prints warp-instructions executed per opcode (and, with --lines, the hottest SASS lines)."""
import collections, csv, io, re, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in txt.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(line)
for b in blocks:
    if not re.search(pat, b["name"]):
        continue
    rd = list(csv.reader(io.StringIO("\n".join(b["rows"]))))
    hdr = rd[0]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    hist, tot, lines = collections.Counter(), 0, []
    for r in rd[1:]:
        if len(r) <= iex:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        n = int(r[iex] or 0)
        hist[m.group(2) if m else "?"] += n
        tot += n
        lines.append((n, int(r[ismp] or 0), r[isrc].strip()))
    print(b["name"][:150])
    print("total warp-instructions executed:", tot)
    for k, v in hist.most_common(top):
        print(f"  {k:24s} {v:12d} {100.0 * v / tot:6.2f}%")
    if "--lines" in sys.argv:
        for n, smp, s in lines:
            print(f"{n:10d} {smp:6d}  {s}")
    break
