"""CPU baseline for BASELINE.json configs 1-5: the C restatement of the
reference's shaders (oracle/cpu_ref.c: one element per "invocation", serial per-output axis loops,
naive matmul loop, float32, OpenMP over all host cores) timed at config C1 exactly and at reduced
sizes of C2-C5, each also scaled linearly to the full size.  The reference's own SPIR-V on Mesa
lavapipe cannot run in this image (no Vulkan loader / ICD / glslc), so every row is labelled
"restatement".  A reported baseline, not an optimisation target.

    python scripts/cpu_baseline_all.py [--out file.json]
"""
import argparse, ctypes as C, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import cpu_ref

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

L = cpu_ref.load()
cores = len(os.sched_getaffinity(0))
L.ref_set_num_threads(cores)
ptr = cpu_ref.ptr
F = np.float32
rs = np.random.default_rng(0)
res = {"kind": "restatement (lavapipe unavailable)", "cores": L.ref_num_threads(), "host": os.uname().nodename,
       "cpu": next((l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?"), "rows": {}}


def timed(fn):
    fn()
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3


def row(name, fn, *, nbytes=None, flops=None, samples=None, scale=1.0, full=None):
    ms = timed(fn)
    r = {"ms": round(ms, 3)}
    if nbytes is not None:
        r["gbs"] = round(nbytes / ms / 1e6, 2)
    if flops is not None:
        r["gflops"] = round(flops / ms / 1e6, 2)
    if samples is not None:
        r["msamples_per_s"] = round(samples / ms / 1e3, 1)
    if full:
        r["full_size"] = full
        r["ms_scaled_to_full_size"] = round(ms * scale, 1)
    res["rows"][name] = r
    print(f"{name:52s} {ms:10.3f} ms  " + json.dumps({k: v for k, v in r.items() if k != 'ms'}), flush=True)


# ---- C1: example/00-arithmetic.py style on 4096 x 4096 (exact size) ------------------------------------
n1 = 4096
a = rs.uniform(0, 1, (n1, n1)).astype(F); b = rs.uniform(0, 1, (n1, n1)).astype(F); c = np.empty_like(a)
col = np.empty(n1, F)
row("C1 a+b 4096^2", lambda: L.ref_binary(0, ptr(a), ptr(b), ptr(c), a.size), nbytes=12 * a.size)
row("C1 a*b 4096^2", lambda: L.ref_binary(2, ptr(a), ptr(b), ptr(c), a.size), nbytes=12 * a.size)
row("C1 a.sum(axis=0) 4096^2", lambda: L.ref_reduce_axis(0, ptr(a), ptr(col), 1, n1, n1, 0), nbytes=4 * a.size + 4 * n1)

# ---- C2 (reduced: 2^24 elements; full 2^28) -> bench.py carries the 12-op list; two more families here ----
row("C2 sin(x) 2^24", lambda: L.ref_unary(2, ptr(a), ptr(c), a.size), nbytes=8 * a.size, scale=16, full="2^28")
row("C2 a**b 2^24", lambda: L.ref_binary(6, ptr(a), ptr(b), ptr(c), a.size), nbytes=12 * a.size, scale=16, full="2^28")

# ---- C3 (reduced: 4096^2; full 16384^2) --------------------------------------------------------------------
rowv = np.empty(n1, F); tmp = np.empty(2 * ((a.size + 63) // 64) + 64, F)
row("C3 sum(axis=1) 4096^2", lambda: L.ref_reduce_axis(0, ptr(a), ptr(rowv), n1, n1, 1, 0), nbytes=4 * a.size, scale=16, full="16384^2")
row("C3 maximum(axis=0) 4096^2", lambda: L.ref_reduce_axis(2, ptr(a), ptr(col), 1, n1, n1, 0), nbytes=4 * a.size, scale=16, full="16384^2")
row("C3 sum(axis=None) 4096^2", lambda: L.ref_reduce_full(0, ptr(a), a.size, ptr(tmp)), nbytes=4 * a.size, scale=16, full="16384^2")
row("C3 sum(axis=1, rebroadcast) 4096^2", lambda: L.ref_reduce_axis(0, ptr(a), ptr(c), n1, n1, 1, 1), nbytes=8 * a.size, scale=16, full="16384^2")

# ---- C4 (reduced: 1024^3 matmul, 2^22 gather indices from a 2048^2 table; full 8192^3 / 2^26 from 8192^2) ----
m = 1024
ma = rs.uniform(-1, 1, (m, m)).astype(F); mb = rs.uniform(-1, 1, (m, m)).astype(F); mc = np.empty_like(ma)
row("C4 matmul 1024^3", lambda: L.ref_matmul(ptr(ma), ptr(mb), ptr(mc), m, m, m), flops=2 * m ** 3, scale=512, full="8192^3")
tab = rs.uniform(0, 1, 2048 * 2048).astype(F)
idx = rs.integers(0, tab.size, 1 << 22, dtype=np.uint32); gout = np.empty(idx.size, F)
row("C4 gather 2^22 random idx (table 2048^2)", lambda: L.ref_gather(ptr(tab), ptr(idx), ptr(gout), idx.size), nbytes=12 * idx.size, scale=16, full="2^26 from 8192^2")

# ---- C5 (reduced: 2^24 samples; MLP batch 1024; full 2^30 / batch 8192 per GPU) --------------------------------
P = 1 << 24
for size in (64, 1 << 20):
    state = np.empty((size, 4), np.uint32)
    L.ref_xoshiro_seed(ptr(state), size, C.c_uint64(7)) if size == 64 else state.__setitem__(slice(None), rs.integers(1, 2 ** 32, (size, 4), dtype=np.uint32))
    out = np.empty(P, np.uint32)
    row(f"C5 random 2^24 (size={size})", (lambda st, o: (lambda: L.ref_xoshiro_fill(ptr(st), st.shape[0], ptr(o), C.c_uint64(P), 1)))(state, out),
        nbytes=4 * P, samples=P, scale=64, full="2^30")
    uf = out.view(F)
    row(f"C5 normal 2^24 (size={size}; uniform + Box-Muller)",
        (lambda st, o, u: (lambda: (L.ref_xoshiro_fill(ptr(st), st.shape[0], ptr(o), C.c_uint64(P), 1),
                                    L.ref_box_muller(ptr(u), ptr(u), P, C.c_float(0), C.c_float(1)))))(state, out, uf),
        nbytes=4 * P, samples=P, scale=64, full="2^30")
B, D, H, Cc = 1024, 1024, 1024, 16
x = rs.normal(size=(B, D)).astype(F); w1 = rs.normal(size=(H, D)).astype(F); b1 = np.zeros(H, F); h = np.empty((B, H), F)
w2 = rs.normal(size=(Cc, H)).astype(F); b2 = np.zeros(Cc, F); y = np.empty((B, Cc), F)
dy = rs.normal(size=(B, H)).astype(F); dw = np.empty((H, D), F); dx = np.empty((B, D), F)
dyT = np.ascontiguousarray(dy.T)


def mlp_gemms():   # the contractions of one Dense(1024,1024)-ReLU-Dense(1024,16) step (forward x2, dW1, dx1)
    L.ref_batch_affine(ptr(w1), ptr(b1), ptr(x), ptr(h), B, D, H)
    L.ref_batch_affine(ptr(w2), ptr(b2), ptr(h), ptr(y), B, H, Cc)
    L.ref_matmul(ptr(dyT), ptr(x), ptr(dw), H, B, D)
    L.ref_matmul(ptr(dy), ptr(w1), ptr(dx), B, H, D)


row("C5 MLP step contractions, batch 1024", mlp_gemms, flops=2 * B * (D * H + H * Cc) + 4 * B * D * H, scale=8, full="batch 8192 per GPU")

txt = json.dumps(res, indent=1)
if args.out:
    open(args.out, "w").write(txt)
print("done")
