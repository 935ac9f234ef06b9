"""Roofline sweep over every SURVEY section-8 row on one B200 (device-resident inputs, CUDA events
on the launch stream around INNER back-to-back calls, median of REPS after warm-up).  Writes JSON to stdout / --out.

  C2 elementwise family (incl. in-place, scalar, unary, clamp, broadcast)   bytes per BASELINE.md 4
  C3 reductions: sum/maximum/mean over axis 0 / 1 / None and rebroadcast on 16384^2
  C4 gather of 2^26 random / sorted indices from an 8192^2 table; 8192^3 matmul
  C5 Xoshiro128pp random / randint / normal, 2^30 samples; MLP train step (Dense 1024-1024-16, batch 8192)
"""
import argparse, json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200 import nn
from vulkpy_b200._backend import Timer

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--inner", type=int, default=4, help="calls enqueued back to back between one event pair (hides the host enqueue latency of the first)")
ap.add_argument("--small", action="store_true", help="quarter-size arrays (debug)")
ap.add_argument("--only", default=None, help="comma-separated substrings; time only the rows whose name contains one")
args = ap.parse_args()

PEAK = 6555.2
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

gpu = vk.GPU(0)
dev = gpu.gpu
R = 8192 if args.small else 16384
N = R * R
res = {"peak_hbm_gbs": PEAK, "n_elements": N, "rows": {}}


def timed(fn, reps=args.reps, warm=2):
    for _ in range(warm):
        out = fn()
        del out
    gpu.wait()
    ts = []
    for _ in range(reps):
        t0, t1 = Timer(dev), Timer(dev)
        l0 = dev.launch_count()
        t0.record()
        for _ in range(args.inner):
            out = fn()
        t1.record()
        ts.append(t0.elapsed_ms(t1) / args.inner)
        launches = (dev.launch_count() - l0) // args.inner
        del out
    return statistics.median(ts), min(ts), launches


def row(name, fn, nbytes=None, flops=None, note=None):
    if args.only and not any(k in name for k in args.only.split(",")):
        res["rows"][name] = {"skipped": True, "tflops": 0.0}
        return
    ms, best, launches = timed(fn)
    r = {"ms": round(ms, 4), "best_ms": round(best, 4), "launches": launches}
    if nbytes is not None:
        r["alg_bytes"] = int(nbytes)
        r["gbs"] = round(nbytes / ms / 1e6, 1)
        r["frac_of_measured_peak"] = round(nbytes / ms / 1e6 / PEAK, 4)
        r["frac_of_8TBs"] = round(nbytes / ms / 1e6 / 8000, 4)
    if flops is not None:
        r["tflops"] = round(flops / ms / 1e9, 2)
    if note:
        r["note"] = note
    res["rows"][name] = r
    print(f"{name:34s} {ms:9.4f} ms  " + (f"{r.get('gbs', 0):8.1f} GB/s {100 * r.get('frac_of_measured_peak', 0):6.1f}%" if nbytes else f"{r.get('tflops', 0):8.2f} TFLOP/s"), flush=True)


rng = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=1234)
a = rng.random(shape=(R, R)); a *= 1.5; a += 0.5
b = rng.random(shape=(R, R)); b *= 4.0; b -= 2.0
rowv = rng.random(shape=(R,))
colv = rng.random(shape=(R, 1))
gpu.wait()
B4 = 4 * N

# ---- C2 -----------------------------------------------------------------------------------------
row("a+b", lambda: a + b, 3 * B4)
row("a-b", lambda: a - b, 3 * B4)
row("a*b", lambda: a * b, 3 * B4)
row("a/b", lambda: a / b, 3 * B4)
row("a.max(b)", lambda: a.max(b), 3 * B4)
c = a + b
row("c+=b (in place)", lambda: c.__iadd__(b), 3 * B4)
row("a*2.5", lambda: a * 2.5, 2 * B4)
row("2.5-a", lambda: 2.5 - a, 2 * B4)
row("c*=1.0 (in place scalar)", lambda: c.__imul__(1.0), 2 * B4)
row("a+row (16384,)", lambda: a + rowv, 2 * B4 + 4 * R)
row("a*col (16384,1)", lambda: a * colv, 2 * B4 + 4 * R)
row("c+=row (in place broadcast)", lambda: c.__iadd__(rowv), 2 * B4 + 4 * R)
row("row.broadcast_to", lambda: rowv.broadcast_to((R, R)), B4 + 4 * R)
uni = rng.random(shape=(R, R)); uni *= 1.8; uni -= 0.9       # (-0.9, 0.9) for asin/acos/atanh
for fn in ("abs", "sign", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "log", "exp2", "log2", "sqrt", "invsqrt"):
    src = a if fn in ("sqrt", "invsqrt", "log", "log2") else (uni if fn in ("asin", "acos", "atanh") else b)
    if fn == "acosh":
        src = a + 0.6
    row(f"{fn}(x)", (lambda f, s: (lambda: getattr(s, f)()))(fn, src), 2 * B4)
del uni
row("a**b", lambda: a ** b, 3 * B4)
row("a**2.7", lambda: a ** 2.7, 2 * B4)
row("1.3**b", lambda: 1.3 ** b, 2 * B4)
row("clamp(s,s)", lambda: a.clamp(0.75, 1.5), 2 * B4)
row("clamp(v,s)", lambda: a.clamp(b, 1.5), 3 * B4)
row("clamp(v,v)", lambda: c.clamp(b, a), 4 * B4)
del c

# ---- C3 -----------------------------------------------------------------------------------------
for op in ("sum", "maximum", "mean"):
    row(f"{op}(axis=0)", (lambda o: (lambda: getattr(a, o)(axis=0)))(op), B4 + 4 * R)
    row(f"{op}(axis=1)", (lambda o: (lambda: getattr(a, o)(axis=1)))(op), B4 + 4 * R)
    row(f"{op}(axis=None)", (lambda o: (lambda: getattr(a, o)()))(op), B4 + 4)
row("sum(axis=1,rebroadcast)", lambda: a.sum(axis=1, rebroadcast=True), 2 * B4)
row("maximum(axis=0,rebroadcast)", lambda: a.maximum(axis=0, rebroadcast=True), 2 * B4)
row("argmax(axis=1)", lambda: a.argmax(axis=1), B4 + 4 * R)
row("argmax(axis=0)", lambda: a.argmax(axis=0), B4 + 4 * R)
row("argmin(axis=None)", lambda: a.argmin(), B4 + 4)
a3 = a
a3.reshape((R // 64, 64, R))
row("sum(axis=1) of (R/64,64,R)", lambda: a3.sum(axis=1), B4 + 4 * N // 64)
a3.reshape((R, R // 8, 8))
row("sum(axis=1) of (R,R/8,8)", lambda: a3.sum(axis=1), B4 + 4 * R * 8)
a.reshape((R, R))

# ---- C4 gather ------------------------------------------------------------------------------------
G = 4096 if args.small else 8192
NI = 1 << (24 if args.small else 26)
table = rng.random(shape=(G, G))
idx_h = np.random.default_rng(99).integers(0, G * G, NI, dtype=np.uint32)
idx = vk.U32Array(gpu, data=idx_h)
idx_sorted = vk.U32Array(gpu, data=np.sort(idx_h))
row("gather 2^26 random idx", lambda: table.gather(idx), 12 * NI, note="table 8192^2 (256 MiB)")
row("gather 2^26 sorted idx", lambda: table.gather(idx_sorted), 12 * NI)
lab = vk.U32Array(gpu, data=np.random.default_rng(1).integers(0, 16, 1 << 20, dtype=np.uint32))
row("to_onehot(16) of 2^20 labels", lambda: lab.to_onehot(16), 4 * (1 << 20) * 17)
del idx, idx_sorted, lab

# ---- C4 matmul --------------------------------------------------------------------------------------
M = 4096 if args.small else 8192
ma = rng.random(shape=(M, M)); ma -= 0.5
mb = rng.random(shape=(M, M)); mb -= 0.5
row(f"matmul {M}^3 (3xTF32 tcgen05)", lambda: ma @ mb, flops=2 * M ** 3)
res["rows"][f"matmul {M}^3 (3xTF32 tcgen05)"]["tf32_pipe_tflops_issued"] = round(3 * res["rows"][f"matmul {M}^3 (3xTF32 tcgen05)"]["tflops"], 1)
del ma, mb, table

# ---- C5 PRNG ----------------------------------------------------------------------------------------
P = 1 << (28 if args.small else 30)
buf = vk.Array(gpu, shape=(P,))
ubuf = vk.U32Array(gpu, shape=(P,))
for size in (64, 1 << 20):
    g = vk.random.Xoshiro128pp(gpu, size=size, seed=7)
    row(f"random 2^30 (size={size})", (lambda gg: (lambda: gg.random(buffer=buf)))(g), 4 * P)
    row(f"randint 2^30 (size={size})", (lambda gg: (lambda: gg.randint(buffer=ubuf)))(g), 4 * P)
    row(f"normal 2^30 (size={size})", (lambda gg: (lambda: gg.normal(buffer=buf)))(g), 4 * P)
del buf, ubuf
g = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=7)
row("permutation(2^24)", lambda: g.permutation(1 << 24), 4 * (1 << 24) * 2, note="bytes = keys written + indices written; the radix sort moves more")

# ---- C5 MLP step ---------------------------------------------------------------------------------------
Bsz, D, H, C = (2048 if args.small else 8192), 1024, 1024, 16
opt = nn.Adam(gpu, lr=1e-3)
net = nn.Sequence([nn.Dense(gpu, D, H, w_opt=opt, b_opt=opt, w_init=nn.HeNormal(gpu, D, seed=1)), nn.ReLU(),
                   nn.Dense(gpu, H, C, w_opt=opt, b_opt=opt, w_init=nn.HeNormal(gpu, H, seed=2)), nn.Softmax()],
                  nn.CrossEntropyLoss())
x = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=3).normal(shape=(Bsz, D))
y = vk.random.Xoshiro128pp(gpu, seed=4).randrange(shape=(Bsz,), low=0, high=C).to_onehot(C)
gpu.wait()
row("MLP train step (B=8192)", lambda: net.train(x, y), flops=6 * Bsz * (D * H + H * C),
    note="Dense(1024,1024)-ReLU-Dense(1024,16)-Softmax, CrossEntropy, Adam; host-enqueued op by op")
t0 = time.perf_counter()
for _ in range(5):
    net.train(x, y)
gpu.wait()
res["rows"]["MLP train step (B=8192)"]["wall_ms_per_step"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)

txt = json.dumps(res, indent=1)
if args.out:
    open(args.out, "w").write(txt)
print("done")
