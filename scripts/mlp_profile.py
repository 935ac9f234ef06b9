"""Config-5 MLP train step: device time per step, host enqueue time per step, cProfile of the host side."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk
from vulkpy_b200 import nn
from vulkpy_b200._backend import Timer

gpu = vk.GPU(0)
dev = gpu.gpu
opt = lambda: nn.Adam(gpu, lr=1e-3)
net = nn.Sequence([nn.Dense(gpu, 1024, 1024, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 1024, seed=1)), nn.ReLU(),
                   nn.Dense(gpu, 1024, 16, w_opt=opt(), b_opt=opt(), w_init=nn.HeNormal(gpu, 1024, seed=2)), nn.Softmax()],
                  nn.CrossEntropyLoss())
B = 8192
x = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=100).normal(shape=(B, 1024))
y = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=200).randrange(shape=(B,), low=0, high=16).to_onehot(16)
for _ in range(5):
    net.train(x, y)
gpu.wait()
steps = int(os.environ.get("STEPS", "50"))
if os.environ.get("ONE_STEP"):
    net.train(x, y); gpu.wait(); sys.exit(0)
l0 = dev.launch_count()
t0, t1 = Timer(dev), Timer(dev)
h0 = time.perf_counter()
t0.record()
for _ in range(steps):
    net.train(x, y)
t1.record()
h1 = time.perf_counter()
ms = t0.elapsed_ms(t1) / steps
print(f"device ms/step {ms:.4f}  host enqueue ms/step {(h1 - h0) * 1e3 / steps:.4f}  launches/step {(dev.launch_count() - l0) / steps:.1f}")
pr = cProfile.Profile()
pr.enable()
for _ in range(steps):
    net.train(x, y)
pr.disable()
gpu.wait()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:6000])
