#!/bin/bash
# round 2, 1-GPU job 33: 128x224 tcgen05 tile -- GEMM tests, timing of a rank's share of the fixed-size 8192^3 product
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_nn.py tests/test_gpu_config5.py -m gpu -q --timeout 600 2>&1 | tail -4
cat > /tmp/bn224.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, vulkpy_b200 as vk
from vulkpy_b200._backend import Timer
gpu = vk.GPU(0); dev = gpu.gpu
rng = vk.random.Xoshiro128pp(gpu, size=1 << 16, seed=1)
for (M, N, K) in [(1024, 8192, 8192), (2048, 8192, 8192), (8192, 8192, 8192)]:
    a = rng.random(shape=(M, K)); bt = rng.random(shape=(N, K)); c = vk.Array(gpu, shape=(M, N))
    for _ in range(3):
        c.job = dev.gemm(False, True, M, N, K, a.buffer, bt.buffer, c.buffer, None, 2)
    gpu.wait()
    t0, t1 = Timer(dev), Timer(dev); t0.record()
    for _ in range(10):
        c.job = dev.gemm(False, True, M, N, K, a.buffer, bt.buffer, c.buffer, None, 2)
    t1.record(); ms = t0.elapsed_ms(t1) / 10
    print(f"{M}x{N}x{K} NT: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
P
{ echo "== default (BN=224 where it saves waves)"; python /tmp/bn224.py; echo "== VKP_TC_BN224=0"; VKP_TC_BN224=0 python /tmp/bn224.py; } > gpurun_out/r02_bn224.txt 2>&1
cat gpurun_out/r02_bn224.txt
