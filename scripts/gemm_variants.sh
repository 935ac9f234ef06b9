#!/bin/bash
# GEMM variants side by side (one process each: the switches are read once per process).
mkdir -p gpurun_out
export GEMM_BENCH_NO_SIMT=1
for v in "VKP_TC_PRESPLIT=0" "VKP_TC_PRESPLIT=1" "VKP_TC_PRESPLIT=1 VKP_TC_BK16=0" "VKP_TC_PRESPLIT=0 VKP_TC_BK16=1"; do
  echo "== $v"
  env $v timeout 120 python scripts/gemm_bench.py 2>&1 | python -c "
import sys, json
try:
    d = json.load(sys.stdin)
    for k, r in d.items(): print(k, r)
except Exception as e:
    print('failed', e)
"
done
echo "== tests with presplit forced"
VKP_TC_PRESPLIT=1 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -5
echo "== tests with presplit + BK32 forced"
VKP_TC_PRESPLIT=1 VKP_TC_BK16=0 timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x 2>&1 | tail -5
echo "== tests default"
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_nn.py -m gpu -q -x 2>&1 | tail -5
