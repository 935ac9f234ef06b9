#!/bin/bash
# round 2, 1-GPU job 32: uploads take a fresh block instead of waiting for compute -- copy tests, e2e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_copy.py tests/test_gpu_array.py -m gpu -q --timeout 300 2>&1 | tail -2
timeout 900 python bench.py --no-matmul --no-configs > gpurun_out/r02_bench_e2e.json 2> gpurun_out/r02_bench_e2e.err
echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_e2e.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'])
print(d['e2e'])
P
