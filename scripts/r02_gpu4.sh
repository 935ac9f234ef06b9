#!/bin/bash
# round 2, 1-GPU job 4: all GPU tests (fusion, config 5, shim), PRNG store / segment variants at the default 64 lanes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
ONLY="random 2^30 (size=64),randint 2^30 (size=64),normal 2^30 (size=64),random 2^30 (size=1048576)"
for v in "base" "VKP_PRNG_STCS=1" "VKP_PRNG_THREADS_PER_SM=2048" "VKP_PRNG_THREADS_PER_SM=4096" "VKP_PRNG_STCS=1 VKP_PRNG_THREADS_PER_SM=4096" "VKP_PRNG_THREADS_PER_SM=512"; do
  echo "== $v"
  if [ "$v" = "base" ]; then env python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "2\^30"; else env $v python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "2\^30"; fi
done > gpurun_out/r02_prng_variants.txt 2>&1
cat gpurun_out/r02_prng_variants.txt
