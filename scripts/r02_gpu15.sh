#!/bin/bash
# round 2, 1-GPU job 15: fast normal with the series off the common path + unit mean/stddev kernel; MUFU error sweep
# with the new threshold; PRNG rows; threads-per-SM sweep at the default 64 lanes
mkdir -p gpurun_out /tmp/pv
timeout 900 python -m pytest tests/test_gpu_random.py tests/test_gpu_reference_trace.py tests/test_gpu_large.py -m gpu -q --timeout 600 -s > gpurun_out/r02_pytest_prng.log 2>&1
echo "pytest exit $?"; grep -E "passed|failed|normal: fast" gpurun_out/r02_pytest_prng.log | tail -4; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_prng.log | head
NV="nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I vulkpy_b200/csrc -diag-suppress 177"
$NV -o /tmp/pv/mufu_error scripts/micro/mufu_error.cu && /tmp/pv/mufu_error > gpurun_out/r02_mufu_error_v2.txt 2>&1; cat gpurun_out/r02_mufu_error_v2.txt
ONLY="random 2^30,randint 2^30,normal 2^30"
{ for v in "base" "VKP_PRNG_THREADS_PER_SM=1536" "VKP_PRNG_THREADS_PER_SM=2048" "VKP_PRNG_THREADS_PER_SM=4096"; do
  echo "== $v"
  if [ "$v" = "base" ]; then python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "GB/s"; else env $v python scripts/bench_all.py --only "(size=64)" 2>&1 | grep -E "GB/s"; fi
done; } > gpurun_out/r02_prng_variants_v2.txt 2>&1
cat gpurun_out/r02_prng_variants_v2.txt
