#!/bin/bash
# round 2, 1-GPU job 21: nibble-sliced jump tables -- bit-exact tests, size sweep with / without the start-state pre-pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_random.py tests/test_gpu_large.py tests/test_gpu_reference_trace.py -m gpu -q --timeout 600 > gpurun_out/r02_pytest_prng4.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r02_pytest_prng4.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_prng4.log | head
{ echo "== default"; python scripts/prng_size_sweep.py; echo "== VKP_PRNG_STARTS=0"; VKP_PRNG_STARTS=0 python scripts/prng_size_sweep.py; } > gpurun_out/r02_prng_size_sweep_v2.txt 2>&1
cat gpurun_out/r02_prng_size_sweep_v2.txt
{ for v in "VKP_PRNG_STARTS=1" "VKP_PRNG_STARTS=0"  "VKP_PRNG_STARTS=0 VKP_PRNG_THREADS_PER_SM=2048" "VKP_PRNG_STARTS=1 VKP_PRNG_THREADS_PER_SM=2048"; do
  echo "== $v"
  env $v python scripts/bench_all.py --only "(size=64)" 2>&1 | grep -E "GB/s"
done; } > gpurun_out/r02_prng_variants_v5.txt 2>&1
cat gpurun_out/r02_prng_variants_v5.txt
