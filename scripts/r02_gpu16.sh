#!/bin/bash
# round 2, 1-GPU job 16: 64-lane PRNG, fewer / fatter segment streams
mkdir -p gpurun_out
{ for v in "VKP_PRNG_THREADS_PER_SM=256" "VKP_PRNG_THREADS_PER_SM=512" "VKP_PRNG_THREADS_PER_SM=768" "VKP_PRNG_THREADS_PER_SM=1024" "VKP_PRNG_THREADS_PER_SM=1024 VKP_PRNG_STCS=0" "VKP_PRNG_THREADS_PER_SM=512 VKP_PRNG_STCS=0"; do
  echo "== $v"
  env $v python scripts/bench_all.py --only "(size=64)" 2>&1 | grep -E "GB/s"
done; } > gpurun_out/r02_prng_variants_v3.txt 2>&1
cat gpurun_out/r02_prng_variants_v3.txt
