#!/bin/bash
# round 2, 1-GPU job 9: GPU tests (reference-suite trace replay, fast/precise normal), MUFU error sweep,
# L2 fetch-granularity gather probe, pow kernel variants, PRNG timings in both normal modes.  nvcc is on the box.
mkdir -p gpurun_out /tmp/pv
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 300 -s -k "not large" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; grep -E "passed|failed|normal: fast" gpurun_out/pytest_gpu.log | tail -8; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
NV="nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I vulkpy_b200/csrc -diag-suppress 177"
$NV -o /tmp/pv/mufu_error scripts/micro/mufu_error.cu && /tmp/pv/mufu_error > gpurun_out/r02_mufu_error.txt 2>&1; cat gpurun_out/r02_mufu_error.txt
$NV -o /tmp/pv/gather_gran scripts/micro/gather_gran.cu && /tmp/pv/gather_gran > gpurun_out/r02_gather_gran.txt 2>&1; cat gpurun_out/r02_gather_gran.txt
{
for v in "v0:" "i2f:-DVKPM_LOG_I2F" "bias:-DVKPM_LOG_BIAS" "mad:-DVKPM_EXP_MAD" "i2f_mad:-DVKPM_LOG_I2F -DVKPM_EXP_MAD" "bias_mad:-DVKPM_LOG_BIAS -DVKPM_EXP_MAD" \
         "i2f_mad_b5:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_MINB=5" "i2f_mad_b6:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_MINB=6" \
         "i2f_mad_u2:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_UNROLL=2" "i2f_mad_u2_b8:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_UNROLL=2 -DPV_MINB=8" \
         "i2f_mad_t128_b12:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_BLOCK=128 -DPV_MINB=12" "i2f_mad_u1_b8:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_UNROLL=1 -DPV_MINB=8"; do
  name=${v%%:*}; flags=${v#*:}
  echo "== $name ($flags)"
  $NV $flags -o /tmp/pv/pow_$name scripts/micro/pow_variants.cu && /tmp/pv/pow_$name
done
} > gpurun_out/r02_pow_variants.txt 2>&1
cat gpurun_out/r02_pow_variants.txt
ONLY="random 2^30 (size=64),randint 2^30 (size=64),normal 2^30 (size=64),normal 2^30 (size=1048576)"
{ echo "== default (fast normal)"; python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "2\^30"
  echo "== VKP_NORMAL_PRECISE=1"; VKP_NORMAL_PRECISE=1 python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "2\^30"; } > gpurun_out/r02_normal_modes.txt 2>&1
cat gpurun_out/r02_normal_modes.txt
