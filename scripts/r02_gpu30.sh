#!/bin/bash
# round 2, 1-GPU job 30: re-create two evidence files lost with an earlier container: L2 fetch-granularity probe of
# the random gather, pow kernel variants (register caps / unrolls)
mkdir -p gpurun_out /tmp/pv
NV="nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I vulkpy_b200/csrc -diag-suppress 177"
$NV -o /tmp/pv/gather_gran scripts/micro/gather_gran.cu && /tmp/pv/gather_gran > gpurun_out/r02_gather_gran.txt 2>&1; cat gpurun_out/r02_gather_gran.txt
{
for v in "v0:" "i2f:-DVKPM_LOG_I2F" "bias:-DVKPM_LOG_BIAS" "mad:-DVKPM_EXP_MAD" "i2f_mad:-DVKPM_LOG_I2F -DVKPM_EXP_MAD" \
         "i2f_mad_b5:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_MINB=5" "i2f_mad_b6:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_MINB=6" \
         "i2f_mad_u2:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_UNROLL=2" "i2f_mad_t128_b12:-DVKPM_LOG_I2F -DVKPM_EXP_MAD -DPV_BLOCK=128 -DPV_MINB=12"; do
  name=${v%%:*}; flags=${v#*:}
  echo "== $name ($flags)"
  $NV $flags -o /tmp/pv/pow_$name scripts/micro/pow_variants.cu && /tmp/pv/pow_$name
done
} > gpurun_out/r02_pow_variants.txt 2>&1
cat gpurun_out/r02_pow_variants.txt | tail -40
