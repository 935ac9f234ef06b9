#!/bin/bash
# round 2, 1-GPU job 5: GPU tests with the MN-major GEMM operands / relaxed pre-split rule, GEMM variants, MLP step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
for v in "base" "VKP_TC_MN=0"; do
  echo "== $v"
  if [ "$v" = "base" ]; then python scripts/gemm_bench.py; python scripts/mlp_profile.py | head -3; else env $v python scripts/gemm_bench.py; env $v python scripts/mlp_profile.py | head -3; fi
done > gpurun_out/r02_gemm_mn.txt 2>&1
cat gpurun_out/r02_gemm_mn.txt | cut -c1-1500
ONE_STEP=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_mlp_launches.csv python scripts/mlp_profile.py > /dev/null 2>&1
echo "ncu mlp exit $?"
