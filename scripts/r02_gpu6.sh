#!/bin/bash
# round 2, 1-GPU job 6: tests after the fused ReLU epilogues / narrow-N tcgen05 tile / skinny_m rewrite; MLP step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=8 --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
python scripts/mlp_profile.py | head -3
VKP_TC_NO_NARROW=1 python scripts/mlp_profile.py | head -1
ONE_STEP=1 timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r02_mlp_launches.csv python scripts/mlp_profile.py > /dev/null 2>&1
echo "ncu mlp exit $?"
