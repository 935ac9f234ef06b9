#!/bin/bash
# round 2, 1-GPU job 13: binomial-series a**s kernel -- parity tests, timing against the general kernel,
# ncu pipe metrics; compute-sanitizer memcheck / racecheck over the GEMM modes, the TMA reduction, element-wise.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_array.py -m gpu -q --timeout 300 -k "pow" > gpurun_out/r02_pytest_pow.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r02_pytest_pow.log
ONLY="a**2.7,a**b,1.3**b"
{ echo "== binomial kernel (default)"; python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "GB/s"
  echo "== VKP_POWS_BINOMIAL=0 (general table kernel)"; VKP_POWS_BINOMIAL=0 python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "GB/s"; } > gpurun_out/r02_pows_timing.txt 2>&1
cat gpurun_out/r02_pows_timing.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'ew_tab|ew_pows|pows_build' -s 0 -c 20 --csv \
  --log-file gpurun_out/r02_ncu_rows_v3.csv python scripts/r02_probe.py pow > gpurun_out/r02_ncu_rows_v3.log 2>&1
echo "ncu rows exit $?"
for t in gemm reduce ew; do
  for tool in memcheck racecheck; do
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_targets.py $t > gpurun_out/r02_sanitize_${tool}_${t}.log 2>&1
    echo "sanitize $tool $t exit $?"; tail -2 gpurun_out/r02_sanitize_${tool}_${t}.log
  done
done
VKP_TC_PRESPLIT=1 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_targets.py gemm > gpurun_out/r02_sanitize_racecheck_gemm_presplit.log 2>&1
echo "sanitize racecheck gemm presplit exit $?"; tail -2 gpurun_out/r02_sanitize_racecheck_gemm_presplit.log
VKP_TC_REWRITE_HI=1 timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_targets.py gemm > gpurun_out/r02_sanitize_racecheck_gemm_rewrite.log 2>&1
echo "sanitize racecheck gemm rewrite_hi exit $?"; tail -2 gpurun_out/r02_sanitize_racecheck_gemm_rewrite.log
