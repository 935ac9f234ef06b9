#!/bin/bash
# round 2, 1-GPU job 14: binomial a**s kernel, second version (32-entry lane tables) -- parity, timing, ncu; PRNG rows
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_array.py -m gpu -q --timeout 300 -k "pow" > gpurun_out/r02_pytest_pow.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/r02_pytest_pow.log
ONLY="a**2.7,a**b,1.3**b"
{ echo "== binomial kernel v2 (default)"; python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "GB/s"
  echo "== VKP_POWS_BINOMIAL=0 (general table kernel)"; VKP_POWS_BINOMIAL=0 python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "GB/s"; } > gpurun_out/r02_pows_timing_v2.txt 2>&1
cat gpurun_out/r02_pows_timing_v2.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none -k regex:'ew_tab|ew_pows|xoshiro' -s 0 -c 40 --csv \
  --log-file gpurun_out/r02_ncu_rows_v4.csv python scripts/r02_probe.py pow prng > gpurun_out/r02_ncu_rows_v4.log 2>&1
echo "ncu rows exit $?"
