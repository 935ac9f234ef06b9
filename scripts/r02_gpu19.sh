#!/bin/bash
# round 2, 1-GPU job 19: segment start states by doubling (xoshiro_starts_kernel) -- bit-exact tests, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_random.py tests/test_gpu_large.py tests/test_gpu_reference_trace.py -m gpu -q --timeout 600 > gpurun_out/r02_pytest_prng3.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r02_pytest_prng3.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_prng3.log | head
{ for v in "VKP_PRNG_STARTS=1" "VKP_PRNG_STARTS=0" "VKP_PRNG_STARTS=1 VKP_PRNG_THREADS_PER_SM=2048" "VKP_PRNG_STARTS=1 VKP_PRNG_THREADS_PER_SM=1536" "VKP_PRNG_STARTS=1 VKP_PRNG_THREADS_PER_SM=512"; do
  echo "== $v"
  env $v python scripts/bench_all.py --only "(size=64)" 2>&1 | grep -E "GB/s"
done; } > gpurun_out/r02_prng_variants_v4.txt 2>&1
cat gpurun_out/r02_prng_variants_v4.txt
M=gpu__time_duration.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'xoshiro' -s 0 -c 40 --csv \
  --log-file gpurun_out/r02_ncu_rows_v5.csv python scripts/r02_probe.py prng > gpurun_out/r02_ncu_rows_v5.log 2>&1
echo "ncu rows exit $?"
