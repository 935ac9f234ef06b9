#!/bin/bash
# round 2, 1-GPU job: GPU tests (incl. config-5 step, compat shim, tall matvec), smoke, the full bench line
# (C2 + roofline_matmul + C3/C4/C5 + cpu_baseline at 2^28) and ncu rows of the reworked pow / log / PRNG kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench exit $?"; tail -c 1500 gpurun_out/r02_bench_n1.json; tail -5 gpurun_out/r02_bench_n1.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'xoshiro|ew_tab|UAsinh' -s 0 -c 40 --csv \
  --log-file gpurun_out/r02_ncu_rows_v2.csv python scripts/r02_probe.py pow log prng > gpurun_out/r02_ncu_rows_v2.log 2>&1
echo "ncu rows exit $?"
