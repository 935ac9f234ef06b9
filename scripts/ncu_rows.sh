#!/bin/bash
# Per-kernel ncu metrics for the SURVEY section-8 rows that bench.py's own capture does not cover
# (reductions, broadcast, gather, PRNG, GEMM, nn).  One GPU.  Output: gpurun_out/ncu_rows.csv
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,lts__t_sector_hit_rate.pct
timeout 900 ncu --metrics $M --clock-control none -k regex:'reduce_|gather|xoshiro|bcast|gemm_tc|transpose|split_lo|arg|softmax|adam|relu' \
  --csv --log-file gpurun_out/ncu_rows.csv \
  python scripts/bench_all.py --reps 1 --inner 1 --only "sum(,maximum(axis=0),mean(axis=1),argmax,gather,a+row,a*col,broadcast_to,matmul,random 2^30 (size=1048576),normal 2^30 (size=1048576),randint 2^30 (size=64),MLP" \
  > gpurun_out/ncu_rows.log 2>&1
echo "ncu rows exit $?"; wc -l gpurun_out/ncu_rows.csv
