#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,sm__cycles_elapsed.max
for st in ${STS:-1 0}; do
VKP_PRNG_STARTS=$st timeout 600 ncu --metrics $M --clock-control none -k regex:'xoshiro' -s 0 -c 40 --csv \
  --log-file gpurun_out/r02_ncu_rows_v6_$st.csv python scripts/r02_probe.py prng > gpurun_out/r02_ncu_rows_v6.log 2>&1
echo "ncu rows exit $?"
done
