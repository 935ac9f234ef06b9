#!/bin/bash
# round 2, GPU job 1: GPU tests after the ADVICE fixes, full ncu captures of the kernels the verdict
# names (pow / log / PRNG / TMA column reduction / random gather), compute-sanitizer over the
# mbarrier / TMA / tcgen05 kernels.  One GPU.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 500 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__sectors_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'reduce_|gather|xoshiro|ew_tab|UAsinh' -s 0 -c 60 --csv \
  --log-file gpurun_out/r02_ncu_rows.csv python scripts/r02_probe.py > gpurun_out/r02_ncu_rows.log 2>&1
echo "ncu rows exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ew_tab_kernel|xoshiro_normal|xoshiro_stream' -c 12 \
  -o gpurun_out/r02_pow_prng -f python scripts/r02_probe.py pow log prng > gpurun_out/r02_pow_prng.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
for t in gemm reduce ew; do
  for tool in memcheck racecheck; do
    timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_targets.py $t > gpurun_out/r02_sanitize_${tool}_${t}.log 2>&1
    echo "sanitize $tool $t exit $?"; tail -3 gpurun_out/r02_sanitize_${tool}_${t}.log
  done
done
VKP_TC_PRESPLIT=1 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_targets.py gemm > gpurun_out/r02_sanitize_racecheck_gemm_presplit.log 2>&1
echo "sanitize racecheck gemm presplit exit $?"; tail -3 gpurun_out/r02_sanitize_racecheck_gemm_presplit.log
VKP_TC_REWRITE_HI=1 timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_targets.py gemm > gpurun_out/r02_sanitize_racecheck_gemm_rewrite.log 2>&1
echo "sanitize racecheck gemm rewrite_hi exit $?"; tail -3 gpurun_out/r02_sanitize_racecheck_gemm_rewrite.log
