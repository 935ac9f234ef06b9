#!/bin/bash
# round 2, 1-GPU job 27 (final N=1 record): full GPU suite, smoke, the N=1 bench line, the reference arm, the ncu
# launch list of the same bench command, one `ncu --set full` capture of the step's kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_box.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 > gpurun_out/r02_pytest_gpu_v5.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r02_pytest_gpu_v5.log; grep -E "passed|failed" gpurun_out/r02_pytest_gpu_v5.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_gpu_v5.log | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_v5.json 2> gpurun_out/r02_bench_n1_v5.err
echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_v5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('frac_of_nominal_8TBs'), d['clocks'])
print(d['roofline']['per_op_frac'], d['roofline']['worst'])
for c in ('C3','C4','C5'):
    print(c, {k:(v.get('ms'),v.get('frac')) for k,v in d['configs'][c]['rows'].items()})
print(d['configs']['C5'].get('mlp_step'))
print(d['e2e'], d.get('gpu_launches'))
P
tail -3 gpurun_out/r02_bench_n1_v5.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1_v5.json 2>&1; tail -1 gpurun_out/r02_bench_ref_n1_v5.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench_v5.csv python bench.py --steps 2 --warmup 1 --e2e-steps 2 --matmul-seconds 0.2 > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r02_launches_bench_v5.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ew_kernel|ew_tab_kernel|ew_pows_kernel|bcast_vec' -s 40 -c 16 -o gpurun_out/r02_prof_full -f \
    python bench.py --steps 1 --warmup 3 --no-matmul --no-configs --e2e-steps 1 > gpurun_out/r02_bench_under_ncu_full.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/*.ncu-rep
