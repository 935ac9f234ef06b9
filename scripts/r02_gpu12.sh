#!/bin/bash
# round 2, 1-GPU job 12 (after the container was re-created): full GPU suite, smoke, the N=1 bench line,
# the reference arm, and the ncu launch list of the same bench command.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_box.txt
timeout 1200 python -m pytest tests -m gpu -q --maxfail=10 --timeout 600 > gpurun_out/r02_pytest_gpu_v3.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/r02_pytest_gpu_v3.log; grep -E "passed|failed" gpurun_out/r02_pytest_gpu_v3.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_gpu_v3.log | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_v3.json 2> gpurun_out/r02_bench_n1_v3.err
echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_v3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('frac_of_nominal_8TBs'), d['clocks'])
print(d['roofline'])
for c in ('C3','C4','C5'):
    print(c, {k:(v.get('ms'),v.get('frac')) for k,v in d['configs'][c]['rows'].items()})
print(d['configs']['C5'].get('mlp_step'))
print({k:v for k,v in d.items() if 'matmul' in k})
print(d['e2e'], d.get('gpu_launches'))
P
tail -3 gpurun_out/r02_bench_n1_v3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_n1_v3.json 2>&1; tail -1 gpurun_out/r02_bench_ref_n1_v3.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_v3.csv python bench.py --steps 2 --warmup 1 --e2e-steps 2 --matmul-seconds 0.2 > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/r02_launches_bench_v3.csv
