#!/bin/bash
# round 2, 1-GPU job 31: new PRNG pre-pass tests, N=1 bench line with the host copy bound beside e2e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_random.py -m gpu -q --timeout 600 > gpurun_out/r02_pytest_prng5.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r02_pytest_prng5.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_prng5.log | head
timeout 900 python bench.py > gpurun_out/r02_bench_n1_v5.json 2> gpurun_out/r02_bench_n1_v5.err
echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_v5.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d.get('frac_of_nominal_8TBs'), d['clocks'])
print(d['roofline']['per_op_frac'])
print({k:(v.get('ms'),v.get('frac')) for k,v in d['configs']['C5']['rows'].items()})
print(d['configs']['C5'].get('mlp_step')['ms'])
print(d['e2e'])
P
