#!/bin/bash
# round 2, 1-GPU job 10: GPU tests after the tab-kernel split / vector scalar-pow / gather shape / fast normal,
# fast-normal segment variants at the default 64 lanes, the full N=1 bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head -20
ONLY="normal 2^30 (size=64),a**2.7,a**b,1.3**b,log(,exp(,gather"
{ for v in "base" "VKP_PRNG_THREADS_PER_SM=1536" "VKP_PRNG_THREADS_PER_SM=2048" "VKP_PRNG_THREADS_PER_SM=3072" "VKP_PRNG_THREADS_PER_SM=768"; do
  echo "== $v"
  if [ "$v" = "base" ]; then python scripts/bench_all.py --only "$ONLY" 2>&1 | grep -E "GB/s"; else env $v python scripts/bench_all.py --only "normal 2^30 (size=64)" 2>&1 | grep -E "GB/s"; fi
done; } > gpurun_out/r02_variants_job10.txt 2>&1
cat gpurun_out/r02_variants_job10.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1_v2.json 2> gpurun_out/r02_bench_n1_v2.err
echo "bench exit $?"; python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_bench_n1_v2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['frac_of_nominal_8TBs'], d['clocks'])
print(d['roofline']['per_op_frac'])
print({k:(v['ms'],v['frac']) for k,v in d['configs']['C4']['rows'].items()})
print({k:(v['ms'],v['frac']) for k,v in d['configs']['C5']['rows'].items()})
print(d['configs']['C5']['mlp_step'])
print(d['e2e'])
P
tail -3 gpurun_out/r02_bench_n1_v2.err
