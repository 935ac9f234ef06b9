#!/bin/bash
# round 2, N-GPU job 11: the bench line with its `sharded` block exactly as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}_v6.json 2> gpurun_out/r02_bench_n${N}_v6.err
echo "bench exit $?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r02_bench_n${N}_v6.json').read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['e2e'])
    for k,v in d['sharded']['rows'].items(): print(k, v.get('ms'), v.get('x_over_one_gpu'), v.get('agg'))
except Exception as e: print('parse failed', e)
P
grep -v "NCCL INFO" gpurun_out/r02_bench_n${N}_v6.err | tail -12
