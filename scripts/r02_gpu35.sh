#!/bin/bash
# round 2, N-GPU job 35: fused row-sharded matmul after doubling the pull's loads in flight (one configuration)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29541 scripts/mm_fused_probe.py 2>&1 | grep "fused row-sharded" > gpurun_out/r02_mm_fused_probe_u16_n$N.txt
cat gpurun_out/r02_mm_fused_probe_u16_n$N.txt
