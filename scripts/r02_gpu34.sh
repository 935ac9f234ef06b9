#!/bin/bash
# round 2, N-GPU job 34: what bounds the fixed-size row-sharded matmul
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{ $TR --master-port 29531 scripts/mm_fused_probe.py
  VKP_TC_BN224=0 $TR --master-port 29532 scripts/mm_fused_probe.py
  VKP_COMM_NO_PULL=1 $TR --master-port 29533 scripts/mm_fused_probe.py
  VKP_COMM_NO_PULL=1 VKP_TC_BN224=0 $TR --master-port 29534 scripts/mm_fused_probe.py; } 2>&1 | grep "fused row-sharded" > gpurun_out/r02_mm_fused_probe_n$N.txt
cat gpurun_out/r02_mm_fused_probe_n$N.txt
