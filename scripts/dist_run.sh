#!/bin/bash
# usage: scripts/dist_run.sh N  -- multi-GPU parity + timing run (one process per GPU)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  scripts/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1
echo "dist_check exit $?"; tail -c 4000 gpurun_out/dist_check_n$N.log
