#!/bin/bash
# round 2, 8-GPU job: sharded parity + the bench line with its `sharded` block at N GPUs; two-shot vs one-shot mailbox A/B
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1
echo "dist_check exit $?"; grep "^{" gpurun_out/r02_dist_check_n$N.log | tail -1
VKP_COMM_TWO_SHOT_MIN=1000000000 timeout 600 $TR --master-port 29514 scripts/dist_check.py > gpurun_out/r02_dist_check_n${N}_oneshot.log 2>&1
echo "dist_check (one-shot only) exit $?"; grep "^{" gpurun_out/r02_dist_check_n${N}_oneshot.log | tail -1
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench exit $?"; tail -c 3000 gpurun_out/r02_bench_n$N.json; grep -v "NCCL INFO" gpurun_out/r02_bench_n$N.err | tail -12
