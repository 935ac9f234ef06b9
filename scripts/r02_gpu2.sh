#!/bin/bash
# round 2, multi-GPU job (gpurun --gpus N): parity of the sharded path incl. the peer mailbox, the bench
# line with its `sharded` block, compute-sanitizer over the pull kernel / mailbox kernel on 2 ranks.
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_check.py > gpurun_out/r02_dist_check_n$N.log 2>&1
echo "dist_check exit $?"; tail -c 2500 gpurun_out/r02_dist_check_n$N.log
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
echo "bench exit $?"; tail -c 6000 gpurun_out/r02_bench_n$N.json; grep -v "NCCL INFO" gpurun_out/r02_bench_n$N.err | tail -15
if [ "$N" = "2" ]; then
  for tool in memcheck racecheck; do
    timeout 400 $TR --master-port 29513 --no-python compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_targets.py pull \
      > gpurun_out/r02_sanitize_${tool}_pull.log 2>&1
    echo "sanitize $tool pull exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pull ok|Error|error" gpurun_out/r02_sanitize_${tool}_pull.log | head -8
  done
fi
