#!/bin/bash
# round 2, 2-GPU job 17: lazy PRNG skip (folded into the draw) -- tests, then the N=2 bench line with the sharded block
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_random.py tests/test_gpu_large.py -m gpu -q --timeout 600 > gpurun_out/r02_pytest_prng2.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/r02_pytest_prng2.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02_pytest_prng2.log | head
bash scripts/r02_gpu11.sh 2
