#!/bin/bash
mkdir -p gpurun_out
{ echo "== default"; python scripts/prng_size_sweep.py; echo "== VKP_PRNG_STARTS=0"; VKP_PRNG_STARTS=0 python scripts/prng_size_sweep.py; } > gpurun_out/r02_prng_size_sweep.txt 2>&1
cat gpurun_out/r02_prng_size_sweep.txt
