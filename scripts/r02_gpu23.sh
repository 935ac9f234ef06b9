#!/bin/bash
mkdir -p gpurun_out
{ echo "== STARTS=1"; python scripts/prng_probe2.py; echo "== STARTS=0"; VKP_PRNG_STARTS=0 python scripts/prng_probe2.py; } > gpurun_out/r02_prng_probe2.txt 2>&1
cat gpurun_out/r02_prng_probe2.txt
