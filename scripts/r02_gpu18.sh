#!/bin/bash
# round 2, 1-GPU job 18: full ncu capture of the uniform generator at 64 lanes vs 2^20 lanes (what limits the former)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'xoshiro_stream' -s 0 -c 4 \
  -o gpurun_out/r02_prng_full -f python scripts/r02_probe.py prng > gpurun_out/r02_prng_full.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/r02_prng_full.ncu-rep --page raw --csv > gpurun_out/r02_prng_full_raw.csv 2>/dev/null
wc -c gpurun_out/r02_prng_full_raw.csv
