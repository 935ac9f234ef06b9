#!/bin/bash
# ncu evidence for the bench step: (1) launch list with per-launch device time, (2) one full
# capture of the kernels named in $1 (regex), read offline with `ncu -i ... --page raw --csv`.
mkdir -p gpurun_out
KREGEX=${1:-ew_kernel}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-matmul --e2e-steps 1 > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; wc -l gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 40 -c 14 -o gpurun_out/prof_full -f \
    python bench.py --steps 1 --warmup 3 --no-matmul --e2e-steps 1 > gpurun_out/bench_under_ncu_full.log 2>&1
echo "full capture exit $?"; ls -la gpurun_out/
