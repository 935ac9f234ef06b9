#!/bin/bash
# round 2, 1-GPU job 26: config-5 MLP step: device / host time, launch list with tensor-pipe %, per-launch order
mkdir -p gpurun_out
STEPS=200 python scripts/mlp_profile.py > gpurun_out/r02_mlp_profile.txt 2>&1; head -3 gpurun_out/r02_mlp_profile.txt
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__inst_executed_pipe_tensor.sum
STEPS=1 timeout 300 ncu --metrics $M --clock-control none -s 140 -c 80 --csv --log-file gpurun_out/r02_mlp_launches.csv python scripts/mlp_profile.py > /dev/null 2>&1
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r02_mlp_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); gi=h.index('Grid Size'); idc=h.index('ID'); bs=h.index('Block Size')
d={}
for r in rows[hi+1:]:
    d.setdefault(r[idc],{'k':r[kn].split('(')[0].replace('void ','').replace('<unnamed>::',''),'g':r[gi],'b':r[bs]})[r[mn]]=r[mv]
for i,v in d.items():
    print(i, v['k'][:60], v['g'], v['b'], v.get('gpu__time_duration.sum'), 'tensor%', v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'), 'dram%', v.get('dram__throughput.avg.pct_of_peak_sustained_elapsed'))
P
