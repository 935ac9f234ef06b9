#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,dram__bytes_write.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct
for st in 1 0; do
VKP_PRNG_STARTS=$st timeout 300 ncu --metrics $M --clock-control none -k regex:'xoshiro' -s 0 -c 60 --csv --log-file gpurun_out/r02_prng_sweep_ncu_$st.csv python scripts/prng_size_sweep.py > /dev/null 2>&1
python - <<P
import csv
rows=list(csv.reader(open('gpurun_out/r02_prng_sweep_ncu_$st.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); gi=h.index('Grid Size'); idc=h.index('ID')
d={}
for r in rows[hi+1:]:
    d.setdefault(r[idc],{'k':r[kn].split('(')[0][-40:],'g':r[gi]})[r[mn]]=r[mv]
print('STARTS=$st')
for i,v in d.items():
    print(i, v['k'], v['g'], v.get('gpu__time_duration.sum'), v.get('smsp__inst_executed.sum'), v.get('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'), v.get('dram__bytes_write.sum'))
P
done
