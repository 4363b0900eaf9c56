#!/bin/bash
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for cfg in C4 C3; do
for v in "tex 0" "tex+linear 0" "tex+linear 1" "linear 0" "linear 1"; do
  set -- $v
  CRN_MIPS_TMA=$2 timeout 300 ncu --metrics $M --clock-control none -k regex:mip_chain -s 2 -c 2 --csv --log-file gpurun_out/mips_ab.csv python profiles/mips_ab.py $1 $cfg 2>&1 | grep -E "stage|ok|Error|assert" 
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/mips_ab.csv')))
hi=next(i for i,r in enumerate(rows) if r and r[0]=='ID')
h=rows[hi]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
d={}
for r in rows[hi+1:]:
    d.setdefault(r[0],{})[r[mn]]=(float(r[mv].replace(',','')), r[mu]); d[r[0]]['k']=r[kn]
for k,v in d.items():
    print("    ", v['k'][:40], {m:x for m,x in v.items() if m!='k'})
PY
done; done
