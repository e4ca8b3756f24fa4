#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_33_rot_launches.csv python scripts/probe_rot_fused.py 10000000 > gpurun_out/r2_33_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_33_rot_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki,vi,ui=hdr.index('Kernel Name'),hdr.index('Metric Value'),hdr.index('Metric Unit')
seq=[]
for r in rows[hi+1:]:
    if len(r)>vi:
        v=float(r[vi].replace(',','')); v = v/1e3 if r[ui]=='ns' else v
        seq.append((r[ki].split('(')[0].replace('void ','').replace('symb::','')[:60], v))
names=[n for n,_ in seq]
idx=[i for i,n in enumerate(names) if 'rotate_split' in n or 'rotate_info' in n]
print(len(seq), idx[-4:])
a=idx[-1]
while a>0 and ('rotate' in names[a-1]): a-=1
for n,v in seq[a:]: print(f"{v:9.1f} {n}")
print("sum", sum(v for _,v in seq[a:]))
PY
