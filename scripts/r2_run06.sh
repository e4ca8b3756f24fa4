#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_06}
timeout ${PYTEST_TIMEOUT:-240} python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q > gpurun_out/${T}_class.log 2>&1; echo "exit $?" >> gpurun_out/${T}_class.log
tail -4 gpurun_out/${T}_class.log
timeout ${PROBE_TIMEOUT:-150} python scripts/probe_class.py > gpurun_out/${T}_probe.json 2> gpurun_out/${T}_probe.err
cat gpurun_out/${T}_probe.json; tail -3 gpurun_out/${T}_probe.err
PROBE_ONLY=class timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"class_dedup|link_work|sum_work|phase_work|group" -c 100 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/probe_class.py > gpurun_out/${T}_ncu_launch.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${T}_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[hi]; ki,vi=hdr.index('Kernel Name'),hdr.index('Metric Value')
seen={}
for r in rows[hi+1:]:
    if len(r)>vi:
        seen.setdefault(r[ki][:48],[]).append(float(r[vi].replace(',',''))/1e6)
for k,v in seen.items(): print(k, ' '.join(f'{x:.3f}' for x in v[2::6][:8]))
PY
PROBE_ONLY=${NCU_VARIANT:-class32} timeout 200 ncu --set full --clock-control none --import-source on -k regex:"class_dedup" -s 2 -c 1 \
    -o gpurun_out/${T}_classk python scripts/probe_class.py > gpurun_out/${T}_ncu1.log 2>&1
tail -1 gpurun_out/${T}_ncu1.log
