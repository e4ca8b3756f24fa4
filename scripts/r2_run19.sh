#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_19}
for M in 100000 300000 1000000 3000000 10000000; do timeout 120 python scripts/probe_rot_fused.py $M; done > gpurun_out/${T}_rot.txt 2>&1
cat gpurun_out/${T}_rot.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_rot_launches.csv python scripts/probe_rot_fused.py 10000000 > gpurun_out/${T}_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_rot_launches.csv 6 2>&1 | tail -22
