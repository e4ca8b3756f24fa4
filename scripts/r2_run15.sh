#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_15}
timeout 200 python scripts/probe_owned.py > gpurun_out/${T}_owned.json 2> gpurun_out/${T}_owned.err; cat gpurun_out/${T}_owned.json; tail -3 gpurun_out/${T}_owned.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches.csv python scripts/probe_owned.py > gpurun_out/${T}_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv 6 2>&1 | tail -22
