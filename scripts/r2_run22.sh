#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_22}
timeout 500 python scripts/bench_paths.py > gpurun_out/${T}_paths.json 2> gpurun_out/${T}_paths.err; echo "paths exit $?"
cat gpurun_out/${T}_paths.json | cut -c1-420; tail -3 gpurun_out/${T}_paths.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_c1_launches.csv python scripts/probe_c1.py > gpurun_out/${T}_c1.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_c1_launches.csv 8 2>&1 | tail -26
