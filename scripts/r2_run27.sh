#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_27
timeout 200 python scripts/probe_owned.py > gpurun_out/${T}_owned.json 2> gpurun_out/${T}_owned.err; cat gpurun_out/${T}_owned.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"class_dedup" -s 3 -c 1 -o gpurun_out/${T}_owned python scripts/probe_owned.py > gpurun_out/${T}_ncu.log 2>&1
tail -1 gpurun_out/${T}_ncu.log
