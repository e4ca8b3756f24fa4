#!/bin/bash
set -u
mkdir -p gpurun_out
PROBE_ONLY=class32 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"class_dedup" -s 2 -c 1 \
    -o gpurun_out/r2_25_class32 python scripts/probe_class.py > gpurun_out/r2_25_ncu.log 2>&1
tail -1 gpurun_out/r2_25_ncu.log
