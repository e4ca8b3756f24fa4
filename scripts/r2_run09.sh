#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_09}
cat > /tmp/span_only.py <<'PY'
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.getcwd())
sys.argv = [sys.argv[0]]
import scripts.probe_class as pc
PY
PROBE_ONLY=class1024 PROBE_SPAN_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"class_dedup_kernel|group_kernel" -s 6 -c 2 \
    -o gpurun_out/${T}_span python scripts/probe_class.py > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
