#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_28
timeout 240 python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q --timeout 200 > gpurun_out/${T}_class.log 2>&1; echo "exit $?" >> gpurun_out/${T}_class.log
tail -3 gpurun_out/${T}_class.log
timeout 150 python scripts/probe_owned.py
PROBE_ONLY=class32 timeout 150 python scripts/probe_class.py
