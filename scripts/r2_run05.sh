#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_05
timeout 900 python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q -k medium > gpurun_out/${T}_class.log 2>&1; echo "exit $?" >> gpurun_out/${T}_class.log
tail -5 gpurun_out/${T}_class.log
PROBE_ONLY=class1024 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"class_dedup_kernel" -s 2 -c 1 \
    -o gpurun_out/${T}_class1024 python scripts/probe_class.py > gpurun_out/${T}_ncu1.log 2>&1
tail -2 gpurun_out/${T}_ncu1.log
PROBE_ONLY=class512 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"class_dedup_kernel" -s 2 -c 1 \
    -o gpurun_out/${T}_class512 python scripts/probe_class.py > gpurun_out/${T}_ncu0.log 2>&1
tail -2 gpurun_out/${T}_ncu0.log
