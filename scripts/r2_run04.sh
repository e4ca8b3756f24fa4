#!/bin/bash
set -u
mkdir -p gpurun_out
T=r2_04
timeout 900 python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q > gpurun_out/${T}_class.log 2>&1; echo "exit $?" >> gpurun_out/${T}_class.log
tail -5 gpurun_out/${T}_class.log
timeout 600 python scripts/probe_class.py > gpurun_out/${T}_probe.json 2> gpurun_out/${T}_probe.err
cat gpurun_out/${T}_probe.json; tail -3 gpurun_out/${T}_probe.err
PROBE_ONLY=class timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"class_dedup|link_work|sum_work|phase_work|cd_|os_pass|tile_classify" -c 200 --csv \
    --log-file gpurun_out/${T}_launches.csv python scripts/probe_class.py > gpurun_out/${T}_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv 1 2>&1 | tail -14
