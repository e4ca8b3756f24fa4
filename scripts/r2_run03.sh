#!/bin/bash
# round 2, call 3: class-local dedup — tests after the ordering fix, launch list and ncu of the class kernel
set -u
mkdir -p gpurun_out
T=r2_03
timeout 900 python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q > gpurun_out/${T}_class.log 2>&1; echo "exit $?" >> gpurun_out/${T}_class.log
tail -15 gpurun_out/${T}_class.log
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "tile or blocks or record or bench_size or owner or exchange or mul or square" > gpurun_out/${T}_ops.log 2>&1; echo "exit $?" >> gpurun_out/${T}_ops.log
tail -8 gpurun_out/${T}_ops.log
SYMMER_BENCH_QUICK=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv 2>&1 | tail -30
SYMMER_BENCH_QUICK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"class_dedup_kernel" -s 2 -c 1 \
    -o gpurun_out/${T}_class python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_class.log 2>&1
tail -2 gpurun_out/${T}_ncu_class.log
