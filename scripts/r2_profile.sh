#!/bin/bash
# ncu evidence of the round: launch list of one bench step, full captures of the hot kernels
set -u
mkdir -p gpurun_out
T=${TAG:-r2_prof}
if [ "${ONLY_HOT:-0}" != "1" ]; then
SYMMER_BENCH_QUICK=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv 8 2>&1 | tail -24
PROBE_ONLY=class32-prefilter PROBE_SPAN_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"class_dedup|group_kernel|tile_emit_kernel" -s 9 -c 3 \
    -o gpurun_out/${T}_span python scripts/probe_class.py > gpurun_out/${T}_ncu_span.log 2>&1
tail -1 gpurun_out/${T}_ncu_span.log
fi
SYMMER_BENCH_QUICK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tile_emit_kernel|class_dedup" -s 8 -c 2 \
    -o gpurun_out/${T}_hot python bench.py --steps 2 --warmup 1 > gpurun_out/${T}_ncu_hot.log 2>&1
tail -1 gpurun_out/${T}_ncu_hot.log
