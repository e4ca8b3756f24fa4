#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_21}
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_api.py tests/test_gpu_class_dedup.py -m gpu -x -q --timeout 200 -k "rotat or class or blocks" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
for M in 1000000 3000000 10000000 30000000; do timeout 120 python scripts/probe_rot_fused.py $M; done > gpurun_out/${T}_rot.txt 2>&1
cat gpurun_out/${T}_rot.txt
