#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q --timeout 300 -k "privatised or structured" > gpurun_out/r2_31_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_31_pytest.log; tail -15 gpurun_out/r2_31_pytest.log
