#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_config_size.py -m gpu -x -q --timeout 150 -k "commute" > gpurun_out/r2_30_pytest.log 2>&1; echo "exit $?" >> gpurun_out/r2_30_pytest.log; tail -3 gpurun_out/r2_30_pytest.log
timeout 90 python scripts/probe_commute_ws.py 2>&1 | cut -c1-330
