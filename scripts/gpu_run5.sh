#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r5_pytest.log
tail -8 gpurun_out/r5_pytest.log
timeout 300 python scripts/probe_expval.py > gpurun_out/r5_expval.txt 2>&1
cat gpurun_out/r5_expval.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:applyw -c 1 -o gpurun_out/r5_applyw \
    python scripts/probe_expval.py > gpurun_out/r5_ncu_applyw.log 2>&1
tail -2 gpurun_out/r5_ncu_applyw.log
