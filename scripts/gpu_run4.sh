#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "binned or gf2 or apply" > gpurun_out/r4_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r4_pytest.log
tail -8 gpurun_out/r4_pytest.log
timeout 300 python scripts/probe_expval.py > gpurun_out/r4_expval.txt 2>&1
cat gpurun_out/r4_expval.txt
timeout 300 python scripts/bench_paths.py gf2 expval > gpurun_out/r4_paths.json 2> gpurun_out/r4_paths.err
cat gpurun_out/r4_paths.json; tail -3 gpurun_out/r4_paths.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:applyw -c 2 -o gpurun_out/r4_applyw \
    python scripts/probe_expval.py > gpurun_out/r4_ncu_applyw.log 2>&1
tail -3 gpurun_out/r4_ncu_applyw.log
ls -la gpurun_out | tail -8
