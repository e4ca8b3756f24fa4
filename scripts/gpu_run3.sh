#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r3_pytest.log
tail -30 gpurun_out/r3_pytest.log
timeout 300 python scripts/bench_paths.py gf2 > gpurun_out/r3_paths.json 2> gpurun_out/r3_paths.err
cat gpurun_out/r3_paths.json; tail -3 gpurun_out/r3_paths.err
SYMMER_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err
tail -c 1500 gpurun_out/r3_bench.json; tail -3 gpurun_out/r3_bench.err
