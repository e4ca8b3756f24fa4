#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_14}
timeout 500 python -m pytest tests/test_gpu_config_size.py -m gpu -x -q --timeout 400 > gpurun_out/${T}_cfg.log 2>&1; echo "exit $?" >> gpurun_out/${T}_cfg.log
tail -12 gpurun_out/${T}_cfg.log
timeout 400 python scripts/bench_dist.py > gpurun_out/${T}_dist_n1.json 2> gpurun_out/${T}_dist_n1.err; echo "dist exit $?"
cat gpurun_out/${T}_dist_n1.json; tail -5 gpurun_out/${T}_dist_n1.err
