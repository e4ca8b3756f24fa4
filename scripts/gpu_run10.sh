#!/bin/bash
# short validation call: the new API cases first, then the whole GPU suite, then a quick bench line
set -u
mkdir -p gpurun_out
TAG=${TAG:-r10}
timeout 200 python -m pytest tests/test_gpu_api_ext.py -q --durations=5 > gpurun_out/${TAG}_ext.log 2>&1
echo "ext exit $?" >> gpurun_out/${TAG}_ext.log
tail -25 gpurun_out/${TAG}_ext.log
timeout 330 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
SYMMER_BENCH_QUICK=1 timeout 120 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
