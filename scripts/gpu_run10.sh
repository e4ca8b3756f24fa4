#!/bin/bash
# short validation call: the new API cases first, then the whole GPU suite, then a quick bench line
set -u
mkdir -p gpurun_out
TAG=${TAG:-r10}
timeout 200 python -m pytest tests/test_gpu_api_ext.py -q --durations=5 > gpurun_out/${TAG}_ext.log 2>&1
echo "ext exit $?" >> gpurun_out/${TAG}_ext.log
tail -25 gpurun_out/${TAG}_ext.log
timeout 330 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
echo skip-paths > gpurun_out/${TAG}_paths.json; : > gpurun_out/${TAG}_paths.err
cat gpurun_out/${TAG}_paths.json; tail -5 gpurun_out/${TAG}_paths.err
