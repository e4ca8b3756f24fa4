#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_16}
timeout 200 python scripts/probe_owned.py > gpurun_out/${T}_owned.json 2> gpurun_out/${T}_owned.err; cat gpurun_out/${T}_owned.json; tail -3 gpurun_out/${T}_owned.err
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_class_dedup.py -m gpu -x -q --timeout 300 -k "tile or blocks or record or bench_size or owner or exchange or mul or square or class or span" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
