#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_18}
timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
