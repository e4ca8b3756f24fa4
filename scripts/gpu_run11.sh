#!/bin/bash
# final validation call of the round: the whole GPU suite (-x, like the driver), then smoke()
set -u
mkdir -p gpurun_out
TAG=${TAG:-r14}
timeout 120 python -m pytest tests -m gpu -x -q --durations=4 > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -v "Greedy\|warnings.warn\|^$\|Docs:\|warnings summary" gpurun_out/${TAG}_pytest.log | tail -14
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
tail -3 gpurun_out/${TAG}_smoke.log
