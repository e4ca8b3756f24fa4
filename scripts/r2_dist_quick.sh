#!/bin/bash
# quick multi-GPU regression: NCCL parity tests + the bench line as the driver launches it
set -u
N=${1:-2}
mkdir -p gpurun_out
T=r2q_n${N}
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "ms/step", round(d["ms_per_step"],3), "value %.4g" % d["value"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "U", d["config"]["unique_terms_total"], "check", d["check"]["all_ranks_ok"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${T}_bench.err").read()[-3000:])
PY
