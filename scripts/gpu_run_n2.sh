#!/bin/bash
# 2-GPU check: NCCL parity test of the sharded paths, then bench.py at N=1 and N=2 launched as the driver does.
set -u
mkdir -p gpurun_out
TAG=${TAG:-n2}
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
SYMMER_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_1.json 2> gpurun_out/${TAG}_bench_1.err
SYMMER_BENCH_QUICK=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2.json 2> gpurun_out/${TAG}_bench_2.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/{tag}_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "GPUs: ms/step", round(d["ms_per_step"], 3), "value", f"{d['value']:.4g}", "e2e ms", round(d["e2e"]["ms_per_step"], 3),
              "emit ms", round(d["roofline"]["kernel_ms"], 3), "U", d["config"]["unique_terms_total"])
    except Exception as e:
        print(n, "FAILED", e)
        print(open(f"gpurun_out/{tag}_bench_{n}.err").read()[-3000:])
PY
