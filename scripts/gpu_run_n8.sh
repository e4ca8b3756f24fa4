#!/bin/bash
# N-GPU check of the bench launched as the driver does (N = number of visible GPUs); at N = 8 this is config C5.
set -u
mkdir -p gpurun_out
TAG=${TAG:-n8}
N=$(nvidia-smi -L | wc -l)
SYMMER_BENCH_QUICK=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_$N.json 2> gpurun_out/${TAG}_bench_$N.err
python - "$TAG" "$N" <<'PY'
import json, sys
tag, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/{tag}_bench_{n}.json").read().strip().splitlines()[-1])
    print(n, "GPUs: ms/step", round(d["ms_per_step"], 3), "value", f"{d['value']:.4g}", "e2e ms", round(d["e2e"]["ms_per_step"], 3),
          "emit ms", round(d["roofline"]["kernel_ms"], 3), "U", d["config"]["unique_terms_total"], "T", d["config"]["cross_terms_total"])
except Exception as e:
    print(n, "FAILED", e)
    print(open(f"gpurun_out/{tag}_bench_{n}.err").read()[-3000:])
PY
