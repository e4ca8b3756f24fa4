#!/bin/bash
# round 2, call 2: class-local dedup — new parity tests, the existing product tests, quick bench, collision probe
set -u
mkdir -p gpurun_out
T=r2_02
timeout 900 python -m pytest tests/test_gpu_class_dedup.py -m gpu -x -q > gpurun_out/${T}_class.log 2>&1; echo "exit $?" >> gpurun_out/${T}_class.log
tail -25 gpurun_out/${T}_class.log
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "tile or blocks or record or bench_size or owner or exchange or mul or square" > gpurun_out/${T}_ops.log 2>&1; echo "exit $?" >> gpurun_out/${T}_ops.log
tail -8 gpurun_out/${T}_ops.log
SYMMER_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_02_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "emit ms", d["roofline"]["kernel_ms"], "launches", d["gpu_launches"], "U", d["config"]["unique_terms_total"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/r2_02_bench.err").read()[-2000:])
PY
timeout 600 python scripts/probe_collisions.py > gpurun_out/${T}_collisions.json 2> gpurun_out/${T}_collisions.err
cat gpurun_out/${T}_collisions.json; tail -3 gpurun_out/${T}_collisions.err
