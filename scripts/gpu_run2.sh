#!/bin/bash
# Two-GPU round-trip: NCCL parity of the sharded paths, 2-GPU bench (owner vs all-to-all), emit variants A/B on one GPU.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest.log
tail -15 gpurun_out/r2_pytest.log
for method in owner alltoall; do
  SYMMER_DIST_METHOD=$method timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench2_$method.json 2> gpurun_out/r2_bench2_$method.err
  tail -c 1500 gpurun_out/r2_bench2_$method.json
done
for v in 0 1 2; do
  SYMMER_BENCH_QUICK=1 SYMMER_EMIT_VARIANT=$v timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_emit_v$v.json 2> gpurun_out/r2_emit_v$v.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_emit_v$v.json").read().strip().splitlines()[-1])
    print("emit_variant $v ms_per_step", j["ms_per_step"], "emit_ms", j["roofline"]["kernel_ms"])
except Exception as e:
    print("variant $v failed", e)
PY
done
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 2500 gpurun_out/r2_bench1.json
