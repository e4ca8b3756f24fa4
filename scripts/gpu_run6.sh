#!/bin/bash
# 8-GPU box: NCCL parity (world 4), weak-scaling bench at 8 / 4 / 2 / 1, sharded expval at 8 / 1.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r6_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r6_pytest.log
tail -15 gpurun_out/r6_pytest.log
for n in 8 4 2; do
  SYMMER_BENCH_QUICK=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r6_bench$n.json 2> gpurun_out/r6_bench$n.err
  tail -c 900 gpurun_out/r6_bench$n.json; echo
done
SYMMER_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r6_bench1.json 2> gpurun_out/r6_bench1.err
tail -c 900 gpurun_out/r6_bench1.json; echo
SYMMER_BENCH_QUICK=1 SYMMER_DIST_METHOD=alltoall timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r6_bench8_alltoall.json 2> gpurun_out/r6_bench8_alltoall.err
tail -c 600 gpurun_out/r6_bench8_alltoall.json; echo
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
      scripts/bench_expval_dist.py > gpurun_out/r6_expval8.json 2> gpurun_out/r6_expval8.err
cat gpurun_out/r6_expval8.json
timeout 200 python scripts/bench_expval_dist.py > gpurun_out/r6_expval1.json 2> gpurun_out/r6_expval1.err
cat gpurun_out/r6_expval1.json
