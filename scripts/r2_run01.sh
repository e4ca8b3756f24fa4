#!/bin/bash
# round 2, call 1: baseline of the duplicate-heavy product (28-generator span) with the round-1 engine + quick bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_01_smi.txt 2>&1
timeout 600 python scripts/probe_collisions.py > gpurun_out/r2_01_collisions.json 2> gpurun_out/r2_01_collisions.err
cat gpurun_out/r2_01_collisions.json; tail -3 gpurun_out/r2_01_collisions.err
SYMMER_BENCH_QUICK=1 timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_01_bench.json 2> gpurun_out/r2_01_bench.err
cat gpurun_out/r2_01_bench.json | head -c 1500; tail -3 gpurun_out/r2_01_bench.err
