#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/bench_dist.py > gpurun_out/r2_32_dist_n1.json 2> gpurun_out/r2_32_dist_n1.err; echo "dist exit $?"
cut -c1-200 gpurun_out/r2_32_dist_n1.json
