#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 90 python scripts/probe_commute_ws.py > gpurun_out/r2_20_commute.json 2> gpurun_out/r2_20_commute.err; echo "exit $?"
cat gpurun_out/r2_20_commute.json; tail -4 gpurun_out/r2_20_commute.err
