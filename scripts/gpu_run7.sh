#!/bin/bash
# Final one-GPU artefacts of the round: tests, smoke, full bench (both arms), secondary paths, ncu launch list,
# ncu --set full of the emit kernel (traffic) and of one radix scatter pass.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r7_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r7_pytest.log
tail -6 gpurun_out/r7_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r7_smoke.log 2>&1; tail -2 gpurun_out/r7_smoke.log
timeout 400 python bench.py > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err
tail -c 2500 gpurun_out/r7_bench.json; tail -3 gpurun_out/r7_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r7_bench_reference.json 2> gpurun_out/r7_bench_reference.err
cat gpurun_out/r7_bench_reference.json
timeout 600 python scripts/bench_paths.py > gpurun_out/r7_paths.json 2> gpurun_out/r7_paths.err
cat gpurun_out/r7_paths.json; tail -3 gpurun_out/r7_paths.err
timeout 200 python scripts/probe_expval.py > gpurun_out/r7_expval.txt 2>&1; cat gpurun_out/r7_expval.txt
SYMMER_BENCH_QUICK=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r7_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r7_ncu_launch.log 2>&1
SYMMER_BENCH_QUICK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"emit_kernel|rs_scatter" -s 12 -c 6 \
    -o gpurun_out/r7_hot python bench.py --steps 2 --warmup 1 > gpurun_out/r7_ncu_hot.log 2>&1
tail -2 gpurun_out/r7_ncu_hot.log
ls -la gpurun_out | grep r7_
