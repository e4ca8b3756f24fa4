#!/bin/bash
# One-GPU artefacts of the round: tests, smoke, full bench (both arms), secondary paths, ncu launch list,
# ncu --set full of the tile emission kernel (traffic) and of one radix pass.
set -u
mkdir -p gpurun_out
TAG=${TAG:-rf}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
cat gpurun_out/${TAG}_bench_reference.json
timeout 600 python scripts/bench_paths.py > gpurun_out/${TAG}_paths.json 2> gpurun_out/${TAG}_paths.err
cat gpurun_out/${TAG}_paths.json; tail -3 gpurun_out/${TAG}_paths.err
SYMMER_BENCH_QUICK=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_launch.log 2>&1
SYMMER_BENCH_QUICK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tile_emit_kernel|os_pass_kernel" -s 15 -c 5 \
    -o gpurun_out/${TAG}_hot python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_hot.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_hot.log
ls -la gpurun_out | grep ${TAG}_
