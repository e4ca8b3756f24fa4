#!/bin/bash
# One-GPU round-trip: parity tests, bench, secondary paths, ncu launch list + full capture of the emit kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_pytest.log
tail -5 gpurun_out/r1_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
tail -c 3000 gpurun_out/r1_bench.json
timeout 400 python scripts/bench_paths.py c1 expval rotate gf2 > gpurun_out/r1_paths.json 2> gpurun_out/r1_paths.err
cat gpurun_out/r1_paths.json
# old two-kernel emit for an A/B on the same box
timeout 200 python - > gpurun_out/r1_ab.txt 2>&1 <<'PY'
import json, subprocess, sys, os
for variant in (0, 1):
    env = dict(os.environ, SYMMER_BENCH_QUICK="1", SYMMER_EMIT_VARIANT=str(variant))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print("emit_variant", variant, "ms_per_step", j["ms_per_step"], "emit_ms", j["roofline"]["kernel_ms"])
    except Exception as e:
        print("variant", variant, "failed", e, out.stderr[-500:])
PY
cat gpurun_out/r1_ab.txt
SYMMER_BENCH_QUICK=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1_ncu_launch.log 2>&1
SYMMER_BENCH_QUICK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:emit_fused -s 3 -c 1 \
    -o gpurun_out/r1_emit python bench.py --steps 2 --warmup 1 > gpurun_out/r1_ncu_emit.log 2>&1
ls -la gpurun_out | tail -20
