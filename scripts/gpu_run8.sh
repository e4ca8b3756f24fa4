#!/bin/bash
# Ordered-tile mode: GPU tests, A/B bench (sorted-hash order vs tiles, B rows per CTA), launch list, ncu of the tile kernel.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r8_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r8_pytest.log
tail -8 gpurun_out/r8_pytest.log
for t in "6=0" "6=1" "6=1,7=8" "6=1,7=32"; do
  SYMMER_BENCH_QUICK=1 SYMMER_TUNING="$t" timeout 300 python bench.py --steps 10 --warmup 3 > "gpurun_out/r8_bench_$t.json" 2> "gpurun_out/r8_bench_$t.err"
  python - "$t" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r8_bench_{t}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(t, "ms/step", round(d["ms_per_step"], 3), "emit kernel ms", round(r["kernel_ms"], 3), "frac", round(r["frac"], 3),
          "phase ms", round(r["emit_phase_ms"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print(t, "FAILED", e)
    print(open(f"gpurun_out/r8_bench_{t}.err").read()[-2000:])
PY
done
SYMMER_BENCH_QUICK=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/r8_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r8_ncu_launch.log 2>&1
SYMMER_BENCH_QUICK=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"tile_emit_kernel|sum_kernel" -s 4 -c 2 \
    -o gpurun_out/r8_tile python bench.py --steps 2 --warmup 1 > gpurun_out/r8_ncu_tile.log 2>&1
tail -2 gpurun_out/r8_ncu_tile.log
ls -la gpurun_out | grep r8_
