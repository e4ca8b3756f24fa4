#!/bin/bash
# quick loop: GPU tests + A/B bench lines (SYMMER_TUNING sets) + launch list
set -u
mkdir -p gpurun_out
TAG=${TAG:-r9}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
IFS=';' read -ra SETS <<< "${TUNINGS:-6=1}"
for t in "${SETS[@]}"; do
  SYMMER_BENCH_QUICK=1 SYMMER_TUNING="$t" timeout 300 python bench.py --steps 10 --warmup 3 > "gpurun_out/${TAG}_bench_$t.json" 2> "gpurun_out/${TAG}_bench_$t.err"
  python - "$t" "$TAG" <<'PY'
import json, sys
t, tag = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/{tag}_bench_{t}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(t, "ms/step", round(d["ms_per_step"], 3), "emit kernel ms", round(r["kernel_ms"], 3), "frac", round(r["frac"], 3),
          "phase ms", round(r["emit_phase_ms"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"])
except Exception as e:
    print(t, "FAILED", e)
    print(open(f"gpurun_out/{tag}_bench_{t}.err").read()[-2000:])
PY
done
SYMMER_BENCH_QUICK=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${TAG}_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv 3 | head -24
