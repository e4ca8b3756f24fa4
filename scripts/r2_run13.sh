#!/bin/bash
set -u
mkdir -p gpurun_out
T=${TAG:-r2_13}
timeout 500 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
tail -c 400 gpurun_out/${T}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"],3), "value %.4g" % d["value"], "e2e ms", round(d["e2e"]["ms_per_step"],3), "emit ms", round(d["roofline"]["kernel_ms"],3), "frac_out_bound", round(d["roofline"]["step_frac_of_output_write_bound"],3), "launches", d["gpu_launches"])
    print("check", d["check"])
    for k in ("secondary_dedup","e2e_host_result","c5_streamed","secondary","cpu_baseline"):
        v = d.get(k)
        print(k, {kk: vv for kk, vv in (v or {}).items() if kk in ("value","ms_per_step","ms_per_pass","unique_terms","model_frac_of_hbm_peak","output_rows_distinct","parity_vs_oracle","survivors","cpu_baseline","d2h_bytes_per_step")})
except Exception as e:
    print("parse failed", e)
PY
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>&1; tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
