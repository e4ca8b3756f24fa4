import sys
import torch
sys.path.insert(0, ".")
import scripts.probe_mul as p
p.ops.device()
for extra in [5, -3, 1]:
    p.ops.set_tuning(3, extra)
    print("sort_extra_bits", extra)
    p.run(1000, 500, 500, square=True)
    p.run(1000, 12500, 10000, reps=3)
