"""commute_mma: warp-specialised variant (knob 12 = 1) against the block-barrier variant — equality and pairs/s."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import ops
dev = ops.device()
g = torch.Generator(device=dev); g.manual_seed(3)
for W2, M, N in [(32, 16384, 65536), (32, 1000, 3001), (2, 20000, 42599), (8, 8192, 16384), (64, 4096, 8192)]:
    a = torch.randint(-2 ** 63, 2 ** 63 - 1, (M, W2), dtype=torch.int64, device=dev, generator=g)
    b = torch.randint(-2 ** 63, 2 ** 63 - 1, (N, W2), dtype=torch.int64, device=dev, generator=g)
    res = {}
    for variant in (0, 1, 2, 3, 4):
        ops.set_tuning(12, variant)
        out = ops.commute_mma(a, b)
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); o2 = ops.commute_mma(a, b); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1)); del o2
        res[variant] = (out, min(ts))
    same = all(bool(torch.equal(res[0][0], res[v][0])) for v in res)
    print(json.dumps({"words": W2, "M": M, "N": N, "equal": same, **{f"pairs_per_s_v{v}": M * N / (res[v][1] * 1e-3) for v in res}}), flush=True)
    assert same
ops.set_tuning(12, 2)
