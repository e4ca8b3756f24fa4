import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from oracle import pauli_oracle as po
ops.device()
for n, M, N in [(64, 128, 256), (100, 130, 300), (1000, 257, 513), (36, 700, 513), (1000, 40, 33), (1, 5, 7), (1100, 70, 50)]:
    a_s, _ = po.random_operator(n, M, seed=n); b_s, _ = po.random_operator(n, N, seed=n + 1)
    a = ops.pack(torch.from_numpy(a_s), n); b = ops.pack(torch.from_numpy(b_s), n)
    ref = ops.commute(a, b)
    got = ops.commute_mma(a, b)
    torch.cuda.synchronize()
    ok = bool(torch.equal(ref, got))
    print(n, M, N, "match" if ok else f"MISMATCH {(ref != got).sum().item()} of {M*N}", flush=True)
    if not ok:
        d = (ref != got).nonzero()[:5]
        print(d.tolist())
n, M = 1000, 100000
a_s, _ = po.random_operator(n, M, seed=3)
a = ops.pack(torch.from_numpy(a_s), n)
blk = a[:20000].contiguous()
for name, fn in [("packed", ops.commute), ("mma", ops.commute_mma)]:
    for _ in range(2): fn(blk, a)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); o = fn(blk, a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del o
    print(name, f"{min(ts):.3f} ms  {20000*M/min(ts)*1e3:.3e} pairs/s", flush=True)
