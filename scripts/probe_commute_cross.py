import sys
import torch
sys.path.insert(0, ".")
import symmer_b200.ops as ops
from oracle import pauli_oracle as po
ops.device()
for n in [64, 128, 256, 512, 1000, 2000]:
    M = 40000
    a_s, _ = po.random_operator(n, M, seed=3)
    a = ops.pack(torch.from_numpy(a_s), n)
    blk = a[:16384].contiguous()
    res = {}
    for name, fn in [("packed", ops.commute), ("mma", ops.commute_mma)]:
        for _ in range(2): fn(blk, a)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); o = fn(blk, a); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1)); del o
        res[name] = 16384 * M / min(ts) * 1e3
    print(f"n={n} W={(n+63)//64} packed {res['packed']:.3e} mma {res['mma']:.3e} pairs/s  ratio {res['mma']/res['packed']:.2f}", flush=True)
