"""One large general rotation through both paths (for an ncu launch list / event timing)."""
import sys, math
import numpy as np
import torch
sys.path.insert(0, ".")
from symmer_b200 import PauliwordOp, ops
import symmer_b200.base as base
ops.device()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(0)
xz = torch.from_numpy(rng.integers(-2**63, 2**63 - 1, size=(M, 32), dtype=np.int64)).cuda()
xz[:, 15] &= (1 << 40) - 1; xz[:, 31] &= (1 << 40) - 1          # 1000 qubits: padding bits clear
c = torch.from_numpy(rng.standard_normal(M) + 1j * rng.standard_normal(M)).cuda()
q = xz[:1].clone()
for name, fn in [("two-step", lambda: ops.cleanup(*ops.rotate(xz, c, q, math.cos(0.7), math.sin(0.7), 0))),
                 ("fused", lambda: ops.rotate_dedup(xz, c, q, math.cos(0.7), math.sin(0.7)))]:
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    print(name, M, "->", out[0].shape[0], f"{e0.elapsed_time(e1):.3f} ms", flush=True)
    del out
