"""One rank's share of config C5 at 8 GPUs on a single GPU: owned_product(A_full, B, log2_parts=3, owner=0) — the
8-block product the multi-GPU step runs per rank. For ncu launch lists and A/B timing."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from oracle import pauli_oracle as po
from symmer_b200 import ops, dist as sdist

dev = ops.device()
lg = int(os.environ.get("OWNED_LG", "3"))
G = 1 << lg
blocks = [po.random_operator(1000, 12500, seed=100 + r) for r in range(G)]
a = ops.pack(torch.from_numpy(np.vstack([b[0] for b in blocks])), 1000)
ac = torch.from_numpy(np.hstack([b[1] for b in blocks])).to(dev)
b_s, b_c = po.random_operator(1000, 10000, seed=7)
b, bc = ops.pack(torch.from_numpy(b_s), 1000), torch.from_numpy(b_c).to(dev)
a_part = sdist.partition_by_owner(a, ac, lg)
b_part = sdist.partition_by_owner(b, bc, lg)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(6):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    xz, c, info = sdist.owned_product(a, ac, b, bc, lg, 0, a_part=a_part, b_part=b_part)
    e1.record()
    torch.cuda.synchronize()
    if it >= 2:
        ts.append(e0.elapsed_time(e1))
    U = xz.shape[0]
    del xz, c, info
print(json.dumps({"path": f"owned_product, 1 of {G} parts of ({12500 * G} x 10000 terms)", "ms": float(np.mean(ts)), "U": U}))
