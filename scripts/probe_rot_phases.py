"""Phases of the fused rotation (rotate_split, block-list product count, emit) timed live with CUDA events and wall clock."""
import sys, math, time
import numpy as np, torch
sys.path.insert(0, ".")
from symmer_b200 import ops
import symmer_b200.ops as O
ops.device()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rng = np.random.default_rng(0)
xz = torch.from_numpy(rng.integers(-2**63, 2**63 - 1, size=(M, 32), dtype=np.int64)).cuda()
xz[:, 15] &= (1 << 40) - 1; xz[:, 31] &= (1 << 40) - 1
c = torch.from_numpy(rng.standard_normal(M) + 1j * rng.standard_normal(M)).cuda()
q = xz[:1].clone()
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(4):
    torch.cuda.synchronize(); w0 = time.perf_counter()
    e0 = ev()
    sxz, sc, sk, yc, n_comm = ops.rotate_split(xz, c, q)
    e1 = ev()
    W = 16
    b_xz = torch.zeros((3, 2 * W), dtype=torch.int64, device=xz.device); b_xz[2] = q.reshape(-1)
    b_c = torch.tensor([1.0, math.cos(0.7), -1j * math.sin(0.7)], dtype=torch.complex128, device=xz.device)
    e2 = ev()
    out = ops.mul_blocks_cleanup(sxz, sc, b_xz, b_c, [(0, n_comm, 0, 1), (n_comm, M, 1, 3)], 1e-15, a_sketch=sk, a_ycount=yc)
    e3 = ev()
    torch.cuda.synchronize(); w1 = time.perf_counter()
    if rep >= 2:
        print(f"split {e0.elapsed_time(e1):.3f}  small tensors {e1.elapsed_time(e2):.3f}  product {e2.elapsed_time(e3):.3f}  total {e0.elapsed_time(e3):.3f} ms  wall {(w1 - w0) * 1e3:.3f} ms", flush=True)
    del out, sxz, sc
