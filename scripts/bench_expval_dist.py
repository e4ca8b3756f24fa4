"""Config C4 on N GPUs: matrix-free <psi|H|psi> of HOOH STO-3G (24 q, 14 905 terms) with the 2^24 basis
rows sharded over the ranks (psi and the terms replicated) and one all-reduce of the partial sums.
Launch: python -m torch.distributed.run --nproc-per-node N scripts/bench_expval_dist.py. Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import PauliwordOp, ops  # noqa: E402
from symmer_b200 import dist as sdist  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = ops.device()
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
d = np.load(os.path.join("tests", "golden", "hamiltonians", "HOOH_STO3G.npz"))
n = int(d["n_qubits"][0])
H = PauliwordOp(np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool), d["coeff"])
rng = np.random.default_rng(0)
psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
psi /= np.linalg.norm(psi)
psi_d = torch.from_numpy(psi).to(dev)
xm, zm, cp = H._terms_sorted()
times = []
for it in range(6):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e = sdist.sharded_expval(xm, zm, cp, n, psi_d)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
t = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t.item())
    print(json.dumps({"path": "C4 matrix-free expval HOOH STO-3G, 2^24 basis sharded", "n_gpus": world, "ms": ms,
                      "sign_evals_per_s": (1 << n) * H.n_terms / (ms * 1e-3), "expval_real": e.real,
                      "mode": "symmetric" if getattr(cp, "_sym_hermitian", False) else "plain"}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
