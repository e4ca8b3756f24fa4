"""One matrix-free expectation value of HOOH STO-3G (24 q, 14 905 terms, dense 2^24 state) per mode;
used under ncu to capture the apply kernels. Prints the timings."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import PauliwordOp, ops  # noqa: E402

d = np.load(os.path.join("tests", "golden", "hamiltonians", "HOOH_STO3G.npz"))
n = int(d["n_qubits"][0])
symp = np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool)
H = PauliwordOp(symp, d["coeff"])
rng = np.random.default_rng(0)
psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
psi /= np.linalg.norm(psi)
psi_d = torch.from_numpy(psi).cuda()
xm, zm, cp = H._terms_sorted()
groups = int(torch.unique_consecutive(xm).numel())
high = int((xm >= 2048).sum())
print(f"n={n} terms={H.n_terms} groups={groups} terms_with_x>=2^11={high} hermitian_real={cp._sym_hermitian}")


def run(label, sym, variant):
    ops.use_symmetric_expval = sym
    ops.set_tuning(4, variant)
    vals = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e = ops.expval_dense(xm, zm, cp, n, psi_d)
        e1.record()
        torch.cuda.synchronize()
        vals.append(e0.elapsed_time(e1))
    print(f"{label}: {min(vals):.2f} ms  expval={complex(e.cpu().numpy()):.12f}")


run("binned symmetric", True, 1)
run("binned", False, 1)
run("4-row", False, 0)
ops.set_tuning(4, 1)
ops.use_symmetric_expval = True
