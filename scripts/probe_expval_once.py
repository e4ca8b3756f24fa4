"""One symmetric-mode expectation value of HOOH STO-3G (for an ncu capture of applyw_kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import PauliwordOp, ops
dev = ops.device()
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "hamiltonians", "HOOH_STO3G.npz"))
n = int(d["n_qubits"][0])
H = PauliwordOp(np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool), d["coeff"])
rng = np.random.default_rng(0)
psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
psi /= np.linalg.norm(psi)
psi_d = torch.from_numpy(psi).to(dev)
xm, zm, cp = H._terms_sorted()
for _ in range(2):
    v = ops.expval_dense(xm, zm, cp, n, psi_d)
print(complex(v.item()))
