"""Matrix-free ground-state energies (symmer_utils.exact_gs_energy on a PauliwordOp): every Lanczos matvec is the
device kernel sym_apply; the CSR matrix of the reference's path is never built. Prints one JSON line per molecule."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from symmer_b200 import PauliwordOp, ops  # noqa: E402
from symmer_b200 import symmer_utils as su  # noqa: E402


def main():
    ops.device()
    for tag in sys.argv[1:] or ["H2O_STO3G", "NH3_STO3G"]:
        d = np.load(os.path.join("tests", "golden", "hamiltonians", tag + ".npz"))
        n = int(d["n_qubits"][0])
        symp = np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool)
        H = PauliwordOp(symp, d["coeff"])
        calls = [0]
        orig = H.apply_dense

        def counted(psi, *a, **k):
            calls[0] += 1
            return orig(psi, *a, **k)

        H.apply_dense = counted
        su.exact_gs_energy(H)                                  # warm-up (allocator, term tables)
        calls[0] = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e, psi = su.exact_gs_energy(H)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dense = psi.to_dense_device()
        resid = float(torch.linalg.vector_norm(orig(dense) - e * dense).cpu())
        print(json.dumps({"path": f"matrix-free exact_gs_energy {tag}", "n_qubits": n, "n_terms": H.n_terms, "energy": float(e),
                          "hf_energy": float(d["hf_energy"][0]), "residual_norm": resid, "wall_s": dt, "matvecs": calls[0],
                          "csr_nnz_of_the_reference_path": int((1 << n) * len(np.unique(symp[:, :n], axis=0)))}), flush=True)


if __name__ == "__main__":
    main()
