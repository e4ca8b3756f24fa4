"""Golden vectors of the tapering workflow from the REAL reference:
QubitTapering(H).taper_it(ref_state=HF) — symmetry generators, sector, Clifford rotations, rotated
stabilizers and the tapered operator. Test infrastructure only; run in the build container:

    python tests/golden/make_golden_taper.py

Reference code exercised: projection/qubit_tapering.py:9-106, projection/base.py:44-124,
operators/independent_op.py:90-314 (through oracle/shim for the uninstalled third-party packages).
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

from symmer import PauliwordOp, QubitTapering  # noqa: E402
from symmer import process  # noqa: E402

process.method = 'single_thread'
HAM_DIR = "/root/reference/tests/hamiltonian_data"
out = {}


def put(name, **arrays):
    for k, v in arrays.items():
        out[f"{name}/{k}"] = np.asarray(v)


for fname in ["H2O_STO-3G_SINGLET_JW.json", "Be_STO-3G_SINGLET_JW.json", "NH3_STO-3G_SINGLET_JW.json",
              "HOOH_STO-3G_SINGLET_JW.json"]:
    with open(os.path.join(HAM_DIR, fname)) as f:
        d = json.load(f)
    H = PauliwordOp.from_dictionary({k: complex(v[0], v[1]) for k, v in d["hamiltonian"].items()})
    hf = np.asarray(d["data"]["hf_array"])
    tag = fname.replace("_SINGLET_JW.json", "").replace("-", "")
    for sqp in ["Z", "X"]:
        qt = QubitTapering(H, target_sqp=sqp)
        tapered = qt.taper_it(ref_state=hf)
        rot = np.array([r.symp_matrix[0] for r, _ in qt.stabilizers.stabilizer_rotations]).reshape(-1, 2 * H.n_qubits)
        put(f"taper_{tag}_{sqp}", gen_symp=qt.symmetry_generators.symp_matrix, sector=qt.stabilizers.coeff_vec.real,
            rotations=rot, rotated_symp=qt.rotated_stabilizers.symp_matrix, rotated_coeff=qt.rotated_stabilizers.coeff_vec.real,
            free=qt.free_qubit_indices, out_symp=np.packbits(tapered.symp_matrix, axis=1), out_coeff=tapered.coeff_vec,
            n_out_qubits=np.array([tapered.n_qubits]), hf=hf)
        print(tag, sqp, H.n_qubits, "->", tapered.n_qubits, "qubits,", H.n_terms, "->", tapered.n_terms, "terms")
    # a user-supplied sector instead of a reference state (all -1)
    qt = QubitTapering(H)
    sector = -np.ones(qt.n_taper, dtype=int)
    tapered = qt.taper_it(sector=sector)
    put(f"taper_{tag}_sector", sector=sector, out_symp=np.packbits(tapered.symp_matrix, axis=1), out_coeff=tapered.coeff_vec,
        n_out_qubits=np.array([tapered.n_qubits]))

np.savez_compressed(os.path.join(HERE, "taper_vectors.npz"), **out)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "taper_vectors.npz")) / 1e6, "MB")
