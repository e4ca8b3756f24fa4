"""Generate the golden vectors under tests/golden/ by running the REAL reference (UCL-CCS/symmer at
/root/reference) in the build container. Test infrastructure only.

    PYTHONPATH=oracle/shim:/root/reference python tests/golden/make_golden.py

The reference is imported unmodified; `oracle/shim` only provides stand-ins for third-party packages
that are not installed here (qiskit, openfermion, ray, quimb, matplotlib, cached_property) — see
DESIGN.md §3. Everything written here is an input/output pair of reference code:
  PauliwordOp.__mul__/_multiply_by_operator   base.py:764-859
  PauliwordOp.cleanup / symplectic_cleanup     base.py:617-638, utils.py:230-279
  commutes_termwise / adjacency_matrix         base.py:938-971, 1054-1062
  _rotate_by_single_Pword / perform_rotations  base.py:1090-1186
  to_sparse_matrix                             base.py:1458-1510
  _rref_binary / rref_binary / cref_binary     utils.py:292-359
  IndependentOp.symmetry_generators            independent_op.py:90-144
  generator_reconstruction                     base.py:523-560
The GPU box has no /root/reference, so the vectors (not this script's imports) are what travels.
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
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from symmer import PauliwordOp, QuantumState  # noqa: E402
from symmer.operators import IndependentOp  # noqa: E402
from symmer.operators.utils import _rref_binary, rref_binary, _cref_binary, cref_binary  # noqa: E402

HAM_DIR = "/root/reference/tests/hamiltonian_data"
out = {}


def put(name, **arrays):
    for k, v in arrays.items():
        out[f"{name}/{k}"] = np.asarray(v)


def load_ham(fname):
    with open(os.path.join(HAM_DIR, fname)) as f:
        d = json.load(f)
    ham = {k: complex(v[0], v[1]) for k, v in d["hamiltonian"].items()}
    return PauliwordOp.from_dictionary(ham), d["data"]


# --- 1. the reference's own golden cases (tests/test_operators/test_base.py) ----------------------
names = []
for a, b in [("X", "Y"), ("Y", "X"), ("Y", "Z"), ("Z", "Y"), ("Z", "X"), ("X", "Z"),     # :596-613
             ("XYZI", "ZZXY"), ("IIII", "YYYY"), ("XZYX", "XZYX")]:
    A = PauliwordOp.from_list([a], [1.0])
    B = PauliwordOp.from_list([b], [1.0])
    C = A * B
    nm = f"mul_single_{a}_{b}"
    put(nm, a_symp=A.symp_matrix, a_coeff=A.coeff_vec, b_symp=B.symp_matrix, b_coeff=B.coeff_vec,
        out_symp=C.symp_matrix, out_coeff=C.coeff_vec)
    names.append(nm)

P = PauliwordOp.from_list(["XXX", "YYY", "XXX", "YYY"], [1, 1, -1, 1])                    # :537-552
C = P.cleanup()
put("cleanup_ref_1", symp=P.symp_matrix, coeff=P.coeff_vec, out_symp=C.symp_matrix, out_coeff=C.coeff_vec)
P = PauliwordOp.from_list(["XXX", "YYY", "ZZZ"], [0, 0, 0])
C = P.cleanup()
put("cleanup_ref_zero", symp=P.symp_matrix, coeff=P.coeff_vec, out_symp=C.symp_matrix, out_coeff=C.coeff_vec)

for i, terms in enumerate([["XYXZ", "YYII", "YYZZ", "XIXZ"], ["IIII", "XXXX", "ZZZZ", "YIYI", "IZXY"]]):
    P = PauliwordOp.from_list(terms, np.ones(len(terms)))
    put(f"adj_ref_{i}", symp=P.symp_matrix, adj=P.adjacency_matrix)                       # :554-579
op1 = PauliwordOp.from_list(["XYXZ", "YYII"], [1, 1])
op2 = PauliwordOp.from_list(["YYZZ", "XIXZ", "XZZI"], [1, 1, 1])
put("commute_ref_doc", a_symp=op1.symp_matrix, b_symp=op2.symp_matrix, out=op1.commutes_termwise(op2))

for s in ["X", "Y", "Z", "XY", "ZY", "II", "XYZ", "YYI"]:                                 # :702-717
    P = PauliwordOp.from_list([s], [1.0])
    put(f"matrix_ref_{s}", symp=P.symp_matrix, coeff=P.coeff_vec, dense=P.to_sparse_matrix.toarray())

# --- 2. seeded random cases ---------------------------------------------------------------------
mul_cases = [(1, 3, 2), (2, 4, 4), (5, 20, 7), (8, 30, 30), (63, 12, 9), (64, 10, 13), (65, 9, 11),
             (130, 16, 5), (1000, 12, 8), (4, 1, 1), (6, 1, 17), (6, 17, 1)]
for k, (n, m1, m2) in enumerate(mul_cases):
    np.random.seed(100 + k)
    A = PauliwordOp.random(n, m1)
    B = PauliwordOp.random(n, m2)
    C = A * B
    put(f"mul_rand_{k}", a_symp=A.symp_matrix, a_coeff=A.coeff_vec, b_symp=B.symp_matrix,
        b_coeff=B.coeff_vec, out_symp=C.symp_matrix, out_coeff=C.coeff_vec)

# squares (heavy duplication + exact cancellation of anticommuting pairs)
for k, (n, m) in enumerate([(3, 20), (10, 40), (100, 30), (1000, 20)]):
    np.random.seed(200 + k)
    A = PauliwordOp.random(n, m)
    C = A * A
    put(f"square_rand_{k}", a_symp=A.symp_matrix, a_coeff=A.coeff_vec,
        out_symp=C.symp_matrix, out_coeff=C.coeff_vec)

# cleanup with many duplicates, thresholds
for k, (n, m, pool) in enumerate([(4, 200, 10), (70, 500, 37), (1000, 300, 100), (3, 64, 64)]):
    np.random.seed(300 + k)
    base = PauliwordOp.random(n, pool)
    idx = np.random.randint(0, pool, size=m)
    symp = base.symp_matrix[idx]
    coeff = np.random.randn(m) + 1j * np.random.randn(m)
    coeff[::7] = 0.0
    P = PauliwordOp(symp, coeff)
    C = P.cleanup()
    put(f"cleanup_rand_{k}", symp=symp, coeff=coeff, out_symp=C.symp_matrix, out_coeff=C.coeff_vec)

# add / sub
np.random.seed(350)
A = PauliwordOp.random(9, 40)
B = PauliwordOp(np.vstack([A.symp_matrix[:15], PauliwordOp.random(9, 10).symp_matrix]), np.random.randn(25))
put("add_rand", a_symp=A.symp_matrix, a_coeff=A.coeff_vec, b_symp=B.symp_matrix, b_coeff=B.coeff_vec,
    sum_symp=(A + B).symp_matrix, sum_coeff=(A + B).coeff_vec,
    diff_symp=(A - B).symp_matrix, diff_coeff=(A - B).coeff_vec)

# commute
for k, (n, m1, m2) in enumerate([(1, 4, 4), (7, 33, 65), (64, 40, 40), (65, 31, 50), (200, 64, 70),
                                 (1000, 50, 3), (1000, 37, 1)]):
    np.random.seed(400 + k)
    A = PauliwordOp.random(n, m1)
    B = PauliwordOp.random(n, m2)
    put(f"commute_rand_{k}", a_symp=A.symp_matrix, b_symp=B.symp_matrix, out=A.commutes_termwise(B))

# rotations: Clifford (k*pi/2 for k = 0..4) and general angles, single and sequences
rot_id = 0
for n, m in [(3, 12), (10, 60), (100, 50), (1000, 40)]:
    np.random.seed(500 + rot_id)
    P = PauliwordOp.random(n, m).cleanup()
    Q = PauliwordOp.random(n, 1)
    Q.coeff_vec[0] = 1
    for ang in [None, 0.0, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi, -np.pi / 2, 0.3, 1.234, -2.2]:
        R = P.perform_rotations([(Q, ang)])
        put(f"rot_single_{rot_id}", symp=P.symp_matrix, coeff=P.coeff_vec, q_symp=Q.symp_matrix,
            angle=np.array([np.nan if ang is None else ang]), out_symp=R.symp_matrix, out_coeff=R.coeff_vec)
        rot_id += 1
for k, (n, m, r) in enumerate([(4, 10, 6), (12, 30, 5), (1000, 25, 4)]):
    np.random.seed(600 + k)
    P = PauliwordOp.random(n, m).cleanup()
    rots = []
    for j in range(r):
        Q = PauliwordOp.random(n, 1)
        Q.coeff_vec[0] = 1
        rots.append((Q, [np.pi / 2, 0.7, None, -1.1, np.pi, 0.25][j % 6]))
    R = P.perform_rotations(rots)
    put(f"rot_seq_{k}", symp=P.symp_matrix, coeff=P.coeff_vec,
        q_symp=np.vstack([q.symp_matrix for q, _ in rots]),
        angle=np.array([np.nan if a is None else a for _, a in rots]),
        out_symp=R.symp_matrix, out_coeff=R.coeff_vec)

# sparse matrices
for k, (n, m) in enumerate([(1, 3), (2, 7), (3, 20), (5, 40), (8, 30), (10, 12)]):
    np.random.seed(700 + k)
    P = PauliwordOp.random(n, m)
    M = P.to_sparse_matrix
    psi = np.random.randn(1 << n) + 1j * np.random.randn(1 << n)
    psi /= np.linalg.norm(psi)
    put(f"matrix_rand_{k}", symp=P.symp_matrix, coeff=P.coeff_vec, dense=M.toarray() if n <= 8 else np.zeros(0),
        psi=psi, Hpsi=M @ psi, expval=np.array([np.vdot(psi, M @ psi)]))

# GF(2)
for k, (r, c, p) in enumerate([(5, 5, 0.5), (12, 30, 0.3), (40, 17, 0.4), (70, 130, 0.2), (130, 70, 0.5),
                               (28, 1114, 0.3), (64, 64, 0.5), (65, 129, 0.1), (10, 8, 0.0), (1, 1, 1.0)]):
    np.random.seed(800 + k)
    m = np.random.rand(r, c) < p
    put(f"gf2_rand_{k}", matrix=m, rref_norows=_rref_binary(m), rref=rref_binary(m) if m.any() else m,
        cref_norows=_cref_binary(m), cref=cref_binary(m) if m.any() else m)

# --- 3. molecular Hamiltonians (configs 2 and 4) -------------------------------------------------
ham_dir = os.path.join(HERE, "hamiltonians")
os.makedirs(ham_dir, exist_ok=True)
for fname in ["H2_STO-3G_SINGLET_JW.json", "H2O_STO-3G_SINGLET_JW.json", "HOOH_STO-3G_SINGLET_JW.json",
              "Be_STO-3G_SINGLET_JW.json", "NH3_STO-3G_SINGLET_JW.json"]:
    path = os.path.join(HAM_DIR, fname)
    if not os.path.exists(path):
        print("missing", fname)
        continue
    H, data = load_ham(fname)
    tag = fname.replace("_SINGLET_JW.json", "").replace("-", "")
    np.savez_compressed(os.path.join(ham_dir, tag + ".npz"),
                        symp=np.packbits(H.symp_matrix, axis=1), n_qubits=np.array([H.n_qubits]),
                        coeff=H.coeff_vec, hf_array=np.asarray(data["hf_array"]),
                        hf_energy=np.array([data["calculated_properties"]["HF"]["energy"]]))
    if H.n_qubits <= 14:
        S = IndependentOp.symmetry_generators(H)
        put(f"symgen_{tag}", symp=H.symp_matrix, gen_symp=S.symp_matrix, adj=H.adjacency_matrix)
        recon, mask = H.generator_reconstruction(H.generators)
        put(f"recon_{tag}", gen_symp=H.generators.symp_matrix, recon=recon, mask=mask)
        hf = QuantumState(np.asarray(data["hf_array"]).reshape(1, -1))
        put(f"hf_expval_{tag}", expval=np.array([H.expval(hf)]),
            psi_index=np.array([int("".join(str(int(b)) for b in data["hf_array"]), 2)]))
    if H.n_qubits <= 10:
        put(f"ham_matrix_{tag}", dense_diag=H.to_sparse_matrix.diagonal(),
            row0=H.to_sparse_matrix.getrow(0).toarray().ravel())

np.savez_compressed(os.path.join(HERE, "golden_vectors.npz"), **out)
print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "golden_vectors.npz")) / 1e6, "MB")
