"""Golden vectors for the callers either side of the hot path (SURVEY.md §8f and the remaining public methods of
PauliwordOp / QuantumState that sit directly on it), from the REAL reference. Test infrastructure only; run in
the build container:

    python tests/golden/make_golden_api.py

Reference code exercised (through oracle/shim for the uninstalled third-party packages):
  qubitwise_commutes_termwise / adjacency_matrix_qwc     operators/base.py:985-1009, 1064-1072
  reindex / tensor                                       base.py:493-521, 1188-1204
  get_graph / largest_clique / clique_cover              base.py:1206-1365
  jordan_generator_reconstruction                        base.py:562-602
  check_jordan_independent                               operators/utils.py:521-565
  QuantumState.random / zero / from_dictionary / from_array / sort / reindex / normalize_counts /
  partial_trace_over_qubits / get_rdm / sample_state / sectors_present /
  measure_state_in_computational_basis                   base.py:1630-2212
  get_PauliwordOp_projector / get_ij_operator / change_of_basis_XY_to_Z   base.py:2275-2537
  PauliwordOp.from_matrix (both strategies)              base.py:238-425
`np.product` (removed in NumPy 2) is aliased to `np.prod` for base.py:2038 — the only change to the
reference's behaviour, needed to run it on this image's NumPy.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shim"))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")
if not hasattr(np, "product"):
    np.product = np.prod

from symmer import PauliwordOp, QuantumState  # noqa: E402
from symmer import process  # noqa: E402
from symmer.operators import IndependentOp  # noqa: E402
from symmer.operators.base import (get_PauliwordOp_projector, get_ij_operator,  # noqa: E402
                                   change_of_basis_XY_to_Z)
from symmer.operators.utils import check_jordan_independent  # noqa: E402

process.method = 'single_thread'
out = {}


def put(name, **arrays):
    for k, v in arrays.items():
        out[f"{name}/{k}"] = np.asarray(v)


def put_cliques(name, cliques):
    keys = sorted(cliques.keys())
    put(name, keys=keys, sizes=[cliques[k].n_terms for k in keys],
        symp=np.vstack([cliques[k].symp_matrix for k in keys]),
        coeff=np.hstack([cliques[k].coeff_vec for k in keys]))


# --- 1. qubit-wise commutation ------------------------------------------------------------------
for i, (n, M, N, seed) in enumerate([(3, 7, 5, 0), (8, 20, 20, 1), (14, 40, 33, 2), (70, 30, 17, 3), (130, 12, 25, 4)]):
    np.random.seed(seed)
    A = PauliwordOp.random(n, M, density=0.25)
    B = PauliwordOp.random(n, N, density=0.25)
    put(f"qwc_{i}", a_symp=A.symp_matrix, b_symp=B.symp_matrix, out=A.qubitwise_commutes_termwise(B),
        adj=A.adjacency_matrix_qwc)
P = PauliwordOp.from_list(['XIZ', 'XZI', 'IYY', 'ZZZ', 'III', 'XYZ'])
put("qwc_doc", a_symp=P.symp_matrix, b_symp=P.symp_matrix, out=P.qubitwise_commutes_termwise(P),
    adj=P.adjacency_matrix_qwc)

# --- 2. reindex / tensor -------------------------------------------------------------------------
np.random.seed(5)
P = PauliwordOp.random(6, 9)
for i, qmap in enumerate([[2, 3, 0], {0: 2, 2: 3, 3: 0}, [5, 4, 3, 2, 1, 0], {1: 4, 4: 1}]):
    R = P.reindex(qmap)
    keys, vals = (list(qmap.keys()), list(qmap.values())) if isinstance(qmap, dict) else ([], qmap)
    put(f"reindex_{i}", symp=P.symp_matrix, coeff=P.coeff_vec, is_dict=[isinstance(qmap, dict)], keys=keys, vals=vals,
        out_symp=R.symp_matrix, out_coeff=R.coeff_vec)
np.random.seed(6)
P = PauliwordOp.random(70, 5)
perm = list(np.random.permutation(70))
R = P.reindex([int(v) for v in perm])
put("reindex_wide", symp=P.symp_matrix, coeff=P.coeff_vec, is_dict=[False], keys=[], vals=perm,
    out_symp=R.symp_matrix, out_coeff=R.coeff_vec)

for i, (nl, ml, nr, mr, seed) in enumerate([(2, 3, 3, 4, 7), (5, 6, 4, 9, 8), (60, 4, 10, 3, 9), (1, 1, 1, 1, 10)]):
    np.random.seed(seed)
    L = PauliwordOp.random(nl, ml)
    Rr = PauliwordOp.random(nr, mr)
    T = L.tensor(Rr)
    put(f"tensor_{i}", a_symp=L.symp_matrix, a_coeff=L.coeff_vec, b_symp=Rr.symp_matrix, b_coeff=Rr.coeff_vec,
        out_symp=T.symp_matrix, out_coeff=T.coeff_vec)

# --- 3. graphs and clique covers -----------------------------------------------------------------
import networkx as nx  # noqa: E402

np.random.seed(11)
H = PauliwordOp.random(5, 24, complex_coeffs=False)
put("graph_op", symp=H.symp_matrix, coeff=H.coeff_vec)
for rel in ['C', 'AC', 'QWC']:
    put(f"graph_{rel}", adj=nx.to_numpy_array(H.get_graph(edge_relation=rel), dtype=bool))
    big = H.largest_clique(edge_relation=rel)
    put(f"largest_clique_{rel}", symp=big.symp_matrix, coeff=big.coeff_vec)
    for strategy in ['largest_first', 'sorted_insertion', 'DSATUR']:
        put_cliques(f"clique_cover_{rel}_{strategy}", H.clique_cover(edge_relation=rel, strategy=strategy))

# --- 4. Jordan-product reconstruction ----------------------------------------------------------------
cases = {
    "jordan_small": (['ZZI', 'IIZ', 'XXI', 'YXI'],
                     ['ZZI', 'IIZ', 'ZZZ', 'XXI', 'YYI', 'YXI', 'YXZ', 'XYZ', 'IIX', 'ZIZ', 'III', 'XXZ']),
    "jordan_symmetric": (['ZIII', 'IZII', 'IIZZ'], ['ZZII', 'ZZZZ', 'XIII', 'IIIZ', 'IIZZ']),
    "jordan_ref_test": (['IIIZ', 'IIZI', 'ZIIZ', 'IXII', 'XIIX'],                      # utils.py:533-541 doc example
                        ['IIIZ', 'IIZI', 'ZIII', 'IXII', 'XIIX', 'ZXZZ', 'XXIX', 'YIIY', 'ZIZI']),
}
for name, (gens, terms) in cases.items():
    G = PauliwordOp.from_list(gens)
    Op = PauliwordOp.from_list(terms)
    recon, ok = Op.jordan_generator_reconstruction(G)
    put(name, gen_symp=G.symp_matrix, op_symp=Op.symp_matrix, recon=recon, ok=ok)
jord = {
    "not_indp": {'XXX': 2, 'XII': 2, 'IIX': 2, 'IXI': 2, 'ZZI': -2, 'IYY': -2},        # test_operator_utils.py:4-16
    "three_n": {'IX': 2, 'IY': 2, 'IZ': 2, 'ZI': 2, 'YI': -2, 'XI': -2},               # :19-31
    "larger_three_n": {'IX': 2, 'IY': 2, 'IZ': 2, 'ZI': 2, 'YI': -2, 'XI': -2, 'XX': -2},   # :34-47
    "xx_yy_zz": {'XX': 2, 'YY': 2, 'ZZ': 2},                                           # :50-59
    "indp": {'XZXIIIZI': 1, 'IZZIZZZX': 1, 'IXXXZXZI': 1, 'IIZIIXZX': 1, 'XIXIIIIZ': 1, 'ZYIIZZIY': 1,
             'IIIXIIII': 1, 'ZIZIIYZZ': 1, 'IIZIIIXY': 1},                             # :62-73
}
for name, d in jord.items():
    op = PauliwordOp.from_dictionary(d)
    put(f"jordan_indep_{name}", symp=op.symp_matrix, out=[bool(check_jordan_independent(op))])

# --- 5. QuantumState ---------------------------------------------------------------------------------
np.random.seed(21)
psi = QuantumState.random(6, 20)
put("qs_random", seed=[21], n_qubits=[6], n_terms=[20], state=psi.state_matrix, coeff=psi.state_op.coeff_vec)
z = QuantumState.zero(5)
put("qs_zero", state=z.state_matrix, coeff=z.state_op.coeff_vec)
d = {'1101': 0.3 + 0.1j, '0110': -0.5j, '1010': 0.7, '0000': 0.2}
s = QuantumState.from_dictionary(d)
put("qs_from_dictionary", keys=list(d.keys()), vals=list(d.values()), state=s.state_matrix, coeff=s.state_op.coeff_vec)
np.random.seed(22)
vec = np.random.randn(32) + 1j * np.random.randn(32)
vec[np.random.rand(32) < 0.4] = 0
vec /= np.linalg.norm(vec)
for kind, arr in [('ket', vec.reshape(-1, 1)), ('bra', vec.reshape(1, -1))]:
    s = QuantumState.from_array(arr)
    put(f"qs_from_array_{kind}", vec=arr, state=s.state_matrix, coeff=s.state_op.coeff_vec, vec_type=[s.vec_type])

np.random.seed(23)
mat = np.unique(np.random.randint(0, 2, (14, 5)), axis=0)
np.random.shuffle(mat)
cf = np.random.rand(mat.shape[0]) + 1j * np.random.rand(mat.shape[0])
cf /= np.linalg.norm(cf)
psi = QuantumState(mat, cf)
put("qs_base", state=mat, coeff=cf)
for key in ['magnitude', 'support']:
    for by in ['decreasing', 'increasing']:
        s = psi.sort(by=by, key=key)
        put(f"qs_sort_{key}_{by}", state=s.state_matrix, coeff=s.state_op.coeff_vec)
for i, qmap in enumerate([[2, 3, 0], {0: 4, 4: 0}]):
    s = psi.reindex(qmap)
    keys, vals = (list(qmap.keys()), list(qmap.values())) if isinstance(qmap, dict) else ([], qmap)
    put(f"qs_reindex_{i}", is_dict=[isinstance(qmap, dict)], keys=keys, vals=vals, state=s.state_matrix,
        coeff=s.state_op.coeff_vec)
counts = QuantumState(mat, np.arange(1, mat.shape[0] + 1).astype(float))
s = counts.normalize_counts
put("qs_normalize_counts", in_coeff=counts.state_op.coeff_vec, coeff=s.state_op.coeff_vec)
put("qs_dense", dense=psi.to_dense_matrix)
for i, qs in enumerate([[0], [1, 3], [0, 2, 4], []]):
    put(f"qs_ptrace_{i}", qubits=qs, rho=psi.partial_trace_over_qubits(qs))
    put(f"qs_rdm_{i}", qubits=qs, rho=psi.get_rdm(qs))
np.random.seed(24)
s = psi.sample_state(1000)
put("qs_sample", seed=[24], n_samples=[1000], state=s.state_matrix, coeff=s.state_op.coeff_vec)
np.random.seed(24)
s = psi.sample_state(1000, return_normalized=True)
put("qs_sample_norm", seed=[24], n_samples=[1000], state=s.state_matrix, coeff=s.state_op.coeff_vec)
S = IndependentOp.from_list(['ZIIII', 'IZZII', 'IIIZZ'])
put("qs_sectors", sym_symp=S.symp_matrix, out=psi.sectors_present(S))
for i, label in enumerate(['XYZIZ', 'ZZIII', 'YYXXI']):
    Pm = PauliwordOp.from_list([label])
    new_psi, Znew = psi.measure_state_in_computational_basis(Pm)
    put(f"qs_measure_{i}", p_symp=Pm.symp_matrix, state=new_psi.state_matrix, coeff=new_psi.state_op.coeff_vec,
        z_symp=Znew.symp_matrix, z_coeff=Znew.coeff_vec)
    U = change_of_basis_XY_to_Z(Pm)
    put(f"change_basis_{i}", p_symp=Pm.symp_matrix, symp=U.symp_matrix, coeff=U.coeff_vec)
put("qs_eq", same=[bool(psi == QuantumState(mat[::-1].copy(), cf[::-1].copy()))],
    different=[bool(psi == QuantumState(mat, cf[::-1].copy()))])

# --- 6. projector helpers ----------------------------------------------------------------------------
for i, (a, b, n) in enumerate([(0, 0, 2), (1, 2, 2), (5, 3, 3), (7, 7, 3), (9, 4, 4)]):
    op = get_ij_operator(a, b, n)
    put(f"ij_{i}", ijn=[a, b, n], symp=op.symp_matrix, coeff=op.coeff_vec)
for i, label in enumerate(['I+0*1II', '01', 'II', '%-*', '1I0']):
    op = get_PauliwordOp_projector(label)
    put(f"projector_{i}", label=[label], symp=op.symp_matrix, coeff=op.coeff_vec)

# --- 7. from_matrix ----------------------------------------------------------------------------------
np.random.seed(31)
for i, (n, kind) in enumerate([(1, 'dense'), (2, 'dense'), (3, 'dense'), (4, 'sparse'), (3, 'ragged')]):
    side = 2 ** n
    if kind == 'dense':
        m = np.random.randn(side, side) + 1j * np.random.randn(side, side)
    elif kind == 'sparse':
        m = np.zeros((side, side), dtype=complex)
        for _ in range(9):
            m[np.random.randint(side), np.random.randint(side)] = np.random.randn() + 1j * np.random.randn()
    else:
        m = np.random.randn(side - 2, side - 3)          # padded with zeros by from_matrix (real matrix)
    for strategy in ['projector', 'full_basis']:
        op = PauliwordOp.from_matrix(m, strategy=strategy, disable_loading_bar=True)
        put(f"from_matrix_{i}_{strategy}", matrix=m, symp=op.symp_matrix, coeff=op.coeff_vec)
np.random.seed(32)
H = PauliwordOp.random(3, 10)
basis = PauliwordOp.from_list(['XXI', 'ZZZ', 'IYI', 'III', 'ZIX'])
op = PauliwordOp.from_matrix(H.to_sparse_matrix.toarray(), operator_basis=basis, disable_loading_bar=True)
put("from_matrix_basis", matrix=H.to_sparse_matrix.toarray(), basis_symp=basis.symp_matrix, symp=op.symp_matrix,
    coeff=op.coeff_vec)

# --- 8. gate library, exponentials, state projection --------------------------------------------------
import json  # noqa: E402
from symmer import QubitTapering  # noqa: E402
from symmer.evolution import trotter  # noqa: E402
from symmer.evolution.exponentiation import exponentiate_single_Pop  # noqa: E402
from symmer.evolution import gate_library as gl  # noqa: E402

gates = {"I": gl.I(3), "X": gl.X(3, 1), "Y": gl.Y(3, 2), "Z": gl.Z(3, 0), "Had": gl.Had(3, 1), "CZ": gl.CZ(3, 0, 2),
         "CX": gl.CX(3, 2, 0), "CY": gl.CY(3, 1, 2), "RX": gl.RX(3, 0, 0.37), "RY": gl.RY(3, 1, -1.2),
         "RZ": gl.RZ(3, 2, 2.5), "U1": gl.U1(3, 1, 0.81), "S": gl.S(3, 2)}
for name, op in gates.items():
    op = op.cleanup()
    put(f"gate_{name}", symp=op.symp_matrix, coeff=op.coeff_vec)
P1 = PauliwordOp.from_list(['XYZI'], [0.3 - 0.8j])
E = exponentiate_single_Pop(P1)
put("exp_single", symp=P1.symp_matrix, coeff=P1.coeff_vec, out_symp=E.symp_matrix, out_coeff=E.coeff_vec)
np.random.seed(41)
T = PauliwordOp.random(3, 4)
for trotnum in [1, 3]:
    E = trotter(T.multiply_by_constant(0.2j), trotnum=trotnum)
    put(f"trotter_{trotnum}", symp=T.symp_matrix, coeff=T.coeff_vec, out_symp=E.symp_matrix, out_coeff=E.coeff_vec)

for tag, fname in [("H3+", "H3+_STO-3G_SINGLET_JW.json"), ("Be", "Be_STO-3G_SINGLET_JW.json")]:
    with open(os.path.join("/root/reference/tests/hamiltonian_data", fname)) as f:
        dd = json.load(f)
    H = PauliwordOp.from_dictionary({k: complex(v[0], v[1]) for k, v in dd["hamiltonian"].items()})
    hf = np.asarray(dd["data"]["hf_array"], dtype=int)
    QT = QubitTapering(H)
    Ht = QT.taper_it(ref_state=hf)
    np.random.seed(42)
    psi = QuantumState(hf) if tag == "Be" else QuantumState.random(H.n_qubits, 6)
    proj = QT.project_state(psi)
    put(f"project_state_{tag}", h_symp=H.symp_matrix, h_coeff=H.coeff_vec, hf=hf, psi_state=psi.state_matrix,
        psi_coeff=psi.state_op.coeff_vec, tapered_symp=Ht.symp_matrix, tapered_coeff=Ht.coeff_vec,
        out_state=proj.state_matrix, out_coeff=proj.state_op.coeff_vec)

# --- 9. symmer/utils.py helpers -----------------------------------------------------------------------
from symmer.utils import exact_gs_energy, get_entanglement_entropy, tensor_list, product_list  # noqa: E402

for tag, fname in [("H3+", "H3+_STO-3G_SINGLET_JW.json"), ("Be", "Be_STO-3G_SINGLET_JW.json")]:
    with open(os.path.join("/root/reference/tests/hamiltonian_data", fname)) as f:
        dd = json.load(f)
    H = PauliwordOp.from_dictionary({k: complex(v[0], v[1]) for k, v in dd["hamiltonian"].items()})
    N = PauliwordOp.from_dictionary({k: complex(v[0], v[1]) for k, v in dd["data"]["auxiliary_operators"]["number_operator"].items()})
    e0, psi0 = exact_gs_energy(H.to_sparse_matrix)
    n_part = int(dd["data"]["n_particles"])
    e_n, psi_n = exact_gs_energy(H.to_sparse_matrix, n_particles=n_part, number_operator=N, n_eigs=12)
    put(f"gs_{tag}", h_symp=H.symp_matrix, h_coeff=H.coeff_vec, n_symp=N.symp_matrix, n_coeff=N.coeff_vec,
        e0=[e0], n_particles=[n_part], e_n=[e_n], fci=[dd["data"]["calculated_properties"]["FCI"]["energy"]])
np.random.seed(51)
psi = QuantumState.random(5, 12)
put("entropy", state=psi.state_matrix, coeff=psi.state_op.coeff_vec,
    out=[get_entanglement_entropy(psi, [0, 1]), get_entanglement_entropy(psi, [2]), get_entanglement_entropy(psi, [0, 2, 4])])
np.random.seed(52)
ops3 = [PauliwordOp.random(2, 3), PauliwordOp.random(1, 2), PauliwordOp.random(2, 2)]
T3 = tensor_list(ops3)
put("util_tensor_list", **{f"symp_{i}": o.symp_matrix for i, o in enumerate(ops3)}, **{f"coeff_{i}": o.coeff_vec for i, o in enumerate(ops3)},
    out_symp=T3.symp_matrix, out_coeff=T3.coeff_vec)
sq = [PauliwordOp.random(3, 4) for _ in range(3)]
P3 = product_list(sq)
put("util_product_list", **{f"symp_{i}": o.symp_matrix for i, o in enumerate(sq)}, **{f"coeff_{i}": o.coeff_vec for i, o in enumerate(sq)},
    out_symp=P3.symp_matrix, out_coeff=P3.coeff_vec)

from symmer.utils import random_anitcomm_2n_1_PauliwordOp, gram_schmidt_from_quantum_state  # noqa: E402

for i, (n, cplx, cliff, seed) in enumerate([(4, False, False, 61), (5, True, True, 62), (3, False, True, 63)]):
    np.random.seed(seed)
    AC = random_anitcomm_2n_1_PauliwordOp(n, complex_coeff=cplx, apply_clifford=cliff)
    put(f"anticomm_{i}", args=[n, int(cplx), int(cliff), seed], symp=AC.symp_matrix, coeff=AC.coeff_vec)
np.random.seed(64)
psi = QuantumState.random(3, 5)
put("gram_schmidt", state=psi.state_matrix, coeff=psi.state_op.coeff_vec, out=gram_schmidt_from_quantum_state(psi))

from symmer.evolution.circuit_symmerlator import CircuitSymmerlator  # noqa: E402

QASM = ("OPENQASM 3.0;\ninclude \"stdgates.inc\";\nqubit[4] q;\nh q[0];\ncx q[0], q[1];\nrz(pi/3) q[1];\n"
        "sdg q[2];\ncz q[1], q[3];\nry(-0.4) q[3];\nswap q[2], q[3];\nsx q[0];\ny q[2];\nrx(3*pi/2) q[0];\n")
np.random.seed(71)
O = PauliwordOp.random(4, 30, complex_coeffs=False)
CS = CircuitSymmerlator.from_qasm(QASM)
rot = CS.apply_sequence(O)
put("circuit_qasm", qasm=[QASM], o_symp=O.symp_matrix, o_coeff=O.coeff_vec, n_steps=[len(CS.sequence)],
    seq_symp=np.vstack([p.symp_matrix for p, _ in CS.sequence]), seq_angle=[a for _, a in CS.sequence],
    rot_symp=rot.symp_matrix, rot_coeff=rot.coeff_vec, expval=[CS.evaluate(O)])

path = os.path.join(HERE, "api_vectors.npz")
np.savez_compressed(path, **out)
print(f"wrote {len(out)} arrays to {path} ({os.path.getsize(path) / 1024:.0f} KB)")
