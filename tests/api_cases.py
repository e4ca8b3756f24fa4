"""Assertions on the public API either side of the hot path, against vectors produced by the REAL reference
(tests/golden/make_golden_api.py -> tests/golden/api_vectors.npz). Each case takes the golden dictionary and runs
through `symmer_b200`'s reference-facing classes; `tests/test_gpu_api_ext.py` runs them on the CUDA kernels
(`-m gpu`), `tests/test_host_logic.py` runs the same cases on the CPU box with the kernels swapped for the
NumPy test double (host logic only)."""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pauli_oracle as po  # noqa: E402


def load_api_golden():
    data = np.load(os.path.join(ROOT, "tests", "golden", "api_vectors.npz"))
    cases = {}
    for key in data.files:
        name, field = key.split("/", 1)
        cases.setdefault(name, {})[field] = data[key]
    return cases


def _same_terms(symp_a, coeff_a, symp_b, coeff_b, scale=1.0):
    ok, why = po.compare_term_sets(symp_a, coeff_a, symp_b, coeff_b, scale=scale)
    assert ok, why


def _same_state(state, ref_matrix, ref_coeff):
    """Equal bit strings with equal amplitudes, order ignored."""
    got = {tuple(r): c for r, c in zip(np.asarray(state.state_matrix).tolist(), state.state_op.coeff_vec)}
    want = {tuple(r): c for r, c in zip(np.asarray(ref_matrix).tolist(), ref_coeff)}
    assert set(got) == set(want)
    for k in want:
        assert abs(got[k] - want[k]) <= 1e-12 * max(1.0, abs(want[k])), (k, got[k], want[k])


def _qmap(g):
    vals = [int(v) for v in g["vals"]]
    if bool(g["is_dict"][0]):
        return dict(zip([int(k) for k in g["keys"]], vals))
    return vals


# ------------------------------------------------------------------------------------------------ cases
def case_qwc(api, G):
    from symmer_b200 import PauliwordOp
    names = [n for n in G if n.startswith("qwc_")]
    assert len(names) >= 6
    for name in names:
        g = G[name]
        A = PauliwordOp(g["a_symp"], np.ones(g["a_symp"].shape[0]))
        B = PauliwordOp(g["b_symp"], np.ones(g["b_symp"].shape[0]))
        got = A.qubitwise_commutes_termwise(B)
        assert got.dtype == bool and np.array_equal(got, g["out"]), name
        assert np.array_equal(A.adjacency_matrix_qwc, g["adj"]), name
        assert np.array_equal(po.qubitwise_commutes_termwise(g["a_symp"], g["b_symp"]), g["out"]), name


def case_reindex(api, G):
    from symmer_b200 import PauliwordOp
    names = [n for n in G if n.startswith("reindex_")]
    assert len(names) >= 5
    for name in names:
        g = G[name]
        P = PauliwordOp(g["symp"], g["coeff"])
        R = P.reindex(_qmap(g))
        assert np.array_equal(R.symp_matrix, g["out_symp"]), name           # no dedup, order kept
        assert np.array_equal(R.coeff_vec, g["out_coeff"]), name
        assert np.array_equal(po.reindex(g["symp"], _qmap(g)), g["out_symp"]), name
    P = PauliwordOp.from_list(['XYZ'])
    for bad in ([0, 0, 1], {0: 1}):
        try:
            P.reindex(bad)
        except AssertionError:
            continue
        raise AssertionError(f"reindex({bad}) must be rejected like the reference (base.py:511-512)")


def case_tensor(api, G):
    from symmer_b200 import PauliwordOp
    names = [n for n in G if n.startswith("tensor_")]
    assert len(names) >= 4
    for name in names:
        g = G[name]
        L = PauliwordOp(g["a_symp"], g["a_coeff"])
        R = PauliwordOp(g["b_symp"], g["b_coeff"])
        T = L.tensor(R)
        assert T.n_qubits == L.n_qubits + R.n_qubits
        _same_terms(T.symp_matrix, T.coeff_vec, g["out_symp"], g["out_coeff"])
        s, c = po.tensor(g["a_symp"], g["a_coeff"], g["b_symp"], g["b_coeff"])
        _same_terms(s, c, g["out_symp"], g["out_coeff"])


def case_graphs(api, G):
    import networkx as nx
    from symmer_b200 import PauliwordOp
    H = PauliwordOp(G["graph_op"]["symp"], G["graph_op"]["coeff"])
    for rel in ['C', 'AC', 'QWC']:
        adj = nx.to_numpy_array(H.get_graph(edge_relation=rel), dtype=bool)
        assert np.array_equal(adj, G[f"graph_{rel}"]["adj"]), rel
        big = H.largest_clique(edge_relation=rel)
        g = G[f"largest_clique_{rel}"]
        _same_terms(big.symp_matrix, big.coeff_vec, g["symp"], g["coeff"])
        for strategy in ['largest_first', 'sorted_insertion', 'DSATUR']:
            g = G[f"clique_cover_{rel}_{strategy}"]
            cover = H.clique_cover(edge_relation=rel, strategy=strategy)
            assert sorted(cover.keys()) == [int(k) for k in g["keys"]], (rel, strategy)
            lo = 0
            for key, size in zip(g["keys"], g["sizes"]):
                clq = cover[int(key)]
                _same_terms(clq.symp_matrix, clq.coeff_vec, g["symp"][lo:lo + size], g["coeff"][lo:lo + size])
                lo += int(size)
    labelled = H.get_graph(edge_relation='C', label_nodes=True)
    assert set(labelled.nodes) == set(po.to_strings(H.symp_matrix))
    try:
        H.get_graph(edge_relation='nope')
    except TypeError:
        pass
    else:
        raise AssertionError("unknown edge relation must raise TypeError (base.py:1241)")
    # duplicate terms: the sorted-insertion cover follows the reference's running sums literally
    D = PauliwordOp.from_list(['XX', 'ZZ', 'XX', 'YI'], [1, 2, 3, 0.5])
    cover = D.clique_cover(strategy='sorted_insertion')
    total = sum(cover.values())
    assert total == D.cleanup()


def case_jordan(api, G):
    from symmer_b200 import PauliwordOp
    from symmer_b200.utils import check_jordan_independent
    for name in ["jordan_small", "jordan_symmetric", "jordan_ref_test"]:
        g = G[name]
        gens = PauliwordOp(g["gen_symp"], np.ones(g["gen_symp"].shape[0]))
        op = PauliwordOp(g["op_symp"], np.ones(g["op_symp"].shape[0]))
        recon, ok = op.jordan_generator_reconstruction(gens)
        assert np.array_equal(ok, g["ok"]), name
        assert np.array_equal(recon[ok], g["recon"][g["ok"]]), name
        assert recon.dtype.kind == 'i'
    names = [n for n in G if n.startswith("jordan_indep_")]
    assert len(names) == 5
    for name in names:
        g = G[name]
        op = PauliwordOp(g["symp"], np.ones(g["symp"].shape[0]))
        assert bool(check_jordan_independent(op)) == bool(g["out"][0]), name


def case_quantum_state_constructors(api, G):
    from symmer_b200 import QuantumState
    g = G["qs_random"]
    np.random.seed(int(g["seed"][0]))
    psi = QuantumState.random(int(g["n_qubits"][0]), int(g["n_terms"][0]))
    _same_state(psi, g["state"], g["coeff"])
    assert psi.vec_type == 'ket' and psi._is_normalized()
    z = QuantumState.zero(5)
    _same_state(z, G["qs_zero"]["state"], G["qs_zero"]["coeff"])
    assert QuantumState.zero(3, vec_type='bra').vec_type == 'bra'
    g = G["qs_from_dictionary"]
    s = QuantumState.from_dictionary(dict(zip([str(k) for k in g["keys"]], g["vals"])))
    assert np.array_equal(s.state_matrix, g["state"]) and np.allclose(s.state_op.coeff_vec, g["coeff"], rtol=1e-15, atol=0)
    s = QuantumState.from_dictionary({'10': (0.6, 0.0), '01': (0.0, 0.8)})
    assert np.allclose(s.state_op.coeff_vec, [0.6, 0.8j])
    for kind in ['ket', 'bra']:
        g = G[f"qs_from_array_{kind}"]
        s = QuantumState.from_array(g["vec"])
        assert s.vec_type == kind == str(g["vec_type"][0])
        assert np.array_equal(s.state_matrix, g["state"]) and np.array_equal(s.state_op.coeff_vec, g["coeff"])
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        QuantumState.from_array(np.array([[1.0], [1.0]]))
        assert any('not normalized' in str(w.message) for w in caught)
    for bad in (np.ones(4), np.ones((3, 1))):
        try:
            QuantumState.from_array(bad)
        except AssertionError:
            continue
        raise AssertionError("from_array must reject non-vectors and non power-of-two sizes")
    np.random.seed(3)
    h = QuantumState.haar_random(3)
    assert h.n_qubits == 3 and h._is_normalized() and h.vec_type == 'ket'
    assert QuantumState.haar_random(2, vec_type='bra').vec_type == 'bra'


def case_quantum_state_methods(api, G):
    from symmer_b200 import IndependentOp, PauliwordOp, QuantumState
    base = G["qs_base"]
    psi = QuantumState(base["state"], base["coeff"])
    for key in ['magnitude', 'support']:
        for by in ['decreasing', 'increasing']:
            g = G[f"qs_sort_{key}_{by}"]
            s = psi.sort(by=by, key=key)
            assert np.array_equal(s.state_matrix, g["state"]), (key, by)
            assert np.array_equal(s.state_op.coeff_vec, g["coeff"]), (key, by)
    for bad in (dict(key='nope'), dict(by='sideways')):
        try:
            psi.sort(**bad)
        except ValueError:
            continue
        raise AssertionError("sort must reject unknown keys / orders (base.py:1902-1907)")
    for i in range(2):
        g = G[f"qs_reindex_{i}"]
        s = psi.reindex(_qmap(g))
        assert np.array_equal(s.state_matrix, g["state"]) and np.array_equal(s.state_op.coeff_vec, g["coeff"])
    g = G["qs_normalize_counts"]
    s = QuantumState(base["state"], g["in_coeff"]).normalize_counts
    assert np.allclose(s.state_op.coeff_vec, g["coeff"], rtol=1e-14, atol=0)
    assert np.allclose(psi.to_dense_matrix, G["qs_dense"]["dense"], rtol=1e-15, atol=0)
    assert psi.to_dense_matrix.shape == (32, 1)
    for i in range(4):
        g = G[f"qs_ptrace_{i}"]
        rho = psi.partial_trace_over_qubits([int(q) for q in g["qubits"]])
        assert rho.shape == g["rho"].shape and np.allclose(rho, g["rho"], rtol=1e-12, atol=1e-15), i
        g = G[f"qs_rdm_{i}"]
        rho = psi.get_rdm([int(q) for q in g["qubits"]])
        assert rho.shape == g["rho"].shape and np.allclose(rho, g["rho"], rtol=1e-12, atol=1e-15), i
    for name, norm in [("qs_sample", False), ("qs_sample_norm", True)]:
        g = G[name]
        np.random.seed(int(g["seed"][0]))
        s = psi.sample_state(int(g["n_samples"][0]), return_normalized=norm)
        assert np.array_equal(s.state_matrix, g["state"]) and np.allclose(s.state_op.coeff_vec, g["coeff"], rtol=1e-15)
    try:
        QuantumState(base["state"], 2 * base["coeff"]).sample_state(10)
    except ValueError:
        pass
    else:
        raise AssertionError("sampling an unnormalised state must raise (base.py:2082-2083)")
    g = G["qs_sectors"]
    S = IndependentOp(g["sym_symp"], np.ones(g["sym_symp"].shape[0]))
    assert np.allclose(psi.sectors_present(S), g["out"], rtol=1e-12, atol=1e-14)
    for i in range(3):
        g = G[f"qs_measure_{i}"]
        Pm = PauliwordOp(g["p_symp"], [1])
        new_psi, Znew = psi.measure_state_in_computational_basis(Pm)
        _same_state(new_psi, g["state"], g["coeff"])
        _same_terms(Znew.symp_matrix, Znew.coeff_vec, g["z_symp"], g["z_coeff"])
        assert not np.any(Znew.X_block)
    assert (psi == QuantumState(base["state"][::-1].copy(), base["coeff"][::-1].copy())) is bool(G["qs_eq"]["same"][0])
    assert (psi == QuantumState(base["state"], base["coeff"][::-1].copy())) is bool(G["qs_eq"]["different"][0])
    assert sum([psi, psi]) == QuantumState(base["state"], 2 * base["coeff"])


def case_projector_helpers(api, G):
    from symmer_b200 import change_of_basis_XY_to_Z, get_ij_operator, get_PauliwordOp_projector, PauliwordOp
    for i in range(5):
        g = G[f"ij_{i}"]
        a, b, n = [int(v) for v in g["ijn"]]
        op = get_ij_operator(a, b, n)
        _same_terms(op.symp_matrix, op.coeff_vec, g["symp"], g["coeff"])
        symp, coeff = get_ij_operator(a, b, n, return_operator=False)
        assert np.array_equal(symp, g["symp"]) and np.allclose(coeff, g["coeff"], rtol=1e-15, atol=0)
        dense = np.zeros((1 << n, 1 << n), dtype=complex)
        dense[a, b] = 1
        assert np.allclose(op.to_sparse_matrix.toarray(), dense, atol=1e-15)
    for i in range(5):
        g = G[f"projector_{i}"]
        op = get_PauliwordOp_projector(str(g["label"][0]))
        _same_terms(op.symp_matrix, op.coeff_vec, g["symp"], g["coeff"])
    assert get_PauliwordOp_projector(list('0+')) == get_PauliwordOp_projector('0+')
    for i in range(3):
        g = G[f"change_basis_{i}"]
        Pm = PauliwordOp(g["p_symp"], [1])
        U = change_of_basis_XY_to_Z(Pm)
        _same_terms(U.symp_matrix, U.coeff_vec, g["symp"], g["coeff"])


def case_from_matrix(api, G):
    import scipy.sparse as sps
    from symmer_b200 import PauliwordOp
    names = [n for n in G if n.startswith("from_matrix_") and n != "from_matrix_basis"]
    assert len(names) == 10
    for name in names:
        g = G[name]
        strategy = 'projector' if name.endswith('projector') else 'full_basis'
        scale = float(np.abs(g["matrix"]).max())
        for mat in (g["matrix"], sps.csr_matrix(g["matrix"])):
            if sps.issparse(mat) and mat.shape[0] != mat.shape[1]:
                continue                                   # the reference pads dense matrices only
            op = PauliwordOp.from_matrix(mat, strategy=strategy, disable_loading_bar=True)
            _same_terms(op.symp_matrix, op.coeff_vec, g["symp"], g["coeff"], scale=scale)
            # ordered by the [X|Z] bit string like the reference's output (base.py:353-356)
            keys = [int(''.join('1' if b else '0' for b in row), 2) for row in op.symp_matrix]
            assert keys == sorted(keys), name
            n = op.n_qubits
            side = 1 << n
            dense = np.zeros((side, side), dtype=complex)
            m = np.asarray(g["matrix"])
            dense[:m.shape[0], :m.shape[1]] = m
            assert np.allclose(op.to_sparse_matrix.toarray(), dense, atol=1e-13 * max(1.0, scale))
    g = G["from_matrix_basis"]
    basis = PauliwordOp(g["basis_symp"], np.ones(g["basis_symp"].shape[0]))
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        op = PauliwordOp.from_matrix(g["matrix"], operator_basis=basis)
        assert any('sufficiently expressive' in str(w.message) for w in caught)
    _same_terms(op.symp_matrix, op.coeff_vec, g["symp"], g["coeff"])
    np.random.seed(12)
    H = PauliwordOp.random(4, 30)
    assert PauliwordOp.from_matrix(H.to_sparse_matrix) == H
    assert PauliwordOp.from_matrix(H.to_sparse_matrix.toarray(), strategy='full_basis') == H
    U = PauliwordOp.haar_random(2)
    assert (U * U.dagger).cleanup(zero_threshold=1e-12) == PauliwordOp.from_list(['II'])
    try:
        PauliwordOp.from_matrix(np.eye(2), strategy='nope')
    except ValueError:
        pass
    else:
        raise AssertionError("unknown strategy must raise ValueError (base.py:423)")
    assert PauliwordOp.from_matrix(np.zeros((4, 4))).n_terms == 0


def case_evolution_and_state_projection(api, G):
    from symmer_b200 import PauliwordOp, QuantumState, QubitTapering
    from symmer_b200 import evolution as ev
    gates = {"I": ev.I(3), "X": ev.X(3, 1), "Y": ev.Y(3, 2), "Z": ev.Z(3, 0), "Had": ev.Had(3, 1), "CZ": ev.CZ(3, 0, 2),
             "CX": ev.CX(3, 2, 0), "CY": ev.CY(3, 1, 2), "RX": ev.RX(3, 0, 0.37), "RY": ev.RY(3, 1, -1.2),
             "RZ": ev.RZ(3, 2, 2.5), "U1": ev.U1(3, 1, 0.81), "S": ev.S(3, 2)}
    for name, op in gates.items():
        g = G[f"gate_{name}"]
        op = op.cleanup()
        _same_terms(op.symp_matrix, op.coeff_vec, g["symp"], g["coeff"])
    cx = ev.CX(2, 0, 1).to_sparse_matrix.toarray()
    assert np.allclose(cx, [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], atol=1e-14)
    g = G["exp_single"]
    E = ev.exponentiate_single_Pop(PauliwordOp(g["symp"], g["coeff"]))
    _same_terms(E.symp_matrix, E.coeff_vec, g["out_symp"], g["out_coeff"])
    try:
        ev.exponentiate_single_Pop(PauliwordOp.from_list(['X', 'Z']))
    except AssertionError:
        pass
    else:
        raise AssertionError("only single Pauli terms can be exponentiated (exponentiation.py:17)")
    for trotnum in [1, 3]:
        g = G[f"trotter_{trotnum}"]
        E = ev.trotter(PauliwordOp(g["symp"], g["coeff"]).multiply_by_constant(0.2j), trotnum=trotnum)
        _same_terms(E.symp_matrix, E.coeff_vec, g["out_symp"], g["out_coeff"])
    for tag in ["H3+", "Be"]:
        g = G[f"project_state_{tag}"]
        H = PauliwordOp(g["h_symp"], g["h_coeff"])
        QT = QubitTapering(H)
        Ht = QT.taper_it(ref_state=g["hf"])
        _same_terms(Ht.symp_matrix, Ht.coeff_vec, g["tapered_symp"], g["tapered_coeff"],
                    scale=float(np.abs(g["h_coeff"]).max()))
        psi = QuantumState(g["psi_state"], g["psi_coeff"])
        proj = QT.project_state(psi)
        assert proj.n_qubits == Ht.n_qubits
        _same_state(proj, g["out_state"], g["out_coeff"])


def case_symmetry_generators_device_path(api, G):
    """The device-resident search (bit transpose + packed identity block + one reduction) against the array-seam
    form of the same search and the reference's generators (tests/golden/golden_vectors.npz, taper_vectors.npz)."""
    import warnings as _w
    from conftest import load_hamiltonian
    from symmer_b200 import IndependentOp, PauliwordOp
    gold = np.load(os.path.join(ROOT, "tests", "golden", "golden_vectors.npz"))
    for tag in ["H2O_STO3G", "Be_STO3G"]:
        symp, coeff, _ = load_hamiltonian(tag)
        H = PauliwordOp(symp, coeff)
        S = IndependentOp.symmetry_generators(H)
        ref = gold[f"symgen_{tag}/gen_symp"]
        assert np.array_equal(S.symp_matrix, ref), tag                          # bit-exact, reference order
        S_host = IndependentOp._symmetry_generators_host(H)
        assert np.array_equal(S_host.symp_matrix, ref), tag
        assert isinstance(S, IndependentOp) and S.target_sqp == 'Z' and np.all(S.coeff_vec == 1)
        assert np.array_equal(S.adjacency_matrix, S_host.adjacency_matrix)
        assert np.all(S.commutes_termwise(H))
        gens = H.generators                                                     # device-resident reduction of the packed rows
        assert np.array_equal(gens.symp_matrix, gold[f"recon_{tag}/gen_symp"]), tag
        recon, mask = H.generator_reconstruction(gens)
        assert np.array_equal(recon, gold[f"recon_{tag}/recon"]) and np.array_equal(mask, gold[f"recon_{tag}/mask"])
    from symmer_b200.utils import check_independent
    assert check_independent(PauliwordOp.from_list(['XX', 'ZZ', 'XI']))
    assert not check_independent(PauliwordOp.from_list(['XX', 'ZZ', 'YY']))
    assert not check_independent(PauliwordOp.from_list(['X', 'Z', 'Y']))
    # wide rows (two words per block) and a term count that is not a multiple of 64
    np.random.seed(77)
    gens = PauliwordOp.random(70, 5, diagonal=True)
    pool = [gens[i] for i in range(5)]
    terms = PauliwordOp.from_list(['I' * 70], [1.0])
    for k in range(1, 32):
        picks = [pool[j] for j in range(5) if (k >> j) & 1]
        prod = picks[0]
        for extra in picks[1:]:
            prod = prod * extra
        terms = terms.append(prod)
    S = IndependentOp.symmetry_generators(terms)
    S_host = IndependentOp._symmetry_generators_host(terms)
    assert np.array_equal(S.symp_matrix, S_host.symp_matrix) and S.n_terms >= 65
    # generators that do not mutually commute: largest commuting subset (independent_op.py:132-144)
    lone = PauliwordOp.from_list(['ZI'])
    S = IndependentOp.symmetry_generators(lone)
    S_host = IndependentOp._symmetry_generators_host(lone)
    assert np.array_equal(S.symp_matrix, S_host.symp_matrix) and np.all(S.adjacency_matrix)
    S_all = IndependentOp.symmetry_generators(lone, commuting_override=True)
    assert S_all.n_terms == 3 and not np.all(S_all.adjacency_matrix)
    with _w.catch_warnings(record=True) as caught:
        _w.simplefilter("always")
        none = IndependentOp.symmetry_generators(PauliwordOp.from_list(['X', 'Z']))
        assert none.n_terms == 0 and any('no Z2 symmetries' in str(w.message) for w in caught)


def case_tapering_intermediates(api, G):
    """QubitTapering(H).taper_it against the reference's vectors (tests/golden/make_golden_taper.py): generators,
    sector, the rotation list in order, rotated stabilizers and free qubits bit-exact, the tapered operator as a
    term set — the regression net of the host-side stabilizer bookkeeping (independent_op.py:146-383)."""
    from conftest import load_hamiltonian
    from symmer_b200 import PauliwordOp, QubitTapering
    data = np.load(os.path.join(ROOT, "tests", "golden", "taper_vectors.npz"))
    for tag in ["H2O_STO3G", "Be_STO3G"]:
        symp, coeff, _ = load_hamiltonian(tag)
        for sqp in ["Z", "X"]:
            g = {k.split("/", 1)[1]: data[k] for k in data.files if k.startswith(f"taper_{tag}_{sqp}/")}
            nq = int(g["n_out_qubits"][0])
            out_symp = np.unpackbits(g["out_symp"], axis=1)[:, :2 * nq].astype(bool)
            H = PauliwordOp(symp, coeff)
            qt = QubitTapering(H, target_sqp=sqp)
            assert qt.n_taper == g["gen_symp"].shape[0]
            assert np.array_equal(qt.symmetry_generators.symp_matrix, g["gen_symp"])
            out = qt.taper_it(ref_state=g["hf"])
            assert np.array_equal(qt.stabilizers.coeff_vec.real, g["sector"]), (tag, sqp)
            rot = np.array([r.symp_matrix[0] for r, _ in qt.stabilizers.stabilizer_rotations]).reshape(-1, symp.shape[1])
            assert np.array_equal(rot, g["rotations"]), (tag, sqp)
            assert np.array_equal(qt.rotated_stabilizers.symp_matrix, g["rotated_symp"]), (tag, sqp)
            assert np.array_equal(qt.rotated_stabilizers.coeff_vec.real, g["rotated_coeff"]), (tag, sqp)
            assert np.array_equal(qt.free_qubit_indices, g["free"])
            assert out.n_qubits == nq and out.n_terms == out_symp.shape[0]
            _same_terms(out.symp_matrix, out.coeff_vec, out_symp, g["out_coeff"], scale=float(np.abs(coeff).max()))


def case_symmer_utils(api, G):
    from symmer_b200 import PauliwordOp, QuantumState
    from symmer_b200.symmer_utils import (exact_gs_energy, get_entanglement_entropy, matrix_allclose, product_list,
                                          tensor_list)
    for tag in ["H3+", "Be"]:
        g = G[f"gs_{tag}"]
        H = PauliwordOp(g["h_symp"], g["h_coeff"])
        N = PauliwordOp(g["n_symp"], g["n_coeff"])
        e_ref = float(g["e0"][0])
        # matrix-free: every H|v> of the Lanczos iteration through the device kernel
        e_dev, psi_dev = exact_gs_energy(H)
        assert abs(e_dev - e_ref) < 1e-8, (tag, e_dev, e_ref)
        assert abs(H.expval(psi_dev) - e_ref) < 1e-8
        # the reference's own call form (CSR matrix in)
        e_csr, _ = exact_gs_energy(H.to_sparse_matrix)
        assert abs(e_csr - e_ref) < 1e-8
        e_n, psi_n = exact_gs_energy(H, n_particles=int(g["n_particles"][0]), number_operator=N, n_eigs=12)
        assert abs(e_n - float(g["e_n"][0])) < 1e-8 and abs(e_n - float(g["fci"][0])) < 1e-6, (tag, e_n)
        assert abs(N.expval(psi_n) - int(g["n_particles"][0])) < 1e-6
    g = G["entropy"]
    psi = QuantumState(g["state"], g["coeff"])
    got = [get_entanglement_entropy(psi, [0, 1]), get_entanglement_entropy(psi, [2]), get_entanglement_entropy(psi, [0, 2, 4])]
    assert np.allclose(got, g["out"], rtol=1e-10, atol=1e-12)
    for name, fn in [("util_tensor_list", tensor_list), ("util_product_list", product_list)]:
        g = G[name]
        factors = [PauliwordOp(g[f"symp_{i}"], g[f"coeff_{i}"]) for i in range(3)]
        out = fn(factors)
        _same_terms(out.symp_matrix, out.coeff_vec, g["out_symp"], g["out_coeff"])
    from symmer_b200.symmer_utils import gram_schmidt_from_quantum_state, random_anitcomm_2n_1_PauliwordOp
    for i in range(3):
        g = G[f"anticomm_{i}"]
        n, cplx, cliff, seed = [int(v) for v in g["args"]]
        np.random.seed(seed)
        AC = random_anitcomm_2n_1_PauliwordOp(n, complex_coeff=bool(cplx), apply_clifford=bool(cliff))
        _same_terms(AC.symp_matrix, AC.coeff_vec, g["symp"], g["coeff"])       # same RNG draws as the reference
        assert AC.n_terms == 2 * n + 1
        assert np.array_equal(AC.adjacency_matrix, np.eye(AC.n_terms, dtype=bool))
    g = G["gram_schmidt"]
    U = gram_schmidt_from_quantum_state(QuantumState(g["state"], g["coeff"]))
    assert np.allclose(U, g["out"], atol=1e-13) and np.allclose(U @ U.conj().T, np.eye(8), atol=1e-12)
    A = PauliwordOp.from_list(['XZ', 'YY'], [0.5, 2]).to_sparse_matrix
    assert matrix_allclose(A, A.copy()) and matrix_allclose(A, A.toarray()) and not matrix_allclose(A, 2 * A)


def case_circuit_symmerlator(api, G):
    from symmer_b200 import PauliwordOp
    from symmer_b200 import evolution as ev
    from symmer_b200.circuit_symmerlator import CircuitSymmerlator
    g = G["circuit_qasm"]
    O = PauliwordOp(g["o_symp"], g["o_coeff"])
    CS = CircuitSymmerlator.from_qasm(str(g["qasm"][0]))
    assert len(CS.sequence) == int(g["n_steps"][0])
    assert np.array_equal(np.vstack([p.symp_matrix for p, _ in CS.sequence]), g["seq_symp"])
    assert np.allclose([a for _, a in CS.sequence], g["seq_angle"], rtol=1e-15, atol=0)
    rot = CS.apply_sequence(O)
    _same_terms(rot.symp_matrix, rot.coeff_vec, g["rot_symp"], g["rot_coeff"])
    assert abs(CS.evaluate(O) - complex(g["expval"][0])) < 1e-12
    # Heisenberg picture: apply_sequence(P) = U^dagger P U with U the gate-library operator, on all two-qubit Paulis
    np.random.seed(5)
    theta = float(np.random.random())
    one = {'X': ev.X, 'Y': ev.Y, 'Z': ev.Z, 'H': ev.Had, 'S': ev.S}
    letters = ['I', 'X', 'Y', 'Z']
    for name, gate in one.items():
        CS = CircuitSymmerlator(2)
        getattr(CS, name)(1)
        U = gate(2, 1)
        for a in letters:
            for b in letters:
                P = PauliwordOp.from_list([a + b])
                assert CS.apply_sequence(P) == (U.dagger * P * U).cleanup(zero_threshold=1e-12), (name, a + b)
    for name, gate in {'CX': ev.CX, 'CY': ev.CY, 'CZ': ev.CZ}.items():
        CS = CircuitSymmerlator(2)
        getattr(CS, name)(0, 1)
        U = gate(2, 0, 1)
        for a in letters:
            for b in letters:
                P = PauliwordOp.from_list([a + b])
                assert CS.apply_sequence(P) == (U.dagger * P * U).cleanup(zero_threshold=1e-12), (name, a + b)
    for name, gate in {'RX': ev.RX, 'RY': ev.RY, 'RZ': ev.RZ}.items():
        CS = CircuitSymmerlator(1)
        getattr(CS, name)(0, theta)
        U = gate(1, 0, theta)
        for a in letters:
            P = PauliwordOp.from_list([a])
            assert CS.apply_sequence(P) == (U.dagger * P * U).cleanup(zero_threshold=1e-12), (name, a)
    # expectation value against the dense state U|0...0>
    CS = CircuitSymmerlator(3)
    CS.H(0); CS.CX(0, 1); CS.S(1); CS.sqrtY(2); CS.CZ(1, 2); CS.SWAP(0, 2); CS.RY(1, 0.3)
    np.random.seed(9)
    O = PauliwordOp.random(3, 20, complex_coeffs=False)
    gates = [ev.Had(3, 0), ev.CX(3, 0, 1), ev.S(3, 1), ev.RY(3, 2, -np.pi / 2), ev.CZ(3, 1, 2), ev.CX(3, 0, 2), ev.CX(3, 2, 0),
             ev.CX(3, 0, 2), ev.RY(3, 1, 0.3)]
    sv = np.zeros(8, dtype=complex)
    sv[0] = 1
    for gate in gates:
        sv = gate.to_sparse_matrix @ sv
    expect = np.vdot(sv, O.to_sparse_matrix @ sv)
    assert abs(CS.evaluate(O) - expect) < 1e-12, (CS.evaluate(O), expect)
    try:
        CS.Toffoli(0, 1, 2)
    except NotImplementedError:
        pass
    else:
        raise AssertionError("Toffoli is not implemented in the reference either")


def case_misc_methods(api, G):
    from symmer_b200 import PauliwordOp
    P = PauliwordOp.from_list(['XX', 'ZY', 'II'], [1, 2j, -0.5])
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        P.set_processing_method('single_thread')
        assert not caught
        P.set_processing_method('ray')
        assert len(caught) == 1
    # NumPy index conventions of base.py:894-928: negative entries wrap, out-of-range raises IndexError
    sub = P[[-1, 0]]
    assert np.array_equal(sub.symp_matrix, P.symp_matrix[[-1, 0]]) and np.array_equal(sub.coeff_vec, P.coeff_vec[[-1, 0]])
    assert P[-2:].n_terms == len(np.arange(-2, 3)) and P[np.array([True, False, True])].n_terms == 2
    try:
        P[[3]]
    except IndexError:
        pass
    else:
        raise AssertionError("index 3 of a 3-term operator must raise IndexError")
    from symmer_b200 import utils as u
    row, c = u.mul_symplectic(np.array([1, 0, 0, 0], dtype=bool), 1.0, np.array([0, 0, 1, 0], dtype=bool), 1.0)   # X * Z = -iY
    assert np.array_equal(row, [True, False, True, False]) and np.isclose(c, -1j)
    assert np.allclose(u.symplectic_to_sparse_matrix(np.array([1, 1], dtype=bool), 2.0).toarray(), [[0, -2j], [2j, 0]])
    assert u.safe_PauliwordOp_to_dict(P) == {'XX': (1.0, 0.0), 'ZY': (0.0, 2.0), 'II': (-0.5, 0.0)}
    assert u.count1_in_int_bitstring(0b1011) == 3 and np.array_equal(u.binary_array_to_int(np.array([[1, 0, 1], [0, 1, 1]])), [5, 3])
    assert np.isclose(np.linalg.norm(u.unit_n_sphere_cartesian_coords([0.3, 1.2, 2.0])), 1) and np.isclose(u.binomial_coefficient(5, 2), 10)
    contextual = PauliwordOp.from_list(['XI', 'IX', 'ZI', 'IZ', 'II'], [5, 4, 3, 2, 1])
    swept = u.perform_noncontextual_sweep(contextual)
    assert swept.is_noncontextual and swept.n_terms == 4 and not contextual.is_noncontextual
    df = P.to_dataframe()
    assert list(df['Pauli terms']) == ['XX', 'ZY', 'II']
    assert np.allclose(df['Coefficients (real)'], [1, 0, -0.5]) and np.allclose(df['Coefficients (imaginary)'], [0, 2, 0])
    assert 'Coefficients (imaginary)' not in PauliwordOp.from_list(['XZ'], [2]).to_dataframe().columns
    for attr in ['to_openfermion', 'to_qiskit']:
        try:
            getattr(P, attr)
        except ImportError:
            pass


CASES = [case_qwc, case_reindex, case_tensor, case_graphs, case_jordan, case_quantum_state_constructors,
         case_quantum_state_methods, case_projector_helpers, case_from_matrix,
         case_evolution_and_state_projection, case_symmetry_generators_device_path, case_tapering_intermediates,
         case_symmer_utils, case_circuit_symmerlator, case_misc_methods]
