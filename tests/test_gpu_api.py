"""GPU tests of the reference-facing API (symmer_b200.PauliwordOp / QuantumState / IndependentOp),
written after the reference's own tests (tests/test_operators/test_base.py etc.) and checked against
golden vectors produced by the real reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from oracle import pauli_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import symmer_b200
    from symmer_b200 import base, independent_op, utils
    return type("NS", (), dict(PauliwordOp=base.PauliwordOp, QuantumState=base.QuantumState,
                               IndependentOp=independent_op.IndependentOp, utils=utils, base=base))


def same_terms(op, symp, coeff, scale=1.0):
    ok, why = po.compare_term_sets(op.symp_matrix, op.coeff_vec, symp, coeff, scale=scale)
    assert ok, why


# ---- constructor contracts (reference tests/test_operators/test_base.py:26-130) ----------------
def test_constructor_contracts(sb):
    P = sb.PauliwordOp
    with pytest.raises(AssertionError):
        P([[0, 1, 2, 0]], [1])                       # not 0/1
    with pytest.raises(AssertionError):
        P([[0, 1, 1]], [1])                          # odd number of columns
    with pytest.raises(AssertionError):
        P([[0, 1, 1, 0]], [1, 2])                    # coefficient count mismatch
    with pytest.raises(TypeError):
        P([[0, 0, 1, 1]], 1)                         # scalar coefficient (test_base.py:99-110)
    with pytest.raises(AssertionError):
        P(np.zeros((2, 4), dtype=float), [1, 1])     # not bool/int
    op = P([0, 1, 1, 0], [2.0])                      # 1-D row is promoted
    assert op.n_terms == 1 and op.n_qubits == 2 and op.symp_matrix.dtype == bool
    assert op.coeff_vec.dtype == complex


def test_empty_and_cleanup_shapes(sb):
    P = sb.PauliwordOp
    E = P.empty(3)
    assert E.n_terms == 1 and E.n_qubits == 3 and np.array_equal(E.coeff_vec, np.array([0]))
    assert E == P([[0] * 6], [0])
    C = E.cleanup()
    assert C.n_qubits == 3 and C.symp_matrix.shape == (0, 6)
    Z = P(np.zeros((0, 6), dtype=bool), [])
    assert Z.cleanup().symp_matrix.shape == (1, 6)   # base.py:631-632


def test_cleanup_known_answers(sb):
    P = sb.PauliwordOp
    op = P.from_list(['XXX', 'YYY', 'XXX', 'YYY'], [1, 1, -1, 1])
    assert op.cleanup() == P.from_list(['YYY'], [2])
    assert P.from_list(['XXX', 'YYY', 'ZZZ'], [0, 0, 0]).cleanup().n_terms == 0
    np.random.seed(0)
    R = P.random(6, 30)
    assert (R + R) == R * 2
    assert (R - R).n_terms == 0


def test_single_qubit_products(sb):
    P = sb.PauliwordOp
    X, Y, Z = (P.from_list([s], [1]) for s in "XYZ")
    assert X * Y == P.from_list(['Z'], [1j])
    assert Y * X == P.from_list(['Z'], [-1j])
    assert Y * Z == P.from_list(['X'], [1j])
    assert Z * Y == P.from_list(['X'], [-1j])
    assert Z * X == P.from_list(['Y'], [1j])
    assert X * Z == P.from_list(['Y'], [-1j])
    assert (X ** 2) == P.from_list(['I'], [1]) and (X ** 0) == P.from_list(['I'], [1])


def test_products_against_reference(sb, golden):
    P = sb.PauliwordOp
    for nm in sorted(k for k in golden if k.startswith(("mul_rand_", "mul_single_"))):
        g = golden[nm]
        A, B = P(g["a_symp"], g["a_coeff"]), P(g["b_symp"], g["b_coeff"])
        same_terms(A * B, g["out_symp"], g["out_coeff"], scale=np.abs(g["a_coeff"]).max() * np.abs(g["b_coeff"]).max())
    for nm in sorted(k for k in golden if k.startswith("square_rand_")):
        g = golden[nm]
        A = P(g["a_symp"], g["a_coeff"])
        S = A * A
        same_terms(S, g["out_symp"], g["out_coeff"], scale=np.abs(g["a_coeff"]).max() ** 2)
        assert S.n_terms == len(g["out_coeff"])


def test_product_matches_dense_matmul(sb):
    P = sb.PauliwordOp
    np.random.seed(3)
    A, B = P.random(4, 12), P.random(4, 9)
    assert np.allclose((A * B).to_sparse_matrix.toarray(), A.to_sparse_matrix.toarray() @ B.to_sparse_matrix.toarray())
    assert np.allclose((A * 2.5j).to_sparse_matrix.toarray(), 2.5j * A.to_sparse_matrix.toarray())


def test_add_sub_against_reference(sb, golden):
    P = sb.PauliwordOp
    g = golden["add_rand"]
    A, B = P(g["a_symp"], g["a_coeff"]), P(g["b_symp"], g["b_coeff"])
    same_terms(A + B, g["sum_symp"], g["sum_coeff"])
    same_terms(A - B, g["diff_symp"], g["diff_coeff"])
    assert sum([A, B]) == A + B


def test_coefficient_write_through(sb):
    """Callers of the reference mutate coeff_vec in place (base.py:746); the device copy must follow."""
    P = sb.PauliwordOp
    op = P.from_list(['XX', 'ZZ'], [1, 2])
    op.coeff_vec[0] = 5
    op.coeff_vec *= -1
    assert (op * 1).coeff_vec.tolist() == [-5, -2]
    assert op + op == P.from_list(['XX', 'ZZ'], [-10, -4])


def test_adjacency_and_noncontextuality(sb, golden):
    P = sb.PauliwordOp
    for nm in sorted(k for k in golden if k.startswith("adj_ref_")):
        g = golden[nm]
        assert np.array_equal(P(g["symp"], np.ones(len(g["symp"]))).adjacency_matrix, g["adj"])
    g = golden["commute_ref_doc"]
    a, b = P(g["a_symp"], [1, 1]), P(g["b_symp"], [1, 1, 1])
    assert np.array_equal(a.commutes_termwise(b), g["out"])
    assert np.array_equal(a.anticommutes_termwise(b), ~g["out"])
    assert P.from_list(['ZZ', 'XX', 'YY', 'II'], [1, 1, 1, 1]).is_noncontextual
    assert not P.from_list(['IX', 'IZ', 'XI', 'ZI', 'XX', 'ZZ'], np.ones(6)).is_noncontextual   # Peres-Mermin type
    with pytest.raises(AssertionError):
        a.commutes_termwise(P.from_list(['XXX'], [1]))


def test_rotations_against_reference(sb, golden):
    P = sb.PauliwordOp
    names = sorted(k for k in golden if k.startswith(("rot_single_", "rot_seq_")))
    assert len(names) >= 43
    for nm in names:
        g = golden[nm]
        op = P(g["symp"], g["coeff"])
        rots = [(P(q.reshape(1, -1), [1]), None if np.isnan(a) else float(a)) for q, a in zip(g["q_symp"], g["angle"])]
        same_terms(op.perform_rotations(rots), g["out_symp"], g["out_coeff"], scale=np.abs(g["coeff"]).max())


def test_fused_rotation_path(sb, golden):
    """General rotations through ops.rotate_dedup (rotation + dedup as one block-list product, the path of
    operators with >= 2^15 terms), forced on the reference's golden cases, then on a 40 000-term operator
    against the oracle, including a sequence that mixes Clifford and general angles."""
    import symmer_b200.base as base
    P = sb.PauliwordOp
    old = base.FUSED_ROTATION_MIN_TERMS
    try:
        base.FUSED_ROTATION_MIN_TERMS = 0
        for nm in sorted(k for k in golden if k.startswith(("rot_single_", "rot_seq_"))):
            g = golden[nm]
            op = P(g["symp"], g["coeff"])
            rots = [(P(q.reshape(1, -1), [1]), None if np.isnan(a) else float(a)) for q, a in zip(g["q_symp"], g["angle"])]
            same_terms(op.perform_rotations(rots), g["out_symp"], g["out_coeff"], scale=np.abs(g["coeff"]).max())
        np.random.seed(5)
        op = P.random(3, 10)
        Q = P.from_list(['XYZ'], [1])
        for t in [0.37, -1.9, 2.5]:
            R = np.cos(t / 2) * np.eye(8) + 1j * np.sin(t / 2) * Q.to_sparse_matrix.toarray()
            expect = R @ op.to_sparse_matrix.toarray() @ R.conj().T
            assert np.allclose(op.perform_rotations([(Q, t)]).to_sparse_matrix.toarray(), expect)
    finally:
        base.FUSED_ROTATION_MIN_TERMS = old
    for n, M in [(1000, 40000), (70, 33000)]:
        s, c = po.random_operator(n, M, seed=n)
        s[M // 2:] = s[:M - M // 2]                              # duplicated rows: the fused dedup must merge them
        qs, _ = po.random_operator(n, 3, seed=n + 1)
        rots = [(qs[0], 0.37), (qs[1], np.pi / 2), (qs[2], -1.1)]
        ref_s, ref_c = po.perform_rotations(s, c, rots)
        out = P(s, c).perform_rotations([(P(q.reshape(1, -1), [1]), a) for q, a in rots])
        same_terms(out, ref_s, ref_c, scale=np.abs(c).max() * 2)


def test_rotation_is_conjugation(sb):
    """R P R^dagger with R = cos(t/2) I + i sin(t/2) Q, checked on dense matrices (reference
    tests/test_evolution/test_circuit_symmerlator.py style)."""
    P = sb.PauliwordOp
    np.random.seed(5)
    op = P.random(3, 10)
    Q = P.from_list(['XYZ'], [1])
    for t in [0.37, np.pi / 2, -1.9, np.pi]:
        R = np.cos(t / 2) * np.eye(8) + 1j * np.sin(t / 2) * Q.to_sparse_matrix.toarray()
        expect = R @ op.to_sparse_matrix.toarray() @ R.conj().T
        assert np.allclose(op.perform_rotations([(Q, t)]).to_sparse_matrix.toarray(), expect)


def test_sparse_matrix_against_reference(sb, golden):
    P = sb.PauliwordOp
    for nm in sorted(k for k in golden if k.startswith(("matrix_ref_", "matrix_rand_"))):
        g = golden[nm]
        if g["dense"].size:
            M = P(g["symp"], g["coeff"]).to_sparse_matrix
            assert M.shape == g["dense"].shape
            assert np.allclose(M.toarray(), g["dense"], rtol=1e-13, atol=1e-13), nm


def test_expval_hartree_fock(sb, golden, hamiltonians):
    for tag in ["H2O_STO3G", "Be_STO3G"]:
        symp, coeff, d = hamiltonians(tag)
        H = sb.PauliwordOp(symp, coeff)
        hf = sb.QuantumState(np.asarray(d["hf_array"]).reshape(1, -1))
        e = H.expval(hf)
        assert np.isclose(e, golden[f"hf_expval_{tag}"]["expval"][0].real, rtol=1e-12)
        assert np.isclose(e, d["hf_energy"][0], atol=1e-6)
        # symbolic route (bra * H * ket) agrees with the dense kernel
        assert np.isclose((hf.dagger * H * hf).real, e, rtol=1e-10)


def test_state_algebra(sb):
    QS, P = sb.QuantumState, sb.PauliwordOp
    psi = QS([[0, 0], [1, 1]], [1 / np.sqrt(2), 1 / np.sqrt(2)])
    assert np.isclose(psi.dagger * psi, 1.0)
    phi = P.from_list(['XI'], [1]) * psi                   # X on qubit 0: |10> + |01>
    assert set(phi.to_dictionary) == {"10", "01"}
    assert np.isclose(psi.dagger * phi, 0.0)
    y = P.from_list(['YI'], [1]) * QS([[0, 0]], [1])       # Y|0> = i|1>
    assert np.allclose(list(y.to_dictionary.values()), [1j])
    zz = P.from_list(['ZZ'], [1])
    assert np.isclose(zz.expval(psi), 1.0)
    wide = QS(np.eye(70, dtype=int)[:3], [0.6, 0.8j, 0.0])  # > 64 qubits: symbolic route, multi-word rows
    assert np.isclose(wide.dagger * wide, 1.0)
    assert np.isclose(P.from_list(['Z' + 'I' * 69], [1]).expval(wide), -0.36 + 0.64)


def test_symmetry_generators_config2(sb, golden, hamiltonians):
    symp, coeff, _ = hamiltonians("H2O_STO3G")
    H = sb.PauliwordOp(symp, coeff)
    S = sb.IndependentOp.symmetry_generators(H)
    g = golden["symgen_H2O_STO3G"]
    assert np.array_equal(S.symp_matrix, g["gen_symp"])           # bit-exact, same order as the reference
    assert set(po.to_strings(S.symp_matrix)) == {"IIIIIIIIZZIIII", "ZIZIIZZIIZZIIZ", "IZIZIZIZIZIZIZ",
                                                 "IIIIZZIIIIIIZZ"}
    assert np.array_equal(H.adjacency_matrix, g["adj"])
    assert np.all(S.commutes_termwise(H))
    r = golden["recon_H2O_STO3G"]
    gens = H.generators
    assert np.array_equal(gens.symp_matrix, r["gen_symp"])
    recon, mask = H.generator_reconstruction(gens)
    assert np.array_equal(recon, r["recon"]) and np.array_equal(mask, r["mask"])
    with pytest.raises(ValueError):
        sb.IndependentOp(np.array([[1, 0, 0, 0], [1, 0, 0, 0]], dtype=bool), [1, 1])   # dependent rows
    with pytest.raises(ValueError):
        sb.IndependentOp(np.array([[1, 0, 0, 0]], dtype=bool), [0.5])                  # coefficient not +/-1


def test_gf2_seams(sb, golden):
    u = sb.utils
    for nm in sorted(k for k in golden if k.startswith("gf2_rand_")):
        g = golden[nm]
        m = g["matrix"]
        assert np.array_equal(u._rref_binary(m), g["rref_norows"]), nm
        assert np.array_equal(u._cref_binary(m), g["cref_norows"]), nm
        if m.any():
            assert np.array_equal(u.rref_binary(m), g["rref"]), nm
            assert np.array_equal(u.cref_binary(m), g["cref"]), nm


def test_array_seams(sb, golden):
    u = sb.utils
    g = golden["cleanup_rand_1"]
    s, c = u.symplectic_cleanup(g["symp"], g["coeff"], zero_threshold=1e-15)
    assert np.array_equal(s, g["out_symp"]) and np.allclose(c, g["out_coeff"], rtol=1e-12)
    rng = np.random.default_rng(4)
    A, B = rng.random((37, 130)) < 0.4, rng.random((130, 21)) < 0.4
    assert np.array_equal(u.matmul_GF2(A, B), po.matmul_gf2(A, B))


def test_indexing_sort_iter(sb):
    P = sb.PauliwordOp
    op = P.from_list(['XX', 'YY', 'ZZ', 'IX'], [4, -3, 2, 1])
    assert op[1] == P.from_list(['YY'], [-3]) and op[-1] == P.from_list(['IX'], [1])
    assert op[1:3] == P.from_list(['YY', 'ZZ'], [-3, 2])
    assert op[[0, 3]] == P.from_list(['XX', 'IX'], [4, 1])
    assert [t.n_terms for t in op] == [1, 1, 1, 1]
    assert op.sort(by='magnitude').coeff_vec.real.tolist() == [4, -3, 2, 1]
    assert op.sort(by='magnitude', key='increasing').coeff_vec.real.tolist() == [1, 2, -3, 4]
    lex = op.sort(by='lex')
    assert np.array_equal(lex.symp_matrix, op.symp_matrix[np.lexsort(op.symp_matrix.T)])
    with pytest.raises(ValueError):
        op.sort(by='nonsense')
    assert op.to_dictionary == {'XX': 4, 'YY': -3, 'ZZ': 2, 'IX': 1}
    assert op.dagger == op and np.array_equal(op.Y_count, [0, 2, 0, 0])


def test_config1_square(sb):
    """Config C1 of BASELINE.json at reduced and full size: P*P for PauliwordOp.random(1000, 500)."""
    P = sb.PauliwordOp
    np.random.seed(1)
    op = P.random(1000, 500)
    S = op * op
    assert S.n_terms == 62415                       # commuting pairs + identity (seed 1)
    sub_s, sub_c = op.symp_matrix[:60], op.coeff_vec[:60]
    ref_s, ref_c = po.multiply(sub_s, sub_c, sub_s, sub_c)
    sub = P(sub_s, sub_c)
    same_terms(sub * sub, ref_s, ref_c, scale=np.abs(sub_c).max() ** 2)


def test_single_pauli_product_at_a_million_qubits(sb):
    """README claim #4 of the reference (single Pauli x single Pauli on very wide registers)."""
    P = sb.PauliwordOp
    n = 1_000_000
    rng = np.random.default_rng(0)
    a = rng.random((1, 2 * n)) < 0.3
    b = rng.random((1, 2 * n)) < 0.3
    A, B = P(a, [1.5]), P(b, [2.0j])
    C = A * B
    ref_s, ref_c = po.multiply(a, np.array([1.5]), b, np.array([2.0j]))
    assert C.n_terms == 1 and np.array_equal(C.symp_matrix, ref_s) and np.allclose(C.coeff_vec, ref_c, rtol=1e-14)
    assert bool(A.commutes_termwise(B)[0, 0]) == bool(po.commutes_termwise(a, b)[0, 0])
    assert (A + A).n_terms == 1 and np.allclose((A + A).coeff_vec, [3.0])


def test_first_occurrence_vs_sorted_hash_switch(sb):
    """Products just below / above the 2^22 cross-term switch agree as term sets; below it the row
    order is the reference's first-occurrence order."""
    P = sb.PauliwordOp
    np.random.seed(11)
    A, B = P.random(40, 2048), P.random(40, 2049)
    lo = A * B[:2047]                                   # 4 192 256 < 2^22
    hi = B * A                                          # 4 196 352 > 2^22 (B is the larger operand)
    ref_s, ref_c = po.multiply(A.symp_matrix, A.coeff_vec, B.symp_matrix[:2047], B.coeff_vec[:2047])
    assert np.array_equal(lo.symp_matrix, ref_s)        # same order as the reference
    assert np.allclose(lo.coeff_vec, ref_c, rtol=1e-12)
    ref_s, ref_c = po.multiply(B.symp_matrix, B.coeff_vec, A.symp_matrix, A.coeff_vec)
    same_terms(hi, ref_s, ref_c, scale=float(np.abs(A.coeff_vec).max() * np.abs(B.coeff_vec).max()))
    if hi.n_terms == len(ref_c):                        # above the switch: ordered-tile mode, same order too
        assert np.array_equal(hi.symp_matrix, ref_s)


def test_molecular_square_heavy_duplication(sb, hamiltonians):
    """H*H for H2O: 1.18M cross terms collapse to a few tens of thousands (group sizes ~ 50)."""
    symp, coeff, _ = hamiltonians("H2O_STO3G")
    H = sb.PauliwordOp(symp, coeff)
    H2 = H * H
    ref_s, ref_c = po.multiply(symp, coeff, symp, coeff)
    assert H2.n_terms == len(ref_c)
    same_terms(H2, ref_s, ref_c, scale=float(np.abs(coeff).max() ** 2))


def test_zero_qubit_and_empty_operands(sb):
    P = sb.PauliwordOp
    Z = P(np.zeros((3, 0), dtype=bool), [1, 2, 3])
    assert Z.n_qubits == 0 and np.allclose(Z.cleanup().coeff_vec, [6])
    E = P(np.zeros((0, 8), dtype=bool), [])
    A = P.from_list(['XYZI'], [1])
    assert (A * E).n_terms == 0 and (E * A).n_terms == 0
    assert (A + E) == A
    assert E.commutes_termwise(A).shape == (0, 1)


def test_generator_reconstruction_device_path(sb):
    """generator_reconstruction (base.py:523-560) on the device: bit-transposed packed rows, blocked
    GF(2) reduction, R and the mask only travel back. Cases: full span, partial span (mask has False
    entries and R rows are only partially built), padding bit positions (n not a multiple of 64), a
    matrix wide enough for the blocked multi-launch path."""
    rng = np.random.default_rng(12)
    for n, n_gen, M in [(6, 4, 50), (40, 11, 300), (100, 30, 2000), (300, 25, 3000), (1000, 12, 200)]:
        gens = rng.random((n_gen, 2 * n)) < 0.3
        assert po.check_independent(gens)
        combos = rng.random((M, n_gen)) < 0.4
        symp = (combos.astype(np.uint8) @ gens.astype(np.uint8) % 2).astype(bool)   # inside the span
        outside = rng.random(M) < 0.2
        symp[outside] ^= rng.random((int(outside.sum()), 2 * n)) < 0.05            # mostly outside the span
        G = sb.PauliwordOp(gens, np.ones(n_gen))
        P = sb.PauliwordOp(symp, np.ones(M))
        recon, mask = P.generator_reconstruction(G)
        ref_recon, ref_mask = po.generator_reconstruction(gens, symp)
        assert recon.shape == (M, n_gen) and recon.dtype == ref_recon.dtype
        assert np.array_equal(mask, ref_mask), n
        assert np.array_equal(recon, ref_recon), n                                   # bit-exact, failed rows included
        assert mask[~outside].all() and not mask.all()
        ok = mask
        assert np.array_equal((recon[ok] @ gens.astype(int)) % 2, symp[ok].astype(int))   # M = R B on the span
    with pytest.raises(AssertionError):
        dep = sb.PauliwordOp(np.vstack([gens[:2], gens[0] ^ gens[1]]), np.ones(3))
        P.generator_reconstruction(dep)


@pytest.mark.parametrize("tag", ["H2O_STO3G", "Be_STO3G", "NH3_STO3G", "HOOH_STO3G"])
def test_qubit_tapering_matches_reference(sb, taper_golden, hamiltonians, tag):
    """QubitTapering(H).taper_it (symmetry generators -> sector from the Hartree-Fock state -> Clifford
    rotations -> device projection) against vectors written by the real reference: every
    intermediate (generators, sector, the rotation list in order, rotated stabilizers, free qubits)
    bit-exact, the tapered operator as a term set."""
    from symmer_b200.projection import QubitTapering
    symp, coeff, d = hamiltonians(tag)
    for sqp in ["Z", "X"]:
        g = taper_golden[f"taper_{tag}_{sqp}"]
        H = sb.PauliwordOp(symp, coeff)
        qt = QubitTapering(H, target_sqp=sqp)
        assert qt.n_taper == g["gen_symp"].shape[0]
        assert np.array_equal(qt.symmetry_generators.symp_matrix, g["gen_symp"])
        out = qt.taper_it(ref_state=g["hf"])
        assert np.array_equal(qt.stabilizers.coeff_vec.real, g["sector"])
        rot = np.array([r.symp_matrix[0] for r, _ in qt.stabilizers.stabilizer_rotations]).reshape(-1, symp.shape[1])
        assert np.array_equal(rot, g["rotations"]), (tag, sqp)
        assert np.array_equal(qt.rotated_stabilizers.symp_matrix, g["rotated_symp"])
        assert np.array_equal(qt.rotated_stabilizers.coeff_vec.real, g["rotated_coeff"])
        assert np.array_equal(qt.free_qubit_indices, g["free"])
        assert out.n_qubits == int(g["n_out_qubits"][0]) and out.n_terms == g["out_symp"].shape[0]
        same_terms(out, g["out_symp"], g["out_coeff"], scale=np.abs(coeff).max())
    g = taper_golden[f"taper_{tag}_sector"]
    out = QubitTapering(sb.PauliwordOp(symp, coeff)).taper_it(sector=g["sector"])
    same_terms(out, g["out_symp"], g["out_coeff"], scale=np.abs(coeff).max())


def test_projection_kernel_against_oracle():
    """sym_project on wide random operators (1000 qubits, free-qubit count not a multiple of 64, mixed
    X/Z stabilizers, duplicates created by the qubit removal) and the all-qubits-stabilized corner."""
    import torch
    from symmer_b200 import ops
    rng = np.random.default_rng(8)
    for n, M, S in [(1000, 5000, 37), (70, 3000, 70), (64, 200, 1), (5, 400, 3)]:
        symp, coeff = po.random_operator(n, M, seed=n)
        qubits = rng.choice(n, size=S, replace=False)
        is_x = rng.random(S) < 0.5
        stab = np.zeros((S, 2 * n), dtype=bool)
        stab[np.arange(S), np.where(is_x, qubits, qubits + n)] = True
        eig = rng.choice([-1.0, 1.0], size=S)
        free = np.setdiff1d(np.arange(n), qubits)
        # make a fair share of rows commute with every stabilizer
        ok_rows = rng.random(M) < 0.5
        symp[np.ix_(ok_rows, qubits[is_x] + n)] = False
        symp[np.ix_(ok_rows, qubits[~is_x])] = False
        xz = ops.pack(torch.from_numpy(symp), n)
        c = torch.from_numpy(coeff).cuda()
        cols = np.where(stab)[1]
        pxz, pc = ops.project(xz, c, n, cols, eig, free)
        keep = np.all(po.commutes_termwise(symp, stab), axis=1)
        assert pxz.shape[0] == int(keep.sum()) and keep.sum() >= ok_rows.sum()
        ref_s, ref_c = po.project_onto_stabilizers(symp, coeff, stab, eig, free)
        if len(free):
            oxz, oc = ops.cleanup(pxz, pc)
            s = ops.unpack(oxz, len(free)).cpu().numpy()
            ok, why = po.compare_term_sets(s, oc.cpu().numpy(), ref_s, ref_c, scale=np.abs(coeff).max())
            assert ok, (n, why)
        else:
            assert np.isclose(complex(pc.sum().cpu().numpy()), ref_c[0], rtol=1e-12)
