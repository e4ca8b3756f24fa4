"""The CPU oracle (oracle/pauli_oracle.py) against the golden vectors produced by the real reference
(tests/golden/make_golden.py). This is what pins the oracle. CPU only."""
import numpy as np
import pytest

from oracle import pauli_oracle as po


def _names(golden, prefix):
    return sorted(k for k in golden if k.startswith(prefix))


def test_golden_file_is_populated(golden):
    assert len(golden) > 100


def test_multiply_matches_reference_bit_for_bit(golden):
    names = _names(golden, "mul_single_") + _names(golden, "mul_rand_")
    assert len(names) >= 20
    for nm in names:
        g = golden[nm]
        s, c = po.multiply(g["a_symp"], g["a_coeff"], g["b_symp"], g["b_coeff"])
        # same algorithm, same order: rows and coefficients must be identical, not merely close
        assert np.array_equal(s, g["out_symp"]), nm
        assert np.array_equal(c, g["out_coeff"]), nm


def test_square_matches_reference(golden):
    for nm in _names(golden, "square_rand_"):
        g = golden[nm]
        s, c = po.multiply(g["a_symp"], g["a_coeff"], g["a_symp"], g["a_coeff"])
        assert np.array_equal(s, g["out_symp"]), nm
        assert np.array_equal(c, g["out_coeff"]), nm


def test_cleanup_matches_reference(golden):
    for nm in _names(golden, "cleanup_"):
        g = golden[nm]
        s, c = po.cleanup(g["symp"], g["coeff"])
        assert s.shape == g["out_symp"].shape, nm
        assert np.array_equal(s, g["out_symp"]), nm
        assert np.array_equal(c, g["out_coeff"]), nm


def test_cleanup_reference_known_answers(golden):
    g = golden["cleanup_ref_1"]                       # ['XXX','YYY','XXX','YYY'],[1,1,-1,1] == 2*YYY
    s, c = po.cleanup(g["symp"], g["coeff"])
    assert po.to_strings(s) == ["YYY"] and c[0] == 2
    g = golden["cleanup_ref_zero"]
    s, c = po.cleanup(g["symp"], g["coeff"])
    assert s.shape == (0, 6) and c.shape == (0,)


def test_add_sub(golden):
    g = golden["add_rand"]
    s, c = po.cleanup(np.vstack([g["a_symp"], g["b_symp"]]), np.hstack([g["a_coeff"], g["b_coeff"]]))
    assert np.array_equal(s, g["sum_symp"]) and np.array_equal(c, g["sum_coeff"])
    s, c = po.cleanup(np.vstack([g["a_symp"], g["b_symp"]]), np.hstack([g["a_coeff"], -g["b_coeff"]]))
    assert np.array_equal(s, g["diff_symp"]) and np.array_equal(c, g["diff_coeff"])


def test_commute_matches_reference(golden):
    names = _names(golden, "commute_")
    assert len(names) >= 8
    for nm in names:
        g = golden[nm]
        assert np.array_equal(po.commutes_termwise(g["a_symp"], g["b_symp"]), g["out"]), nm
    for nm in _names(golden, "adj_ref_"):
        g = golden[nm]
        assert np.array_equal(po.commutes_termwise(g["symp"], g["symp"]), g["adj"]), nm


def test_rotations_match_reference(golden):
    names = _names(golden, "rot_single_") + _names(golden, "rot_seq_")
    assert len(names) >= 40
    for nm in names:
        g = golden[nm]
        rots = [(q, None if np.isnan(a) else float(a)) for q, a in zip(g["q_symp"], g["angle"])]
        s, c = po.perform_rotations(g["symp"], g["coeff"], rots)
        assert np.array_equal(s, g["out_symp"]), nm
        assert np.allclose(c, g["out_coeff"], rtol=1e-14, atol=0), nm


def test_sparse_matrix_matches_reference(golden):
    for nm in _names(golden, "matrix_ref_") + _names(golden, "matrix_rand_"):
        g = golden[nm]
        M = po.to_sparse_matrix(g["symp"], g["coeff"])
        if g["dense"].size:
            assert np.allclose(M.toarray(), g["dense"], rtol=1e-13, atol=1e-13), nm
            single = sum(po.single_term_matrix(r, c) for r, c in zip(g["symp"], g["coeff"]))
            assert np.allclose(single.toarray(), g["dense"], rtol=1e-13, atol=1e-13), nm
        if "psi" in g:
            assert np.allclose(M @ g["psi"], g["Hpsi"], rtol=1e-12, atol=1e-13), nm
            assert np.allclose(po.pauli_apply_dense(g["symp"], g["coeff"], g["psi"]), g["Hpsi"],
                               rtol=1e-12, atol=1e-13), nm
            assert np.isclose(po.expval_dense(g["symp"], g["coeff"], g["psi"]), g["expval"][0], rtol=1e-12)


def test_gf2_matches_reference(golden):
    names = _names(golden, "gf2_rand_")
    assert len(names) >= 10
    for nm in names:
        g = golden[nm]
        m = g["matrix"]
        assert np.array_equal(po._rref_binary(m), g["rref_norows"]), nm
        assert np.array_equal(po._cref_binary(m), g["cref_norows"]), nm
        if m.any():
            assert np.array_equal(po.rref_binary(m), g["rref"]), nm
            assert np.array_equal(po.cref_binary(m), g["cref"]), nm


def test_symmetry_generators_match_reference(golden, hamiltonians):
    for tag in ["H2O_STO3G", "Be_STO3G"]:
        symp, coeff, _ = hamiltonians(tag)
        g = golden[f"symgen_{tag}"]
        assert np.array_equal(symp, g["symp"])
        S = po.symmetry_generator_rows(symp)
        assert np.array_equal(S, g["gen_symp"]), tag
        assert np.array_equal(po.commutes_termwise(symp, symp), g["adj"]), tag
        r = golden[f"recon_{tag}"]
        recon, mask = po.generator_reconstruction(r["gen_symp"], symp)
        assert np.array_equal(recon, r["recon"]) and np.array_equal(mask, r["mask"])
    # config 2 known answer
    symp, _, _ = hamiltonians("H2O_STO3G")
    assert set(po.to_strings(po.symmetry_generator_rows(symp))) == {
        "IIIIIIIIZZIIII", "ZIZIIZZIIZZIIZ", "IZIZIZIZIZIZIZ", "IIIIZZIIIIIIZZ"}


def test_hf_expval_matches_reference(golden, hamiltonians):
    for tag in ["H2O_STO3G", "Be_STO3G"]:
        symp, coeff, d = hamiltonians(tag)
        n = symp.shape[1] // 2
        psi = np.zeros(1 << n, dtype=complex)
        psi[int(golden[f"hf_expval_{tag}"]["psi_index"][0])] = 1.0
        e = po.expval_dense(symp, coeff, psi)
        assert np.isclose(e.real, golden[f"hf_expval_{tag}"]["expval"][0].real, rtol=1e-12), tag
        assert np.isclose(e.real, d["hf_energy"][0], atol=1e-6), tag


def test_unordered_unique_c_and_numpy_agree():
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 2, size=(500, 9)).astype("uint16")
    first_c, inv_c = po.unordered_unique(rows)
    saved, po._CLIB = po._CLIB, False
    try:
        first_n, inv_n = po.unordered_unique(rows)
    finally:
        po._CLIB = saved
    assert np.array_equal(first_c, first_n) and np.array_equal(inv_c, inv_n)


def test_pack_unpack_roundtrip():
    rng = np.random.default_rng(1)
    for n in [1, 5, 63, 64, 65, 128, 1000]:
        symp = rng.random((7, 2 * n)) < 0.3
        xz = po.pack_bits(symp)
        assert xz.shape == (7, 2 * ((n + 63) // 64))
        assert np.array_equal(po.unpack_bits(xz, n), symp)
        q = n - 1
        assert bool((xz[0, q // 64] >> np.uint64(q % 64)) & np.uint64(1)) == bool(symp[0, q])


def test_projection_oracle_against_reference_tapering(taper_golden, hamiltonians):
    """The oracle's rotations + stabilizer-subspace projection reproduce the real reference's
    QubitTapering.taper_it outputs (H2O, Be, NH3; target Pauli Z and X)."""
    for tag in ["H2O_STO3G", "Be_STO3G", "NH3_STO3G"]:
        symp, coeff, _ = hamiltonians(tag)
        for sqp in ["Z", "X"]:
            g = taper_golden[f"taper_{tag}_{sqp}"]
            s, c = po.taper(symp, coeff, g["rotations"], g["rotated_symp"], g["rotated_coeff"], g["free"])
            assert s.shape == g["out_symp"].shape, (tag, sqp)
            ok, why = po.compare_term_sets(s, c, g["out_symp"], g["out_coeff"], scale=np.abs(coeff).max())
            assert ok, (tag, sqp, why)
