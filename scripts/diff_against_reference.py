"""Development check (build container only: needs /root/reference): differential run of the reference-facing API
against the REAL reference (imported through oracle/shim) on seeded random inputs — same calls on both sides,
results compared field by field. On the CPU box the kernels are swapped for the NumPy test double
(tests/_host_double.py), so this exercises the host logic; pass --device on a B200 to run the CUDA kernels.

    python scripts/diff_against_reference.py [--device] [--rounds N]
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle", "shim"), "/root/reference"]
warnings.simplefilter("ignore")
if not hasattr(np, "product"):
    np.product = np.prod


def main():
    if not os.path.isdir("/root/reference/symmer"):
        print("reference not available on this machine")
        return 0
    if "--device" not in sys.argv:
        from _host_double import host_double
        main.double = host_double()
        main.double.__enter__()
    rounds = int(sys.argv[sys.argv.index("--rounds") + 1]) if "--rounds" in sys.argv else 12
    import symmer as ref
    from symmer import process
    process.method = 'single_thread'
    from symmer.operators import IndependentOp as RefInd
    import symmer_b200 as new
    from symmer_b200 import IndependentOp as NewInd

    checks = [0]
    failures = []

    def same_op(tag, a, b, ordered=True, tol=1e-12):
        """a: reference object, b: engine object."""
        checks[0] += 1
        sa, ca, sb_, cb = a.symp_matrix, a.coeff_vec, b.symp_matrix, b.coeff_vec
        ok = sa.shape == sb_.shape
        if ok and ordered:
            ok = np.array_equal(sa, sb_) and np.allclose(ca, cb, rtol=tol, atol=tol)
        elif ok:
            ok = {tuple(r): 0 for r in sa.tolist()}.keys() == {tuple(r): 0 for r in sb_.tolist()}.keys()
            if ok:
                da = {tuple(r): c for r, c in zip(sa.tolist(), ca)}
                db = {tuple(r): c for r, c in zip(sb_.tolist(), cb)}
                ok = all(abs(da[k] - db[k]) <= tol * max(1, abs(da[k])) for k in da)
        if not ok:
            failures.append(tag)

    def same(tag, a, b):
        checks[0] += 1
        if isinstance(a, (np.ndarray, list, tuple)) or isinstance(b, (np.ndarray, list, tuple)):
            ok = np.shape(a) == np.shape(b) and np.allclose(np.asarray(a, dtype=complex), np.asarray(b, dtype=complex),
                                                           rtol=1e-12, atol=1e-12)
        elif isinstance(a, (int, float, complex, np.number)) and not isinstance(a, (bool, np.bool_)):
            ok = bool(np.isclose(a, b, rtol=1e-12, atol=1e-12))
        else:
            ok = a == b
        if not ok:
            failures.append(f"{tag}: {a!r} vs {b!r}"[:300])

    for rnd in range(rounds):
        n = int(np.random.default_rng(rnd).integers(1, 9))
        m = int(np.random.default_rng(rnd + 100).integers(1, 14))
        np.random.seed(rnd)
        A_r = ref.PauliwordOp.random(n, m)
        np.random.seed(rnd)
        A_n = new.PauliwordOp.random(n, m)
        np.random.seed(rnd + 1000)
        B_r = ref.PauliwordOp.random(n, max(1, m // 2), complex_coeffs=False)
        np.random.seed(rnd + 1000)
        B_n = new.PauliwordOp.random(n, max(1, m // 2), complex_coeffs=False)
        same_op(f"{rnd} random", A_r, A_n)
        for by in ['magnitude', 'weight', 'support', 'X', 'Y', 'Z', 'lex']:
            for key in ['decreasing', 'increasing']:
                # ties are broken by position in both; coefficients are continuous, rows may repeat
                same_op(f"{rnd} sort {by} {key}", A_r.cleanup().sort(by=by, key=key), A_n.cleanup().sort(by=by, key=key),
                        ordered=(by in ('magnitude', 'lex')))
        same_op(f"{rnd} cleanup", A_r.cleanup(), A_n.cleanup())
        same_op(f"{rnd} mul", A_r * B_r, A_n * B_n, ordered=False)
        same_op(f"{rnd} add", A_r + B_r, A_n + B_n, ordered=False)
        same_op(f"{rnd} sub", A_r - B_r, A_n - B_n, ordered=False)
        same_op(f"{rnd} pow2", A_r ** 2, A_n ** 2, ordered=False)
        same_op(f"{rnd} pow0", A_r ** 0, A_n ** 0)
        same_op(f"{rnd} commutator", A_r.commutator(B_r), A_n.commutator(B_n), ordered=False)
        same_op(f"{rnd} anticommutator", A_r.anticommutator(B_r), A_n.anticommutator(B_n), ordered=False)
        same(f"{rnd} commutes", bool(A_r.commutes(B_r)), bool(A_n.commutes(B_n)))
        same_op(f"{rnd} dagger", A_r.dagger, A_n.dagger)
        same_op(f"{rnd} const", A_r.multiply_by_constant(0.3 - 2j), A_n.multiply_by_constant(0.3 - 2j))
        same_op(f"{rnd} scalar mul", A_r * 2.5, A_n * 2.5)
        same_op(f"{rnd} append", A_r.append(B_r), A_n.append(B_n))
        same(f"{rnd} Y_count", A_r.Y_count, A_n.Y_count)
        same(f"{rnd} commutes_termwise", A_r.commutes_termwise(B_r), A_n.commutes_termwise(B_n))
        same(f"{rnd} anticommutes_termwise", A_r.anticommutes_termwise(B_r), A_n.anticommutes_termwise(B_n))
        same(f"{rnd} qwc", A_r.qubitwise_commutes_termwise(B_r), A_n.qubitwise_commutes_termwise(B_n))
        same(f"{rnd} adjacency", A_r.adjacency_matrix, A_n.adjacency_matrix)
        same(f"{rnd} noncontextual", bool(A_r.is_noncontextual), bool(A_n.is_noncontextual))
        same(f"{rnd} str", str(A_r), str(A_n))
        same(f"{rnd} dict", A_r.to_dictionary, A_n.to_dictionary)
        same(f"{rnd} hash-eq", A_r == B_r, A_n == B_n)
        same_op(f"{rnd} generators", A_r.generators, A_n.generators)
        same_op(f"{rnd} getitem slice", A_r[1:], A_n[1:])
        same_op(f"{rnd} getitem list", A_r[[0, -1]], A_n[[0, -1]])
        same(f"{rnd} sparse", A_r.to_sparse_matrix.toarray(), A_n.to_sparse_matrix.toarray())
        g_r = A_r.generators
        g_n = A_n.generators
        rr, mr = A_r.generator_reconstruction(g_r)
        rn, mn = A_n.generator_reconstruction(g_n)
        same(f"{rnd} recon", rr, rn)
        same(f"{rnd} recon mask", mr, mn)
        Q_r = ref.PauliwordOp(B_r.symp_matrix[0], [1])
        Q_n = new.PauliwordOp(B_n.symp_matrix[0], [1])
        for angle in [None, 0.37, np.pi, -np.pi / 2, 3 * np.pi / 2]:
            same_op(f"{rnd} rotate {angle}", A_r.perform_rotations([(Q_r, angle)]), A_n.perform_rotations([(Q_n, angle)]),
                    ordered=False)
        same_op(f"{rnd} tensor", A_r.tensor(B_r), A_n.tensor(B_n), ordered=False)
        from symmer.operators import utils as ru
        from symmer_b200 import utils as nu
        same(f"{rnd} single csr", ru.symplectic_to_sparse_matrix(A_r.symp_matrix[0], 0.7 - 0.2j).toarray(),
             nu.symplectic_to_sparse_matrix(A_n.symp_matrix[0], 0.7 - 0.2j).toarray())
        v_r, c_r = ru.mul_symplectic(A_r.symp_matrix[0], 0.5j, B_r.symp_matrix[0], 2.0)
        v_n, c_n = nu.mul_symplectic(A_n.symp_matrix[0], 0.5j, B_n.symp_matrix[0], 2.0)
        same(f"{rnd} mul_symplectic row", v_r, v_n)
        same(f"{rnd} mul_symplectic coeff", c_r, c_n)
        same(f"{rnd} safe dict", {k: complex(*v) for k, v in ru.safe_PauliwordOp_to_dict(A_r).items()},
             {k: complex(*v) for k, v in nu.safe_PauliwordOp_to_dict(A_n).items()})
        same(f"{rnd} bits to int", ru.binary_array_to_int(A_r.symp_matrix), nu.binary_array_to_int(A_n.symp_matrix))
        same(f"{rnd} popcount", [ru.count1_in_int_bitstring(v) for v in (0, 1, 255, 2 ** 31 + 5, rnd * 977)],
             [nu.count1_in_int_bitstring(v) for v in (0, 1, 255, 2 ** 31 + 5, rnd * 977)])
        ang = np.random.default_rng(rnd).random(4) * 3
        same(f"{rnd} sphere", ru.unit_n_sphere_cartesian_coords(ang), nu.unit_n_sphere_cartesian_coords(ang))
        same(f"{rnd} binomial", ru.binomial_coefficient(4.5, 3), nu.binomial_coefficient(4.5, 3))
        same_op(f"{rnd} noncontextual sweep", ru.perform_noncontextual_sweep(A_r.cleanup().sort()),
                nu.perform_noncontextual_sweep(A_n.cleanup().sort()))
        # states
        np.random.seed(rnd + 7)
        psi_r = ref.QuantumState.random(n, 5)
        np.random.seed(rnd + 7)
        psi_n = new.QuantumState.random(n, 5)
        same(f"{rnd} state dict", {k: np.round(v, 12) for k, v in psi_r.to_dictionary.items()},
             {k: np.round(v, 12) for k, v in psi_n.to_dictionary.items()})
        same(f"{rnd} expval", A_r.expval(psi_r), A_n.expval(psi_n))
        same(f"{rnd} bra-ket", psi_r.dagger * psi_r, psi_n.dagger * psi_n)
        same(f"{rnd} op-ket dict", {k: np.round(v, 10) for k, v in (A_r * psi_r).to_dictionary.items()},
             {k: np.round(v, 10) for k, v in (A_n * psi_n).to_dictionary.items()})
        same(f"{rnd} state str", str(psi_r.sort()), str(psi_n.sort()))
        same(f"{rnd} dense", psi_r.to_dense_matrix, psi_n.to_dense_matrix)
        same(f"{rnd} normalized", psi_r._is_normalized(), psi_n._is_normalized())
    # ---- structured cases: duplicates and vanishing coefficients, stabilizer workflows, clique covers, states
    import json
    from symmer import QubitTapering as RefQT
    from symmer_b200 import QubitTapering as NewQT

    def both(fn_r, fn_n):
        return fn_r(), fn_n()

    for rnd in range(rounds):
        n = 3 + rnd % 4
        np.random.seed(500 + rnd)
        base_symp = np.random.rand(6, 2 * n) < 0.4
        symp = np.vstack([base_symp, base_symp[[0, 2, 2]]])
        coeff = np.random.randn(9) + 1j * np.random.randn(9)
        coeff[1] = 0
        coeff[7] = -coeff[2]                                   # cancels with its duplicate ... partially (8 also dups 2)
        D_r, D_n = ref.PauliwordOp(symp, coeff), new.PauliwordOp(symp, coeff)
        same_op(f"{rnd} dup cleanup", D_r.cleanup(), D_n.cleanup())
        same_op(f"{rnd} dup cleanup thr", D_r.cleanup(zero_threshold=0.5), D_n.cleanup(zero_threshold=0.5))
        same_op(f"{rnd} dup square", D_r * D_r, D_n * D_n, ordered=False)
        same(f"{rnd} dup dict", D_r.to_dictionary, D_n.to_dictionary)
        same(f"{rnd} dup eq", D_r == D_r.cleanup(), D_n == D_n.cleanup())
        same(f"{rnd} dup adjacency", D_r.adjacency_matrix, D_n.adjacency_matrix)
        same(f"{rnd} dup qwc", D_r.adjacency_matrix_qwc, D_n.adjacency_matrix_qwc)
        for rel in ['C', 'AC', 'QWC']:
            for strategy in ['largest_first', 'sorted_insertion', 'independent_set', 'DSATUR']:
                cr, cn = D_r.cleanup().clique_cover(rel, strategy), D_n.cleanup().clique_cover(rel, strategy)
                same(f"{rnd} cover keys {rel} {strategy}", sorted(cr.keys()), sorted(cn.keys()))
                for k in cr:
                    if k in cn:
                        same_op(f"{rnd} cover {rel} {strategy} {k}", cr[k], cn[k], ordered=False)
            same_op(f"{rnd} largest clique {rel}", D_r.cleanup().largest_clique(rel), D_n.cleanup().largest_clique(rel),
                    ordered=False)
        # a commuting independent set: Clifford images of Z_0..Z_{k-1}
        k = 1 + rnd % n
        z = np.zeros((k, 2 * n), dtype=bool)
        z[np.arange(k), n + np.arange(k)] = True
        signs = np.random.choice([1, -1], size=k)
        S_r, S_n = ref.PauliwordOp(z, signs), new.PauliwordOp(z, signs)
        rots_r, rots_n = [], []
        for _ in range(2 * n):
            q = np.random.rand(2 * n) < 0.5
            if not q.any():
                q[0] = True
            rots_r.append((ref.PauliwordOp(q, [1]), None))
            rots_n.append((new.PauliwordOp(q, [1]), None))
        S_r, S_n = S_r.perform_rotations(rots_r), S_n.perform_rotations(rots_n)
        same_op(f"{rnd} clifford images", S_r, S_n, ordered=False)
        for sqp in ['Z', 'X']:
            I_r, I_n = RefInd(S_r.symp_matrix, S_r.coeff_vec, target_sqp=sqp), NewInd(S_r.symp_matrix, S_r.coeff_vec, target_sqp=sqp)
            R_r, R_n = I_r.rotate_onto_single_qubit_paulis(), I_n.rotate_onto_single_qubit_paulis()
            same_op(f"{rnd} onto sqp {sqp}", R_r, R_n)
            same(f"{rnd} rotation list {sqp}", np.array([p.symp_matrix[0] for p, _ in I_r.stabilizer_rotations]).reshape(-1, 2 * n),
                 np.array([p.symp_matrix[0] for p, _ in I_n.stabilizer_rotations]).reshape(-1, 2 * n))
            same_op(f"{rnd} ind getitem", I_r[0], I_n[0])
            same_op(f"{rnd} ind perform_rotations", I_r.perform_rotations(I_r.stabilizer_rotations),
                    I_n.perform_rotations(I_n.stabilizer_rotations))
        # states
        np.random.seed(900 + rnd)
        p_r, q_r = ref.QuantumState.random(n, 6), ref.QuantumState.random(n, 4)
        np.random.seed(900 + rnd)
        p_n, q_n = new.QuantumState.random(n, 6), new.QuantumState.random(n, 4)

        def sd(state):
            return {kk: np.round(vv, 10) for kk, vv in state.to_dictionary.items()}
        same(f"{rnd} state add", sd(p_r + q_r), sd(p_n + q_n))
        same(f"{rnd} state sub", sd(p_r - q_r), sd(p_n - q_n))
        same(f"{rnd} state scalar", sd(p_r * 0.5j), sd(p_n * 0.5j))
        same(f"{rnd} state normalize", sd((p_r + q_r).normalize), sd((p_n + q_n).normalize))
        same(f"{rnd} state overlap", q_r.dagger * p_r, q_n.dagger * p_n)
        same(f"{rnd} state sort", str(p_r.sort(key='support')), str(p_n.sort(key='support')))
        same(f"{rnd} state rdm", p_r.get_rdm([0]), p_n.get_rdm([0]))
        M_r, M_n = both(lambda: ref.PauliwordOp.from_matrix(D_r.to_sparse_matrix.toarray(), disable_loading_bar=True),
                        lambda: new.PauliwordOp.from_matrix(D_n.to_sparse_matrix.toarray()))
        same_op(f"{rnd} from_matrix", M_r.cleanup(zero_threshold=1e-12), M_n.cleanup(zero_threshold=1e-12), ordered=False)
    ham_dir = "/root/reference/tests/hamiltonian_data"
    for fname in ["H2_STO-3G_SINGLET_JW.json", "H3+_STO-3G_SINGLET_JW.json", "H4_STO-3G_SINGLET_JW.json",
                  "HeH+_3-21G_SINGLET_JW.json", "LiH_STO-3G_SINGLET_JW.json", "H2_3-21G_SINGLET_JW.json"]:
        path = os.path.join(ham_dir, fname)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            dd = json.load(f)
        ham = {kk: complex(v[0], v[1]) for kk, v in dd["hamiltonian"].items()}
        hf = np.asarray(dd["data"]["hf_array"], dtype=int)
        H_r, H_n = ref.PauliwordOp.from_dictionary(ham), new.PauliwordOp.from_dictionary(ham)
        for sqp in ['Z', 'X']:
            T_r, T_n = RefQT(H_r, target_sqp=sqp), NewQT(H_n, target_sqp=sqp)
            same_op(f"{fname} generators {sqp}", T_r.symmetry_generators, T_n.symmetry_generators)
            O_r, O_n = T_r.taper_it(ref_state=hf), T_n.taper_it(ref_state=hf)
            same_op(f"{fname} tapered {sqp}", O_r, O_n, ordered=False, tol=1e-10)
            same(f"{fname} sector {sqp}", T_r.stabilizers.coeff_vec, T_n.stabilizers.coeff_vec)
            if sqp == 'Z':
                ps_r, ps_n = T_r.project_state(ref.QuantumState(hf)), T_n.project_state(new.QuantumState(hf))
                same(f"{fname} projected state", {kk: np.round(vv, 10) for kk, vv in ps_r.to_dictionary.items()},
                     {kk: np.round(vv, 10) for kk, vv in ps_n.to_dictionary.items()})
        same(f"{fname} hf energy", H_r.expval(ref.QuantumState(hf)), H_n.expval(new.QuantumState(hf)))
        same(f"{fname} noncontextual", bool(H_r.is_noncontextual), bool(H_n.is_noncontextual))
    print(f"{checks[0]} comparisons, {len(failures)} differences")
    for f in failures[:40]:
        print("  DIFF:", f)
    return 1 if failures else 0


if __name__ == "__main__":
    sys.exit(main())
