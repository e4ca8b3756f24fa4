"""One script of API calls, two executions: `run(api, sink)` is written purely against the reference's public API
(PauliwordOp / QuantumState / IndependentOp / QubitTapering), so the SAME code drives

  * the REAL reference in the build container (tests/golden/make_golden_replay.py, api = symmer through
    oracle/shim, sink = Recorder) — every result is written to tests/golden/replay_vectors.npz, and
  * this engine (api = symmer_b200, sink = Checker): on the B200 through the CUDA kernels
    (tests/test_gpu_replay.py) and on the CPU box through the NumPy test double (tests/test_host_logic.py),
    every result compared with what the reference returned for the same call.

Operators are compared as ordered (rows bit-exact in the reference's order, coefficients rtol 1e-12) wherever the
reference defines an order, and as term sets otherwise (SURVEY.md §8c caveat 1)."""
import numpy as np


def _canonical(symp, coeff):
    symp = np.asarray(symp, dtype=bool)
    coeff = np.asarray(coeff, dtype=complex)
    if symp.shape[0] == 0:
        return symp, coeff
    rev = np.ascontiguousarray(symp[:, ::-1]).view(np.uint8)
    order = np.argsort(rev.view(np.dtype((np.void, rev.shape[1]))).ravel(), kind="stable") if symp.shape[1] else np.arange(symp.shape[0])
    return symp[order], coeff[order]


class Recorder:
    """Record mode: stores what the reference returns."""
    replay = False

    def __init__(self):
        self.out = {}

    def _put(self, tag, **arrays):
        assert f"{tag}/kind" not in self.out, f"duplicate tag {tag}"
        for k, v in arrays.items():
            self.out[f"{tag}/{k}"] = np.asarray(v)

    def op(self, tag, op, ordered=True, tol=1e-12):
        symp, coeff = (op.symp_matrix, op.coeff_vec) if ordered else _canonical(op.symp_matrix, op.coeff_vec)
        self._put(tag, kind=["op"], symp=symp, coeff=coeff)

    def value(self, tag, value, tol=1e-12):
        if isinstance(value, str):
            self._put(tag, kind=["str"], value=[value])
        elif isinstance(value, dict):
            keys = sorted(value.keys())
            self._put(tag, kind=["dict"], keys=keys, vals=np.asarray([value[k] for k in keys], dtype=complex))
        else:
            self._put(tag, kind=["array"], value=np.asarray(value))

    def state(self, tag, state, tol=1e-10):
        self.value(tag, state.to_dictionary, tol)

    def given(self, tag, producer):
        arrays = producer()
        for k, v in arrays.items():
            self.out[f"given/{tag}/{k}"] = np.asarray(v)
        return arrays


class Checker:
    """Replay mode: compares with what the reference returned."""
    replay = True

    def __init__(self, stored):
        self.stored = stored
        self.checked = 0

    def _get(self, tag, field):
        return self.stored[f"{tag}/{field}"]

    def op(self, tag, op, ordered=True, tol=1e-12):
        want_s, want_c = self._get(tag, "symp"), self._get(tag, "coeff")
        got_s, got_c = (op.symp_matrix, op.coeff_vec) if ordered else _canonical(op.symp_matrix, op.coeff_vec)
        scale = max(1.0, float(np.abs(want_c).max())) if want_c.size else 1.0
        # SURVEY.md §8c caveat 3: terms at or below tau on either side are cancellation residues (the reference's
        # FMA-contracted complex products leave them, the engine's commutative multiply cancels exactly): ignored
        tau = 1e-12 * scale
        if want_c.size and got_c.size and (np.any(np.abs(want_c) <= tau) or np.any(np.abs(got_c) <= tau)) \
                and not (want_c.size == 1 and got_c.size == 1):
            want_s, want_c = want_s[np.abs(want_c) > tau], want_c[np.abs(want_c) > tau]
            got_s, got_c = got_s[np.abs(got_c) > tau], got_c[np.abs(got_c) > tau]
        assert got_s.shape == want_s.shape, (tag, got_s.shape, want_s.shape)
        assert np.array_equal(got_s, want_s), tag
        assert np.allclose(got_c, want_c, rtol=tol, atol=tol * scale), (tag, np.abs(got_c - want_c).max())
        self.checked += 1

    def value(self, tag, value, tol=1e-12):
        kind = str(self._get(tag, "kind")[0])
        if kind == "str":
            assert value == str(self._get(tag, "value")[0]), tag
        elif kind == "dict":
            want = {str(k): v for k, v in zip(self._get(tag, "keys"), self._get(tag, "vals")) if abs(v) > 1e-12 or len(value) <= 1}
            value = {k: v for k, v in value.items() if abs(v) > 1e-12 or len(value) <= 1}       # residues, as for operators
            keys = sorted(value.keys())
            assert keys == sorted(want.keys()), tag
            assert np.allclose(np.asarray([value[k] for k in keys], dtype=complex),
                               np.asarray([want[k] for k in keys], dtype=complex), rtol=tol, atol=tol), tag
        else:
            want = self._get(tag, "value")
            got = np.asarray(value)
            assert got.shape == want.shape, (tag, got.shape, want.shape)
            if want.dtype == bool:
                assert np.array_equal(got.astype(bool), want), tag
            else:
                assert np.allclose(got.astype(complex), want.astype(complex), rtol=tol, atol=tol), tag
        self.checked += 1

    def state(self, tag, state, tol=1e-10):
        self.value(tag, state.to_dictionary, tol)

    def given(self, tag, producer):
        prefix = f"given/{tag}/"
        return {k[len(prefix):]: self.stored[k] for k in self.stored if k.startswith(prefix)}


def run(api, sink, rounds=6, hamiltonians=None):
    """`api`: namespace with PauliwordOp, QuantumState, IndependentOp, QubitTapering. `hamiltonians`: record mode only,
    {tag: callable returning dict(symp=, coeff=, hf=)}; in replay mode the stored inputs are used."""
    P, Q = api.PauliwordOp, api.QuantumState
    for rnd in range(rounds):
        n = int(np.random.default_rng(rnd).integers(1, 9))
        m = int(np.random.default_rng(rnd + 100).integers(1, 14))
        np.random.seed(rnd)
        A = P.random(n, m)
        np.random.seed(rnd + 1000)
        B = P.random(n, max(1, m // 2), complex_coeffs=False)
        t = f"r{rnd}"
        sink.op(f"{t}/random", A)
        for by in ['magnitude', 'weight', 'support', 'X', 'Y', 'Z', 'lex']:
            for key in ['decreasing', 'increasing']:
                sink.op(f"{t}/sort_{by}_{key}", A.cleanup().sort(by=by, key=key), ordered=(by in ('magnitude', 'lex')))
        sink.op(f"{t}/cleanup", A.cleanup())
        sink.op(f"{t}/mul", A * B, ordered=False)
        sink.op(f"{t}/add", A + B, ordered=False)
        sink.op(f"{t}/sub", A - B, ordered=False)
        sink.op(f"{t}/pow2", A ** 2, ordered=False)
        sink.op(f"{t}/pow0", A ** 0)
        sink.op(f"{t}/commutator", A.commutator(B), ordered=False)
        sink.op(f"{t}/anticommutator", A.anticommutator(B), ordered=False)
        sink.value(f"{t}/commutes", [bool(A.commutes(B))])
        sink.op(f"{t}/dagger", A.dagger)
        sink.op(f"{t}/const", A.multiply_by_constant(0.3 - 2j))
        sink.op(f"{t}/scalar", A * 2.5)
        sink.op(f"{t}/append", A.append(B))
        sink.value(f"{t}/Y_count", A.Y_count)
        sink.value(f"{t}/commutes_termwise", A.commutes_termwise(B))
        sink.value(f"{t}/anticommutes_termwise", A.anticommutes_termwise(B))
        sink.value(f"{t}/qwc", A.qubitwise_commutes_termwise(B))
        sink.value(f"{t}/adjacency", A.adjacency_matrix)
        sink.value(f"{t}/noncontextual", [bool(A.is_noncontextual)])
        sink.value(f"{t}/str", str(A))
        sink.value(f"{t}/dict", A.to_dictionary)
        sink.value(f"{t}/eq", [bool(A == B), bool(A == A.cleanup())])
        gens = A.generators
        sink.op(f"{t}/generators", gens)
        sink.op(f"{t}/slice", A[1:])
        sink.op(f"{t}/list_index", A[[0, -1]])
        if n <= 5:
            sink.value(f"{t}/sparse", A.to_sparse_matrix.toarray())
        recon, mask = A.generator_reconstruction(gens)
        sink.value(f"{t}/recon", recon)
        sink.value(f"{t}/recon_mask", mask)
        Qr = P(B.symp_matrix[0], [1])
        for i, angle in enumerate([None, 0.37, np.pi, -np.pi / 2, 3 * np.pi / 2]):
            sink.op(f"{t}/rotate_{i}", A.perform_rotations([(Qr, angle)]), ordered=False)
        sink.op(f"{t}/tensor", A.tensor(B), ordered=False)
        np.random.seed(rnd + 7)
        psi = Q.random(n, 5)
        sink.state(f"{t}/state", psi)
        sink.value(f"{t}/expval", [A.expval(psi)])
        sink.value(f"{t}/braket", [psi.dagger * psi])
        sink.state(f"{t}/op_ket", A * psi)
        sink.value(f"{t}/state_str", str(psi.sort()))
        sink.value(f"{t}/dense", psi.to_dense_matrix)

        # duplicates and vanishing coefficients
        n2 = 3 + rnd % 4
        np.random.seed(500 + rnd)
        base_symp = np.random.rand(6, 2 * n2) < 0.4
        symp = np.vstack([base_symp, base_symp[[0, 2, 2]]])
        coeff = np.random.randn(9) + 1j * np.random.randn(9)
        coeff[1] = 0
        coeff[7] = -coeff[2]
        D = P(symp, coeff)
        sink.op(f"{t}/dup_cleanup", D.cleanup())
        sink.op(f"{t}/dup_cleanup_thr", D.cleanup(zero_threshold=0.5))
        sink.op(f"{t}/dup_square", D * D, ordered=False)
        sink.value(f"{t}/dup_dict", D.to_dictionary)
        sink.value(f"{t}/dup_adjacency", D.adjacency_matrix)
        sink.value(f"{t}/dup_qwc", D.adjacency_matrix_qwc)
        Dc = D.cleanup()
        for rel in ['C', 'AC', 'QWC']:
            for strategy in ['largest_first', 'sorted_insertion', 'DSATUR']:
                cover = Dc.clique_cover(rel, strategy)
                sink.value(f"{t}/cover_keys_{rel}_{strategy}", sorted(cover.keys()))
                for k in sorted(cover.keys()):
                    sink.op(f"{t}/cover_{rel}_{strategy}_{k}", cover[k], ordered=False)
            sink.op(f"{t}/largest_clique_{rel}", Dc.largest_clique(rel), ordered=False)
        M = P.from_matrix(D.to_sparse_matrix.toarray(), disable_loading_bar=True)
        sink.op(f"{t}/from_matrix", M.cleanup(zero_threshold=1e-12), ordered=False, tol=1e-10)

        # a commuting independent set (Clifford images of Z_0..Z_{k-1}) rotated back onto single-qubit Paulis
        k = 1 + rnd % n2
        z = np.zeros((k, 2 * n2), dtype=bool)
        z[np.arange(k), n2 + np.arange(k)] = True
        signs = np.random.choice([1, -1], size=k)
        S = P(z, signs)
        rots = []
        for _ in range(2 * n2):
            q = np.random.rand(2 * n2) < 0.5
            if not q.any():
                q[0] = True
            rots.append((P(q, [1]), None))
        S = S.perform_rotations(rots)
        sink.op(f"{t}/clifford_images", S, ordered=False)
        S_symp, S_coeff = S.symp_matrix.copy(), S.coeff_vec.copy()
        if sink.replay:      # both sides start the stabilizer workflow from the reference's row order
            S_symp, S_coeff = sink._get(f"{t}/stab_input", "symp"), sink._get(f"{t}/stab_input", "coeff")
        else:
            sink._put(f"{t}/stab_input", kind=["op"], symp=S_symp, coeff=S_coeff)
        for sqp in ['Z', 'X']:
            Ind = api.IndependentOp(S_symp, S_coeff, target_sqp=sqp)
            sink.op(f"{t}/onto_sqp_{sqp}", Ind.rotate_onto_single_qubit_paulis())
            sink.value(f"{t}/rotation_list_{sqp}",
                       np.array([p.symp_matrix[0] for p, _ in Ind.stabilizer_rotations], dtype=bool).reshape(-1, 2 * n2))
            sink.op(f"{t}/ind_item_{sqp}", Ind[0])
            sink.op(f"{t}/ind_rotations_{sqp}", Ind.perform_rotations(Ind.stabilizer_rotations))

        # state algebra
        np.random.seed(900 + rnd)
        p_state, q_state = Q.random(n2, 6), Q.random(n2, 4)
        sink.state(f"{t}/state_add", p_state + q_state)
        sink.state(f"{t}/state_sub", p_state - q_state)
        sink.state(f"{t}/state_scalar", p_state * 0.5j)
        sink.state(f"{t}/state_normalize", (p_state + q_state).normalize)
        sink.value(f"{t}/state_overlap", [q_state.dagger * p_state])
        sink.value(f"{t}/state_sort_support", str(p_state.sort(key='support')))
        sink.value(f"{t}/state_rdm", p_state.get_rdm([0]))
        sink.state(f"{t}/bra_op", (p_state.dagger * D).dagger)

    # molecular Hamiltonians: symmetry generators, sector, tapering for both target Paulis, state projection
    tags = sorted(hamiltonians) if hamiltonians is not None else sorted(
        {k.split("/")[1] for k in sink.stored if k.startswith("given/")})
    for tag in tags:
        g = sink.given(tag, hamiltonians[tag] if hamiltonians is not None else None)
        H = P(g["symp"], g["coeff"])
        hf = np.asarray(g["hf"], dtype=int)
        for sqp in ['Z', 'X']:
            T = api.QubitTapering(H, target_sqp=sqp)
            sink.op(f"{tag}/generators_{sqp}", T.symmetry_generators)
            sink.op(f"{tag}/tapered_{sqp}", T.taper_it(ref_state=hf), ordered=False, tol=1e-10)
            sink.value(f"{tag}/sector_{sqp}", T.stabilizers.coeff_vec)
            if sqp == 'Z':
                sink.state(f"{tag}/projected_state", T.project_state(Q(hf)))
        sink.value(f"{tag}/hf_energy", [H.expval(Q(hf))])
        sink.value(f"{tag}/noncontextual", [bool(H.is_noncontextual)])
