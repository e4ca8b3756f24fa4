"""Config-size parity (BASELINE.json configs C3 and C4) and the symmetric adjacency path, on the GPU.

C3: one general and one Clifford rotation of PauliwordOp.random(1000 q, 100 000 terms) against the CPU oracle
    (symmer/operators/base.py:1090-1186 restated) on the WHOLE operator.
C4: matrix-free <psi|H|psi> of HOOH STO-3G (24 q, 14 905 terms) on a dense 2^24 state against three independent
    evaluations: a product state (the expectation value factorises over qubits: O(terms * n) on the CPU), the sum
    of 8 basis shards against the unsharded value, and the Hartree-Fock basis state against the diagonal terms.
"""
import os

import numpy as np
import pytest
import torch

from oracle import pauli_oracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def ops():
    import symmer_b200.ops as o
    o.device()
    return o


def test_commute_self_upper_triangle_matches_full(ops):
    """adjacency_matrix of a large operator computes the upper block triangle and mirrors it (sym_mirror_upper)."""
    for n, M, blk in [(130, 9001, 2048), (1000, 8200, 4096), (64, 8192, 2048)]:
        s, _ = po.random_operator(n, M, seed=n + M)
        s[M // 2:] = s[: M - M // 2] ^ (np.arange(M - M // 2)[:, None] % 7 == 0)      # structure: near-duplicates
        xz = ops.pack(torch.from_numpy(s), n)
        full = ops.commute(xz, xz)
        assert torch.equal(full, full.T)
        sym = ops.commute_self(xz, block_rows=blk)
        assert sym.shape == full.shape and torch.equal(sym, full)
    # a slice the oracle can check
    ref = po.commutes_termwise(s[:300], s[:300])
    assert np.array_equal(sym[:300, :300].cpu().numpy(), ref)


@pytest.mark.timeout(600)
def test_config_c3_rotations_of_100k_terms_against_oracle():
    """BASELINE config C3 at full size: P = random(1000 q, 100 000 terms), one non-Clifford and one Clifford rotation,
    every output term against the oracle (rows bit-exact, coefficients rtol 1e-12)."""
    from symmer_b200 import PauliwordOp
    n, M = 1000, 100_000
    p_s, p_c = po.random_operator(n, M, seed=3)
    q_s, _ = po.random_operator(n, 1, seed=4)
    P = PauliwordOp(p_s, p_c)
    Q = PauliwordOp(q_s, [1])
    for angle in (0.37, np.pi / 2, None, 2.1):
        rot = P.perform_rotations([(Q, angle)])
        ref_s, ref_c = po.perform_rotations(p_s, p_c, [(q_s[0], angle)])
        assert rot.n_terms == len(ref_c)
        ok, why = po.compare_term_sets(rot.symp_matrix, rot.coeff_vec, ref_s, ref_c, scale=float(np.abs(p_c).max()))
        assert ok, (angle, why)
    # two rotations in sequence (the second acts on ~1.5e5 rows)
    q2_s, _ = po.random_operator(n, 1, seed=5)
    rot = P.perform_rotations([(Q, 0.37), (PauliwordOp(q2_s, [1]), 1.1)])
    ref_s, ref_c = po.perform_rotations(p_s, p_c, [(q_s[0], 0.37), (q2_s[0], 1.1)])
    ok, why = po.compare_term_sets(rot.symp_matrix, rot.coeff_vec, ref_s, ref_c, scale=float(np.abs(p_c).max()))
    assert ok, why


def _pauli_factor(a, b):
    """<phi|sigma|phi> of one qubit state a|0> + b|1> for sigma = I, X, Y, Z."""
    return np.array([abs(a) ** 2 + abs(b) ** 2, 2 * (np.conj(a) * b).real, 2 * (np.conj(a) * b).imag, abs(a) ** 2 - abs(b) ** 2])


@pytest.mark.timeout(600)
def test_config_c4_expval_24_qubits_independent_checks(ops):
    from symmer_b200 import PauliwordOp
    d = np.load(os.path.join(ROOT, "tests", "golden", "hamiltonians", "HOOH_STO3G.npz"))
    n = int(d["n_qubits"][0])
    symp = np.unpackbits(d["symp"], axis=1)[:, :2 * n].astype(bool)
    coeff = d["coeff"]
    assert n == 24 and symp.shape[0] == 14905
    H = PauliwordOp(symp, coeff)
    xm, zm, cp = H._terms_sorted()
    dev = ops.device()
    # (1) product state: the expectation value factorises over the qubits
    rng = np.random.default_rng(7)
    amps = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    amps /= np.linalg.norm(amps, axis=1, keepdims=True)
    psi = np.ones(1, dtype=complex)
    for q in range(n):                         # qubit 0 = most significant bit of the basis index (base.py:1502-1503)
        psi = np.kron(psi, amps[q])
    fac = np.array([_pauli_factor(amps[q, 0], amps[q, 1]) for q in range(n)])     # [n, 4]
    x, z = symp[:, :n], symp[:, n:]
    kind = np.where(x & z, 2, np.where(x, 1, np.where(z, 3, 0)))                   # I X Y Z -> 0 1 2 3
    ref = np.sum(coeff * np.prod(fac[np.arange(n)[None, :], kind], axis=1))
    psi_d = torch.from_numpy(psi).to(dev)
    e = complex(ops.expval_dense(xm, zm, cp, n, psi_d).cpu().numpy())
    assert np.isclose(e.real, ref.real, rtol=1e-11, atol=1e-12) and abs(e.imag) < 1e-10
    # (2) the sum over 8 basis shards equals the unsharded value
    side = 1 << n
    parts = [complex(ops.expval_dense(xm, zm, cp, n, psi_d, k * side // 8, (k + 1) * side // 8).cpu().numpy()) for k in range(8)]
    assert np.isclose(sum(parts), e, rtol=1e-12, atol=1e-13)
    # (3) a basis state: only the diagonal terms contribute
    hf = int("".join(str(int(b)) for b in d["hf_array"]), 2)                      # qubit 0 = most significant bit
    onehot = torch.zeros(side, dtype=torch.complex128, device=dev)
    onehot[hf] = 1.0
    diag = ~x.any(axis=1)
    bits = np.array([(hf >> (n - 1 - q)) & 1 for q in range(n)], dtype=bool)
    e_ref = np.sum(coeff[diag] * (-1.0) ** np.count_nonzero(z[diag] & bits[None, :], axis=1))
    e_hf = complex(ops.expval_dense(xm, zm, cp, n, onehot).cpu().numpy())
    assert np.isclose(e_hf.real, e_ref.real, rtol=1e-12) and abs(e_hf.imag) < 1e-12
    assert np.isclose(e_hf.real, float(d["hf_energy"][0]), rtol=1e-8)         # the reference data's own Hartree-Fock energy


@pytest.mark.parametrize("n,m", [(1, 7), (5, 200), (64, 300), (65, 257), (130, 1000), (1000, 5000)])
def test_lex_order_on_device_matches_numpy_lexsort(ops, n, m):
    """sort('lex') (base.py:469-470, the canonical order behind every `==`) as a device radix sort of the packed rows."""
    from symmer_b200 import PauliwordOp
    s, c = po.random_operator(n, m, seed=n * 7 + m)
    s[m // 2:] = s[: m - m // 2]                    # ties: the sort must be stable like np.lexsort
    s[:, -3:] = False                                # all-zero high columns: skipped word bits
    order = np.lexsort(s.T)
    perm = ops.lex_order(ops.pack(torch.from_numpy(s), n)).cpu().numpy()
    assert np.array_equal(perm, order)
    P = PauliwordOp(s, c)
    for key in ("decreasing", "increasing"):
        srt = P.sort(by="lex", key=key)
        ref = order if key == "decreasing" else order[::-1]
        assert np.array_equal(srt.symp_matrix, s[ref]) and np.array_equal(srt.coeff_vec, c[ref])
    Q = PauliwordOp(s[::-1].copy(), c[::-1].copy())
    assert P == Q and not (P == PauliwordOp(s, c * 1.5))


def test_state_inner_product_join_is_exact(ops):
    """bra * ket (base.py:1781-1830) joins the two states on equal bit strings: against the dict join of the oracle
    semantics, with repeated basis states and at widths beyond one word."""
    from symmer_b200 import QuantumState
    rng = np.random.default_rng(3)
    for n, k1, k2 in [(10, 300, 200), (70, 500, 400), (130, 64, 64)]:
        pool = rng.integers(0, 2, size=(150, n))
        left, right = pool[rng.integers(0, 150, k1)], pool[rng.integers(0, 150, k2)]
        cl = rng.standard_normal(k1) + 1j * rng.standard_normal(k1)
        cr = rng.standard_normal(k2) + 1j * rng.standard_normal(k2)
        bra = QuantumState(left, cl, vec_type='bra')
        ket = QuantumState(right, cr, vec_type='ket')
        dl, dr = {}, {}
        for r, v in zip(left, cl):
            dl[r.tobytes()] = dl.get(r.tobytes(), 0) + v
        for r, v in zip(right, cr):
            dr[r.tobytes()] = dr.get(r.tobytes(), 0) + v
        ref = sum(v * dr[k] for k, v in dl.items() if k in dr)
        assert np.isclose(bra * ket, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("n_qubits,n_terms", [(1000, 20000), (128, 5000), (100, 3000), (40, 3000), (130, 2000)])
def test_padded_general_rotation_cleans_up_to_the_compact_one(n_qubits, n_terms):
    """sym_rotate mode 4 (one pass, 2M rows, zero-coefficient copies of the commuting rows) followed by the cleanup
    gives the same rows, in the same order, with the same sums as mode 0 followed by the cleanup."""
    from symmer_b200 import ops
    s, c = po.random_operator(n_qubits, n_terms, seed=n_terms)
    s[n_terms // 2:n_terms // 2 + 50] = s[:50]                      # some duplicates inside the operand
    q_s, _ = po.random_operator(n_qubits, 1, seed=99)
    xz = ops.pack(torch.from_numpy(s), n_qubits)
    cc = torch.from_numpy(c).cuda()
    q = ops.pack(torch.from_numpy(q_s), n_qubits)
    ca, sa = np.cos(0.37), np.sin(0.37)
    r0 = ops.rotate(xz, cc, q, ca, sa, 0)
    r4 = ops.rotate(xz, cc, q, ca, sa, 0, padded_ok=True)
    W = xz.shape[1] // 2
    if (W % 2 == 0 and W <= 16) or W == 1:
        assert r4[0].shape[0] == 2 * n_terms and r0[0].shape[0] < 2 * n_terms
    else:
        assert r4[0].shape[0] == r0[0].shape[0]                      # odd word counts keep the compact form
    a_xz, a_c = ops.cleanup(*r0)
    b_xz, b_c = ops.cleanup(*r4)
    assert torch.equal(a_xz, b_xz)
    assert torch.equal(a_c, b_c)
    # and both are the reference's rotation followed by its cleanup, as a set of terms
    o_s, o_c = po.perform_rotations(s, c, [(q_s[0], 0.37)])
    got_s = po.unpack_bits(b_xz.cpu().numpy().view(np.uint64), n_qubits)
    got_c = b_c.cpu().numpy()
    assert got_s.shape == o_s.shape
    order_o = np.lexsort(o_s.T[::-1])
    order_g = np.lexsort(got_s.T[::-1])
    assert np.array_equal(got_s[order_g], o_s[order_o])
    np.testing.assert_allclose(got_c[order_g], o_c[order_o], rtol=1e-12, atol=1e-15)
