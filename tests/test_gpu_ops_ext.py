"""Kernel-level parity of the entry points added for rows g1-g3 at the sizes the API cases do not
reach: every row-width template of the qubit-wise commutation kernel, multi-word qubit gathers, and the
Walsh-Hadamard decomposition beyond one shared-memory tile (global butterfly stages) — against NumPy restatements."""
import numpy as np
import pytest
import torch

from oracle import pauli_oracle as po

pytestmark = pytest.mark.gpu


def _wht_rows(table):
    """Natural-order Walsh-Hadamard transform of every row, scaled by 1/len (NumPy restatement)."""
    t = np.array(table, dtype=complex)
    side = t.shape[-1]
    h = 1
    while h < side:
        t = t.reshape(t.shape[0], -1, 2, h)
        t = np.stack([t[:, :, 0] + t[:, :, 1], t[:, :, 0] - t[:, :, 1]], axis=2)
        h *= 2
    return t.reshape(t.shape[0], side) / side


@pytest.mark.parametrize("n_qubits", [1, 63, 64, 65, 128, 200, 500, 1000, 1024, 1500])
def test_qwc_every_row_width(n_qubits):
    """W = 1, 2, 4, 8, 16 templates and the generic kernel (W > 16), ragged M and N, sparse rows so that both
    outcomes occur."""
    from symmer_b200 import ops
    rng = np.random.default_rng(n_qubits)
    M, N = 70, 301
    density = min(0.3, 1.5 / np.sqrt(n_qubits))
    a = rng.random((M, 2 * n_qubits)) < density
    b = rng.random((N, 2 * n_qubits)) < density
    b[:M // 2] = a[:M // 2]                                   # identical rows always commute qubit-wise
    want = po.qubitwise_commutes_termwise(a, b)
    got = ops.commute_qwc(ops.pack(torch.from_numpy(a), n_qubits), ops.pack(torch.from_numpy(b), n_qubits)).cpu().numpy()
    assert got.shape == (M, N) and np.array_equal(got, want)
    assert want.any() and not want.all()
    empty = ops.commute_qwc(ops.pack(torch.from_numpy(a[:0]), n_qubits), ops.pack(torch.from_numpy(b), n_qubits))
    assert tuple(empty.shape) == (0, N)


@pytest.mark.parametrize("n_in,n_out", [(5, 5), (64, 64), (70, 70), (200, 200), (130, 260), (300, 40), (1000, 1000)])
def test_gather_qubits_multiword(n_in, n_out):
    from symmer_b200 import ops
    rng = np.random.default_rng(n_in * 7 + n_out)
    M = 37
    symp = rng.random((M, 2 * n_in)) < 0.4
    if n_in == n_out:
        src = rng.permutation(n_in)
    else:
        src = rng.integers(-1, n_in, size=n_out)              # repeats and holes (-1 -> identity)
    got = ops.unpack(ops.gather_qubits(ops.pack(torch.from_numpy(symp), n_in), src, n_in), n_out).cpu().numpy()
    want = np.zeros((M, 2 * n_out), dtype=bool)
    live = np.flatnonzero(src >= 0)
    want[:, live] = symp[:, src[live]]
    want[:, n_out + live] = symp[:, n_in + src[live]]
    assert np.array_equal(got, want)
    # padding bits of the packed output stay zero (they are hashed and compared by the dedup kernels)
    packed = ops.gather_qubits(ops.pack(torch.from_numpy(symp), n_in), src, n_in).cpu().numpy().view(np.uint64)
    assert np.array_equal(packed, po.pack_bits(want))


@pytest.mark.parametrize("n_qubits,K", [(0, 2), (1, 3), (5, 4), (11, 3), (12, 2), (13, 3), (15, 2), (17, 1)])
def test_walsh_hadamard_diagonals(n_qubits, K):
    """In-place transform of K XOR-diagonals: one shared-memory tile (n <= 12) and tile + global stages (n > 12)."""
    from symmer_b200 import ops
    rng = np.random.default_rng(100 + n_qubits)
    side = 1 << n_qubits
    diag = rng.standard_normal((K, side)) + 1j * rng.standard_normal((K, side))
    want = _wht_rows(diag)
    got = ops.pauli_decompose_diagonals(torch.from_numpy(diag.copy()).to(ops.device()), n_qubits).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13 * np.abs(diag).max())


@pytest.mark.parametrize("n_qubits", [1, 3, 6, 9])
def test_dense_decomposition_inverts_to_sparse_matrix(n_qubits):
    """Dense form (the kernel gathers the diagonals itself) against NumPy, and the round trip through the operator."""
    from symmer_b200 import PauliwordOp, ops
    rng = np.random.default_rng(n_qubits)
    side = 1 << n_qubits
    m = rng.standard_normal((side, side)) + 1j * rng.standard_normal((side, side))
    r = np.arange(side)
    want = _wht_rows(np.stack([m[r, r ^ x] for x in range(side)]))
    got = ops.pauli_decompose_dense(torch.from_numpy(m).to(ops.device()), n_qubits).cpu().numpy()
    assert np.allclose(got, want, rtol=1e-12, atol=1e-13)
    if n_qubits <= 6:
        op = PauliwordOp.from_matrix(m)
        assert op.n_terms == 4 ** n_qubits
        assert np.allclose(op.to_sparse_matrix.toarray(), m, atol=1e-12)


def test_from_matrix_of_a_wide_sparse_operator():
    """14 qubits: sparse path, diagonals longer than one shared-memory tile; from_matrix(to_sparse_matrix(H)) == H."""
    from symmer_b200 import PauliwordOp
    np.random.seed(3)
    H = PauliwordOp.random(14, 40)
    back = PauliwordOp.from_matrix(H.to_sparse_matrix)
    clean = back.cleanup(zero_threshold=1e-12)
    ok, why = po.compare_term_sets(clean.symp_matrix, clean.coeff_vec, *po.cleanup(H.symp_matrix, H.coeff_vec))
    assert ok, why


def test_rows_from_masks_inverts_term_masks():
    from symmer_b200 import PauliwordOp, ops
    from symmer_b200.base import _unsorted_masks
    for n in [1, 7, 31, 32, 62]:
        np.random.seed(n)
        P = PauliwordOp.random(n, 50)
        xm, zm, cp = _unsorted_masks(P)
        xz, c = ops.rows_from_masks(xm, zm, cp, n)
        assert torch.equal(xz, P.device_rows)
        assert np.allclose(c.cpu().numpy(), P.coeff_vec, rtol=1e-15, atol=0)
