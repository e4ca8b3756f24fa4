"""Helpers of symmer/utils.py that sit directly on the hot path, with the reference's names: ground-state energies
(`exact_gs_energy`, symmer/utils.py:14-76), entanglement entropy (:78-94), `tensor_list` / `product_list` (:160-184)
and `matrix_allclose` (:300-323).

`exact_gs_energy` also accepts a `PauliwordOp` instead of its CSR matrix: the Lanczos iteration then runs
matrix-free, every H|v> through the device kernel behind `PauliwordOp.apply_dense` (`sym_apply`), so the 2^n x
(number of distinct X rows) CSR matrix of the reference (0.93 TB for the 24-qubit HOOH Hamiltonian)
is never built.
"""
from functools import reduce
from typing import List, Tuple, Union

import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla
import torch

from . import ops
from .base import PauliwordOp, QuantumState


def _device_operator(op: PauliwordOp) -> spla.LinearOperator:
    """Hermitian LinearOperator whose matvec is the matrix-free device kernel (host vector in, host vector out)."""
    side = 1 << op.n_qubits
    dev = ops.device()

    def matvec(v):
        psi = torch.from_numpy(np.ascontiguousarray(v, dtype=complex).reshape(-1)).to(dev)
        return op.apply_dense(psi).cpu().numpy()

    return spla.LinearOperator((side, side), matvec=matvec, dtype=complex)


def exact_gs_energy(sparse_matrix, initial_guess=None, n_particles=None, number_operator=None,
                    n_eigs=6) -> Tuple[float, QuantumState]:
    """symmer/utils.py:14-76: smallest eigenvalue and its eigenvector as a QuantumState; with `n_particles` and a
    diagonal `number_operator` the first of the `n_eigs` lowest eigenvectors with that particle number."""
    if number_operator is None:
        n_eigs = 1
    if isinstance(sparse_matrix, PauliwordOp):
        operator = sparse_matrix
        side = 1 << operator.n_qubits
        matrix = operator.to_sparse_matrix if side <= 2 ** 5 else _device_operator(operator)
    else:
        matrix, side = sparse_matrix, sparse_matrix.shape[0]
    if side > 2 ** 5:
        eigvals, eigvecs = spla.eigsh(matrix, k=n_eigs, v0=initial_guess, which='SA', maxiter=1e7)
    else:
        eigvals, eigvecs = np.linalg.eigh(matrix.toarray())
    order = np.argsort(eigvals)
    eigvals, eigvecs = eigvals[order], eigvecs[:, order]
    if n_particles is None:
        return eigvals[0], QuantumState.from_array(eigvecs[:, 0].reshape([-1, 1]))
    assert (number_operator is not None), 'Must specify the number operator.'
    assert (~np.any(number_operator.X_block)), 'Number operator not diagonal'
    for evl, evc in zip(eigvals, eigvecs.T):
        psi = QuantumState.from_array(evc.reshape([-1, 1])).cleanup(zero_threshold=1e-5)
        weights = np.square(abs(psi.state_op.coeff_vec))
        signs = 1 - 2 * ((psi.state_matrix.astype(int) @ number_operator.Z_block.astype(int).T) % 2)   # [K, terms]
        expval_n_particle = np.sum(number_operator.coeff_vec * (weights @ signs))
        if np.round(expval_n_particle) == n_particles:
            return evl, QuantumState.from_array(evc.reshape([-1, 1]))
    raise RuntimeError('No eigenvector of the correct particle number was identified - try increasing n_eigs.')


def get_entanglement_entropy(psi: QuantumState, qubits: List[int]) -> float:
    """symmer/utils.py:78-94: von Neumann entropy of the reduced density matrix on `qubits`."""
    eigvals = np.linalg.eigvals(psi.get_rdm(qubits))
    eigvals = eigvals[eigvals > 0]
    return -np.sum(eigvals * np.log(eigvals)).real


def random_anitcomm_2n_1_PauliwordOp(n_qubits: int, complex_coeff: bool = False, apply_clifford: bool = True) -> PauliwordOp:
    """symmer/utils.py:96-158: a maximal (2n+1)-term pairwise anticommuting operator with normal coefficients — the
    Jordan-Wigner-like strings Z..Z Y_i and Z..Z X_i plus Z..Z — scrambled by 5n random Clifford rotations (device
    kernel `sym_rotate`). Draws from the global NumPy RNG in the reference's order."""
    q = np.arange(n_qubits)
    z_below = q[:, None] > q[None, :]                       # Z on every qubit before i
    here = np.eye(n_qubits, dtype=bool)
    y_rows = np.hstack([here, z_below | here])
    x_rows = np.hstack([here, z_below])
    z_row = np.hstack([np.zeros(n_qubits, dtype=bool), np.ones(n_qubits, dtype=bool)])
    symp = np.vstack([y_rows, x_rows, z_row])
    coeff_vec = np.random.randn(symp.shape[0]).astype(complex)
    if complex_coeff:
        coeff_vec += 1j * np.random.randn(2 * n_qubits + 1).astype(complex)
    P_anticomm = PauliwordOp(symp, coeff_vec)
    if apply_clifford:
        rotations = []
        for _ in range(n_qubits * 5):
            P_rand = PauliwordOp.random(n_qubits, n_terms=1)
            P_rand.coeff_vec = [1]
            rotations.append((P_rand, np.random.choice([np.pi / 2, -np.pi / 2])))
        P_anticomm = P_anticomm.perform_rotations(rotations)
    assert P_anticomm.n_terms == 2 * n_qubits + 1
    return P_anticomm


def gram_schmidt_from_quantum_state(state: Union[np.ndarray, list, QuantumState]) -> np.ndarray:
    """symmer/utils.py:186-298: a unitary whose first column is the given state (the remaining columns come from a
    Gram-Schmidt sweep over the identity); small dense host linear algebra, as in the reference."""
    if isinstance(state, QuantumState):
        n_qubits = state.n_qubits
        state = state.to_sparse_matrix.toarray().reshape([-1])
    else:
        state = np.asarray(state).reshape([-1])
        n_qubits = round(np.log2(state.shape[0]))
        state = np.hstack((state, np.zeros(2 ** n_qubits - state.shape[0], dtype=complex)))
    assert state.shape[0] == 2 ** n_qubits, 'state is not defined on power of two'
    assert np.isclose(np.linalg.norm(state), 1), 'state is not normalized'
    M = np.eye(2 ** n_qubits, dtype=complex)
    if np.isclose(state[0], 0):
        swap = np.argmax(state)
        M[:, [0, swap]] = M[:, [swap, 0]]
    M[:, 0] = state
    for a in range(M.shape[0]):
        for b in range(a):
            M[:, a] -= (M[:, b].conj() @ M[:, a]) * M[:, b]
        M[:, a] /= np.linalg.norm(M[:, a])
    return M


def tensor_list(factor_list: List[PauliwordOp]) -> PauliwordOp:
    """symmer/utils.py:160-171."""
    return reduce(lambda x, y: x.tensor(y), factor_list)


def product_list(product_list: List[PauliwordOp]) -> PauliwordOp:
    """symmer/utils.py:173-184."""
    return reduce(lambda x, y: x * y, product_list)


def matrix_allclose(A: Union[sps.csr_matrix, np.ndarray], B: Union[sps.csr_matrix, np.ndarray], tol: float = 1e-15) -> bool:
    """symmer/utils.py:300-323: largest absolute difference <= tol, for dense or sparse matrices."""
    if sps.issparse(A) and sps.issparse(B):
        return bool(abs(A - B).max() <= tol)
    A = A.toarray() if sps.issparse(A) else A
    B = B.toarray() if sps.issparse(B) else B
    return bool(np.allclose(A, B, atol=tol))
