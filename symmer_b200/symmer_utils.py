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
