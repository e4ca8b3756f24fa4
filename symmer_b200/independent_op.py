"""IndependentOp with the reference's API (symmer/operators/independent_op.py) on the B200 engine:
algebraically independent stabilizer sets and the symmetry-generator search (config C2)."""
import warnings

import numpy as np

from .base import PauliwordOp
from .utils import _cref_binary, check_independent


class IndependentOp(PauliwordOp):
    """independent_op.py:9-43."""

    def __init__(self, symp_matrix, coeff_vec=None, target_sqp='Z') -> None:
        symp_matrix = np.asarray(symp_matrix)
        if coeff_vec is None:
            coeff_vec = np.ones(symp_matrix.reshape(-1, symp_matrix.shape[-1]).shape[0], dtype=complex)
        super().__init__(symp_matrix, coeff_vec)
        if target_sqp not in ['X', 'Z', 'Y']:
            raise ValueError('Target single-qubit Pauli not recognised - must be X or Z')
        self.target_sqp = target_sqp
        self._check_stab()
        self.coeff_vec = self.coeff_vec.real.astype(int).astype(complex) if np.all(self.coeff_vec.imag == 0) \
            else self.coeff_vec
        self._check_independent()

    @classmethod
    def from_PauliwordOp(cls, PwordOp: PauliwordOp) -> "IndependentOp":
        return cls(PwordOp.symp_matrix, PwordOp.coeff_vec)

    def _check_stab(self) -> None:
        """independent_op.py:146-151."""
        if not set(self.coeff_vec).issubset({0, +1, -1}):
            raise ValueError(f'Stabilizer coefficients not +/-1: {self.coeff_vec}')

    def _check_independent(self) -> None:
        """independent_op.py:153-159."""
        if not check_independent(self):
            raise ValueError('The supplied stabilizers are not independent')

    @classmethod
    def symmetry_generators(cls, PwordOp: PauliwordOp, commuting_override: bool = False,
                            largest_clique: bool = False) -> "IndependentOp":
        """independent_op.py:90-144: column reduction of [[Z X],[I]]; the rows of I below the zero
        columns of the reduced top block span the symmetry group. The GF(2) reduction and the
        commutation check run on the device; the clique search (only when the generators do not
        mutually commute) is host graph logic as in the reference."""
        n = PwordOp.n_qubits
        to_reduce = np.vstack([np.hstack([PwordOp.Z_block, PwordOp.X_block]), np.eye(2 * n, dtype=bool)])
        cref_matrix = _cref_binary(to_reduce)
        S_symp = cref_matrix[PwordOp.n_terms:, np.all(~cref_matrix[:PwordOp.n_terms], axis=0)].T
        S = cls(S_symp, np.ones(S_symp.shape[0]))
        if S.n_terms == 0:
            warnings.warn('The input PauliwordOp has no Z2 symmetries.')
            return S
        adj = S.adjacency_matrix
        if np.all(adj) or commuting_override:
            return S
        # largest mutually commuting subset (independent_op.py:132-144)
        import networkx as nx
        graph = nx.from_numpy_array(adj & ~np.eye(S.n_terms, dtype=bool))
        if S.n_terms < 10 or largest_clique:
            keep = sorted(max(nx.find_cliques(graph), key=len))
        else:
            colouring = nx.greedy_color(nx.complement(graph), strategy='independent_set')
            groups = {}
            for node, col in colouring.items():
                groups.setdefault(col, []).append(node)
            keep = sorted(groups[0])
            warnings.warn('Greedy method may identify non-optimal commuting symmetry terms; might be able to taper again.')
        return cls(S.symp_matrix[keep], np.ones(len(keep), dtype=complex))
