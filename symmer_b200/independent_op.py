"""IndependentOp with the reference's API (symmer/operators/independent_op.py) on the B200 engine:
algebraically independent stabilizer sets and the symmetry-generator search (config C2)."""
import warnings
from typing import List, Tuple, Union

import numpy as np

from .base import PauliwordOp, QuantumState, single_term_expval
from .utils import _cref_binary, check_independent, symplectic_to_string


class IndependentOp(PauliwordOp):
    """independent_op.py:9-43."""

    def __init__(self, symp_matrix, coeff_vec=None, target_sqp='Z') -> None:
        symp_matrix = np.asarray(symp_matrix)
        if coeff_vec is None:
            coeff_vec = np.ones(symp_matrix.reshape(-1, symp_matrix.shape[-1]).shape[0], dtype=complex)
        if target_sqp not in ['X', 'Z', 'Y']:
            raise ValueError('Target single-qubit Pauli not recognised - must be X or Z')
        # the coefficient checks of independent_op.py:33-36, 146-151 run on the host array BEFORE it is uploaded
        coeff = np.asarray(coeff_vec, dtype=complex)
        if coeff.ndim == 1:
            if not set(coeff).issubset({0, +1, -1}):
                raise ValueError(f'Stabilizer coefficients not +/-1: {coeff}')
            if np.all(coeff.imag == 0):
                coeff = coeff.real.astype(int).astype(complex)
        super().__init__(symp_matrix, coeff)
        self.target_sqp = target_sqp
        self.stabilizer_rotations = None          # independent_op.py:41-42: filled by generate_stabilizer_rotations
        self.used_indices = None
        self.coeff_vec = coeff                    # host view authoritative from the start (callers set sectors in place)
        self._check_independent()

    @classmethod
    def _trusted(cls, xz, c, n_qubits: int, target_sqp: str = 'Z', nonzero=None) -> "IndependentOp":
        """Wrap device rows known to satisfy the constructor's checks — a subset or a Clifford rotation of a valid
        set is independent with coefficients in {0, +1, -1} — without the host round trip and the GF(2) reduction
        the validating constructor costs. `nonzero`: the coefficients are known to be all non-zero."""
        self = PauliwordOp._from_device(xz, c, n_qubits)
        self.__class__ = cls
        self.target_sqp = target_sqp
        self.stabilizer_rotations = None
        self.used_indices = None
        self._nz = nonzero
        return self

    def _all_nonzero(self) -> bool:
        """No zero-valued sector: Clifford rotations then never need the reference's per-step cleanup."""
        if self._c_host is not None:
            return bool(np.all(np.abs(self._c_host) > 1e-15))
        if getattr(self, '_nz', None) is None:
            self._nz = bool((self._c.abs() > 1e-15).all().item())
        return self._nz

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
    def _from_independent_rows(cls, xz, n_qubits: int) -> "IndependentOp":
        """Device rows that are independent by construction (distinct pivots of a GF(2) reduction), unit coefficients."""
        import torch
        ones = torch.ones(xz.shape[0], dtype=torch.complex128, device=xz.device)
        return cls._trusted(xz, ones, n_qubits, nonzero=True)

    @classmethod
    def symmetry_generators(cls, PwordOp: PauliwordOp, commuting_override: bool = False,
                            largest_clique: bool = False) -> "IndependentOp":
        """independent_op.py:90-144: column reduction of [[Z X],[I]]; the rows of I below the zero columns of the
        reduced top block span the symmetry group.

        Device-resident (config C2): the packed rows are bit-transposed (column reduction = row reduction of the
        transpose, utils.py:337-347), the identity block is appended directly in the packed-row layout (row of
        Z-bit q carries the packed row of X_q and vice versa, so the reduced identity part IS the generator in
        packed form), one bit-exact GF(2) reduction and one commutation kernel run back to back, and a single
        small copy (pivots + commutation matrix) comes back to the host. The clique search (only when the
        generators do not mutually commute) is host graph logic as in the reference."""
        import torch
        from . import ops
        n, M = PwordOp.n_qubits, PwordOp.n_terms
        if n == 0 or M == 0:
            return cls._symmetry_generators_host(PwordOp, commuting_override, largest_clique)
        W = ops.words_for(n)
        T = ops.bit_transpose(PwordOp.device_rows)                       # [128 W bit positions][ceil(M/64)]
        Cw_terms = T.shape[1]
        reduced = torch.cat([torch.cat([T[64 * W:64 * W + n], T[:n]], dim=0), _single_qubit_rows(n, T.device)],
                            dim=1).contiguous()
        piv = ops.rref_packed(reduced, 64 * Cw_terms + 128 * W)
        rows = reduced[:, Cw_terms:].contiguous()
        adj = ops.commute(rows, rows)
        back = torch.cat([piv.view(torch.uint8), adj.reshape(-1).view(torch.uint8)]).cpu().numpy()
        piv_host = back[:8 * n].view(np.int32)
        adj_host = back[8 * n:].view(bool).reshape(2 * n, 2 * n)
        sel = np.flatnonzero(piv_host >= 64 * Cw_terms)                  # top block reduced to zero
        if len(sel) == 0:
            warnings.warn('The input PauliwordOp has no Z2 symmetries.')
            return cls(np.zeros((0, 2 * n), dtype=bool), np.ones(0))
        adj_sel = adj_host[np.ix_(sel, sel)]
        if not (np.all(adj_sel) or commuting_override):
            sel = sel[_largest_commuting_subset(adj_sel, largest_clique)]
            adj_sel = adj_host[np.ix_(sel, sel)]
        S = cls._from_independent_rows(rows.index_select(0, torch.as_tensor(sel, dtype=torch.int64, device=rows.device)), n)
        S._cache['adjacency_matrix'] = adj_sel
        return S

    @classmethod
    def _symmetry_generators_host(cls, PwordOp: PauliwordOp, commuting_override: bool = False,
                                  largest_clique: bool = False) -> "IndependentOp":
        """The same search through the array seams (`_cref_binary` on a host matrix, device reduction inside):
        degenerate operands (no qubits / no terms), and the cross-check of the device-resident path in the tests."""
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
        keep = _largest_commuting_subset(adj, largest_clique)
        return cls(S.symp_matrix[keep], np.ones(len(keep), dtype=complex))

    # ------------------------------------------------------------------ container behaviour
    def __str__(self) -> str:
        return ' \n'.join(f'{c} {symplectic_to_string(row)}' for row, c in zip(self.symp_matrix, self.coeff_vec))

    def __repr__(self) -> str:
        return str(self)

    def __add__(self, Pword: "IndependentOp") -> "IndependentOp":
        return self.from_PauliwordOp(PauliwordOp.__add__(self, Pword))

    def __getitem__(self, key) -> "IndependentOp":
        """independent_op.py:316-350: indexing re-validates the selection as an IndependentOp."""
        sub = PauliwordOp.__getitem__(self, key)
        if sub.n_terms <= self.n_terms and not _has_repeats(key, self.n_terms):
            return IndependentOp._trusted(sub.device_rows, sub.device_coeffs, self.n_qubits, self.target_sqp)
        return IndependentOp(sub.symp_matrix, sub.coeff_vec)      # repeated indices: let the constructor reject them

    def __iter__(self):
        return iter([self[i] for i in range(self.n_terms)])

    # ------------------------------------------------------------------ Clifford rotations onto single-qubit Paulis
    def _rotate_by_single_Pword(self, Pword: PauliwordOp, angle: float = None) -> "IndependentOp":
        """independent_op.py:186-190. The rotation search below picks pivots by row position, so the row order of
        the reference's Clifford branch (anticommuting rows first, then the commuting ones; base.py:1151-1154) is
        reproduced. With no zero-valued sector the whole step stays on the device (rotation kernel, commutation
        kernel, stable partition) and the result is wrapped without re-validation: a Clifford rotation maps an
        independent +/-1 set onto one."""
        import torch
        multiple = (np.pi / 2 if angle is None else complex(angle).real) * 2 / np.pi
        clifford = abs(round(multiple) - multiple) <= 1e-18
        if clifford and self._all_nonzero():
            rotated, _ = PauliwordOp._rotation_step(self, Pword, angle)
            commutes = self.commutes_termwise_device(Pword)[:, 0]
            order = torch.argsort(commutes.to(torch.int8), stable=True)          # anticommuting rows first
            return IndependentOp._trusted(rotated.device_rows.index_select(0, order),
                                          rotated.device_coeffs.index_select(0, order), self.n_qubits, self.target_sqp,
                                          nonzero=True)
        rotated = PauliwordOp._rotate_by_single_Pword(self, Pword, angle)
        if rotated.n_terms == self.n_terms and clifford:
            commutes = self.commutes_termwise(Pword)[:, 0]
            if not commutes.all():
                anti = ~commutes
                if round(multiple) % 2:
                    # odd multiples form the anticommuting part with `*`, whose cleanup drops zero coefficients
                    anti &= abs(self.coeff_vec) > 1e-15
                rotated = rotated._take(np.concatenate([np.flatnonzero(anti), np.flatnonzero(commutes)]))
        return self.from_PauliwordOp(rotated)

    def perform_rotations(self, rotations: List[Tuple[PauliwordOp, float]]) -> "IndependentOp":
        """independent_op.py:192-204 (a dedup after every rotation, like base.py:1184-1185; a Clifford rotation of
        a set without zero-valued sectors cannot create duplicates or zeros, so that dedup is the identity)."""
        op = self
        for generator, angle in rotations:
            step = op._rotate_by_single_Pword(generator, angle)
            if getattr(step, '_nz', None) and step._c_host is None:
                op = step
            else:
                op = self.from_PauliwordOp(step.cleanup())
        return self.from_PauliwordOp(op) if op is self else op

    def _recursive_rotations(self, basis: "IndependentOp") -> None:
        """independent_op.py:204-241. Peel off the rows that already are single-qubit Paulis, then
        rotate the lightest remaining row onto a single-qubit Pauli on its least-supported unused
        qubit, and repeat on the rotated remainder."""
        n = self.n_qubits
        symp, coeff = basis.symp_matrix, basis.coeff_vec
        weight = symp.sum(axis=1)
        single = weight == 1
        # the reference finds the single-qubit rows as `basis - rest`, whose cleanup drops zero coefficients
        done = np.flatnonzero(single & (abs(coeff) > 1e-15))
        qubits = np.array([np.flatnonzero(symp[i])[0] % n for i in done], dtype=int)
        self.used_indices += np.append(qubits, qubits + n).tolist()
        rest_symp, rest_coeff = symp[~single], coeff[~single]
        if rest_symp.shape[0] == 0:
            return None
        rest = IndependentOp._trusted(*_subset(basis, np.flatnonzero(~single)), n, self.target_sqp)
        pivot_row = rest_symp[np.argsort(rest_symp.sum(axis=1))][0]
        candidates = np.setdiff1d(np.flatnonzero(pivot_row), np.array(self.used_indices))
        support = pivot_row * rest_symp.sum(axis=0)
        pivot_point = candidates[np.argmin(support[candidates])]
        # rotate onto the OTHER Pauli type of that qubit: X column q -> Z_q, Z column n+q -> X_q
        target = np.zeros(2 * n, dtype=bool)
        target[pivot_point + n if pivot_point < n else pivot_point - n] = True
        rotation = PauliwordOp(target ^ pivot_row, [1])
        self.stabilizer_rotations.append((rotation, None))
        return self._recursive_rotations(rest._rotate_by_single_Pword(rotation))

    def generate_stabilizer_rotations(self) -> None:
        """independent_op.py:243-273: the pi/2 rotations mapping this set onto single-qubit Paulis of
        type target_sqp."""
        assert (self.n_terms <= self.n_qubits), 'Too many terms in basis to reduce to single-qubit Paulis'
        assert (np.all(self.adjacency_matrix)), 'The basis is not commuting, hence the rotation is not possible'
        self.stabilizer_rotations = []
        self.used_indices = []
        basis = IndependentOp._trusted(self.device_rows, self.device_coeffs, self.n_qubits)
        self._recursive_rotations(basis)
        rotated = basis.perform_rotations(self.stabilizer_rotations)
        n = self.n_qubits
        for row in rotated.symp_matrix:
            q = np.flatnonzero(row)[0] % n
            target = np.zeros(2 * n, dtype=bool)
            target[q] = self.target_sqp in ['X', 'Y']
            target[q + n] = self.target_sqp in ['Y', 'Z']
            fix = target ^ row
            if fix.any():                                   # already the target Pauli otherwise
                self.stabilizer_rotations.append((PauliwordOp(fix, [1]), None))

    def rotate_onto_single_qubit_paulis(self) -> "IndependentOp":
        """independent_op.py:299-314: the stabilizers after the rotations, in their original order
        (the device Clifford kernel rewrites rows in place, so one pass keeps the order that the
        reference obtains by rotating the stabilizers one at a time)."""
        self.generate_stabilizer_rotations()
        if not self.stabilizer_rotations:
            return self
        op = PauliwordOp._from_device(self.device_rows, self.device_coeffs, self.n_qubits)
        for generator, angle in self.stabilizer_rotations:
            op, _ = op._rotation_step(generator, angle)
        if self._all_nonzero():
            return IndependentOp._trusted(op.device_rows, op.device_coeffs, self.n_qubits, self.target_sqp, nonzero=True)
        keep = np.flatnonzero(abs(op.coeff_vec) > 1e-15)    # the per-stabilizer cleanup of the reference
        return IndependentOp(op.symp_matrix[keep], op.coeff_vec[keep])

    # ------------------------------------------------------------------ sector assignment
    def update_sector(self, ref_state: Union[List[int], np.ndarray, QuantumState], threshold: float = 0.5) -> None:
        """independent_op.py:275-297: measure every stabilizer on the reference state; +/-1 when the
        expectation value is decisive, 0 otherwise (with a warning)."""
        if not isinstance(ref_state, QuantumState):
            ref_state = QuantumState(ref_state)
        assert ref_state._is_normalized(), 'Reference state is not normalized.'
        self.coeff_vec = np.array(assign_value(self, ref_state), dtype=complex)
        if np.any(self.coeff_vec == 0):
            zero = [symplectic_to_string(r) for r in self.symp_matrix[self.coeff_vec == 0]]
            warnings.warn(f'The stabilizers {zero} were assigned zero values - bad reference state.')


def assign_value(S: PauliwordOp, ref_state: QuantumState, threshold: float = 0.5) -> List[int]:
    """independent_op.py:364-383, evaluated term by term on the device (never through process.parallelize: no fork
    after CUDA initialisation). With a dense state the per-generator kernels are queued back to back and their
    results come back in one copy."""
    import torch
    from .base import dense_state_pays
    if S.n_terms and dense_state_pays(S.n_qubits, ref_state.n_terms):
        dense = ref_state.to_dense_device()
        ones = torch.ones(1, dtype=torch.complex128, device=dense.device)
        parts = []
        for i in range(S.n_terms):
            unit = PauliwordOp._from_device(S.device_rows[i:i + 1], ones, S.n_qubits)
            parts.append(unit.expval_dense(dense))
        expvals = torch.stack(parts).cpu().numpy().real
    else:
        expvals = [single_term_expval(PauliwordOp.__getitem__(S, i), ref_state) for i in range(S.n_terms)]
    return [int(np.sign(e)) if abs(e) > threshold else 0 for e in expvals]


def _subset(op: PauliwordOp, index: np.ndarray):
    """(rows, coefficients) of the selected terms, gathered on the device."""
    import torch
    idx = torch.as_tensor(np.asarray(index, dtype=np.int64), device=op.device_rows.device)
    return op.device_rows.index_select(0, idx), op.device_coeffs.index_select(0, idx)


def _has_repeats(key, n_terms: int) -> bool:
    if isinstance(key, (int, np.integer, slice)):
        return False
    idx = np.asarray(key)
    if idx.dtype == bool:
        return False
    idx = np.where(idx < 0, idx + n_terms, idx)
    return len(np.unique(idx)) != idx.size


_SINGLE_QUBIT_ROWS = {}


def _single_qubit_rows(n_qubits: int, device):
    """Packed rows of X_0..X_{n-1}, Z_0..Z_{n-1} (int64[2n, 2W]), cached per (n, device)."""
    import torch
    key = (n_qubits, str(device))
    if key not in _SINGLE_QUBIT_ROWS:
        W = max(1, (n_qubits + 63) // 64)
        rows = np.zeros((2 * n_qubits, 2 * W), dtype=np.uint64)
        q = np.arange(n_qubits)
        rows[q, q // 64] = np.uint64(1) << (q % 64).astype(np.uint64)
        rows[n_qubits + q, W + q // 64] = np.uint64(1) << (q % 64).astype(np.uint64)
        _SINGLE_QUBIT_ROWS[key] = torch.from_numpy(rows.view(np.int64)).to(device)
    return _SINGLE_QUBIT_ROWS[key]


def _largest_commuting_subset(adj: np.ndarray, largest_clique: bool) -> List[int]:
    """independent_op.py:132-144: indices of a largest mutually commuting subset (exact clique search below ten
    generators or on request, greedy colouring of the complement graph otherwise)."""
    import networkx as nx
    k = adj.shape[0]
    graph = nx.from_numpy_array(adj & ~np.eye(k, dtype=bool))
    if k < 10 or largest_clique:
        return sorted(max(nx.find_cliques(graph), key=len))
    colouring = nx.greedy_color(nx.complement(graph), strategy='independent_set')
    groups = {}
    for node, col in colouring.items():
        groups.setdefault(col, []).append(node)
    warnings.warn('Greedy method may identify non-optimal commuting symmetry terms; might be able to taper again.')
    return sorted(groups[0])
