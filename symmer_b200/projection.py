"""Stabilizer-subspace projection and qubit tapering with the reference's API
(symmer/projection/base.py: S3Projection; symmer/projection/qubit_tapering.py: QubitTapering) on the
B200 engine — SURVEY.md §8f-2: the step right after config C2/C3 in the real workflow.

The operator never leaves the GPU between the Clifford rotations (`sym_rotate`), the projection
(`sym_project`: commutation filter, eigenvalue signs, removal of the stabilized qubits) and the final
duplicate merge (`sym_cleanup`). The rotation search on the (at most n-row) stabilizer set is host
logic, as in the reference.
"""
import warnings
from typing import List, Union

import numpy as np

from . import ops
from .base import PauliwordOp, QuantumState
from .independent_op import IndependentOp


class S3Projection:
    """projection/base.py:7-124."""
    rotated_flag = False

    def __init__(self, stabilizers: IndependentOp) -> None:
        self.stabilizers = stabilizers

    def _perform_projection(self, operator: PauliwordOp) -> PauliwordOp:
        """projection/base.py:44-84 on the device."""
        assert (operator.n_qubits == self.stabilizers.n_qubits), \
            'The input operator does not have the same number of qubits as the stabilizers'
        assert (self.rotated_flag), 'The operator has not been rotated - intended for use with perform_projection method'
        self.rotated_flag = False
        stab = self.rotated_stabilizers
        rows, cols = np.nonzero(stab.symp_matrix)
        assert len(rows) == stab.n_terms and np.array_equal(rows, np.arange(stab.n_terms)), \
            'projection needs single-qubit X or Z stabilizers (ill-defined for Y, as in the reference)'
        eigs = stab.coeff_vec.real.astype(float)
        xz, c = ops.project(operator.device_rows, operator.device_coeffs, operator.n_qubits, cols, eigs,
                            self.free_qubit_indices)
        n_free = len(self.free_qubit_indices)
        if n_free == 0:   # everything stabilized: a scalar (projection/base.py:83-84)
            return PauliwordOp(np.array([], dtype=bool), [complex(c.sum().cpu().numpy())])
        return PauliwordOp._from_device(xz, c, n_free).cleanup()

    def _project_state(self, state: QuantumState) -> QuantumState:
        """projection/base.py:126-158: the state seen from the stabilizer subspace. Hadamards on the qubits whose
        stabilizer was rotated onto X, the projectors (S^2 + S)/2 = (I + S)/2 of the rotated single-qubit
        stabilizers, then the stabilizer rotations as exponentials exp(i pi/4 R); the stabilized qubit positions are
        dropped and duplicates summed. Every product runs on the device."""
        from functools import reduce
        from .evolution import Had, trotter
        rotated = self.stabilizers.rotate_onto_single_qubit_paulis()
        transformation_list = [Had(self.stabilizers.n_qubits, int(i)) for i in
                               np.where(np.sum(rotated.X_block & ~rotated.Z_block, axis=0))[0]]
        for i in range(rotated.n_terms):
            sq = PauliwordOp._from_device(rotated[i].device_rows, rotated[i].device_coeffs, rotated.n_qubits)
            transformation_list.append((sq * sq + sq) * .5)
        for rot in self.stabilizers.stabilizer_rotations:
            transformation_list.append(trotter(rot[0] * (np.pi / 4 * 1j)))
        transformation = reduce(lambda x, y: x * y, transformation_list)
        transformed_state = transformation * state
        stab = np.where(rotated.symp_matrix)[1] % self.stabilizers.n_qubits
        free = np.setdiff1d(np.arange(self.stabilizers.n_qubits), stab)
        return QuantumState(transformed_state.state_matrix[:, free],
                            transformed_state.state_op.coeff_vec).cleanup(zero_threshold=1e-12)

    def perform_projection(self, operator: PauliwordOp, ref_state: Union[List[int], np.ndarray] = None,
                           sector: Union[List[int], np.ndarray] = None) -> PauliwordOp:
        """projection/base.py:86-124."""
        if sector is None and ref_state is not None:
            self.stabilizers.update_sector(ref_state)
        elif sector is not None:
            self.stabilizers.coeff_vec = np.array(sector, dtype=int)
        self.rotated_stabilizers = self.stabilizers.rotate_onto_single_qubit_paulis()
        self.stab_qubit_indices = np.where(self.rotated_stabilizers.symp_matrix)[1] % operator.n_qubits
        self.free_qubit_indices = np.setdiff1d(np.arange(operator.n_qubits), self.stab_qubit_indices)
        rotations = getattr(self.stabilizers, 'stabilizer_rotations', [])
        # the stabilizer rotations are Clifford: pure relabels of the rows on the device, and the dedup the reference
        # runs after them (base.py:1185) is absorbed by the one that follows the projection
        op_rotated = operator
        for generator, angle in rotations:
            op_rotated, status = op_rotated._rotation_step(generator, angle)
            if status == 'dirty':
                op_rotated = op_rotated.cleanup()
        self.rotated_flag = True
        return self._perform_projection(operator=op_rotated)


class QubitTapering(S3Projection):
    """projection/qubit_tapering.py:9-111."""
    name = 'qubit_tapering'

    def __init__(self, operator: PauliwordOp, target_sqp: str = 'Z') -> None:
        self.operator = operator
        self.target_sqp = target_sqp
        self._symmetry_generators = None
        self.n_taper = self.symmetry_generators.n_terms
        super().__init__(self.symmetry_generators)

    @property
    def symmetry_generators(self) -> IndependentOp:
        """qubit_tapering.py:42-52 (cached)."""
        if self._symmetry_generators is None:
            stabilizers = IndependentOp.symmetry_generators(self.operator)
            stabilizers.target_sqp = self.target_sqp
            self._symmetry_generators = stabilizers
        return self._symmetry_generators

    @symmetry_generators.setter
    def symmetry_generators(self, stabilizers: IndependentOp) -> None:
        """The reference's cached_property is assignable (tests/test_projection/test_qubit_tapering.py:72): a caller
        may taper with a subset of the symmetry generators; `taper_it` then re-initialises the parent projection."""
        self._symmetry_generators = stabilizers

    def project_state(self, state_to_project: QuantumState) -> QuantumState:
        """qubit_tapering.py:108-111."""
        return self._project_state(state_to_project)

    def taper_it(self, ref_state: Union[List[int], np.ndarray, QuantumState] = None,
                 sector: Union[List[int], np.ndarray] = None, aux_operator: PauliwordOp = None) -> PauliwordOp:
        """qubit_tapering.py:54-106."""
        if ref_state is not None:
            if not isinstance(ref_state, QuantumState):
                ref_state = QuantumState(ref_state)
            assert ref_state._is_normalized(), 'Reference state is not normalized.'
        if not (self.symmetry_generators is self.stabilizers or self.symmetry_generators == self.stabilizers):
            warnings.warn('the defined symmetry generators have been updated from parent class stabilizers')
            super().__init__(self.symmetry_generators)
        operator_to_taper = aux_operator if aux_operator is not None else self.operator
        return self.perform_projection(operator=operator_to_taper, ref_state=ref_state, sector=sector)
