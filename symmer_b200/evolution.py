"""Operator exponentials and the gate library as PauliwordOps (symmer/evolution/exponentiation.py:6-38,
symmer/evolution/gate_library.py:9-198) on the B200 engine: thin host constructors, every product and sum
runs through the device kernels behind `PauliwordOp`. Needed by `S3Projection._project_state`
(symmer/projection/base.py:126-158), which maps a state into the stabilizer subspace.
"""
from functools import reduce

import numpy as np

from .base import PauliwordOp


def _single(n_qubits: int, index: int, pauli: str) -> PauliwordOp:
    """One Pauli on one qubit, identity elsewhere."""
    symp = np.zeros((1, 2 * n_qubits), dtype=bool)
    if pauli in 'XY':
        symp[0, index] = True
    if pauli in 'ZY':
        symp[0, n_qubits + index] = True
    return PauliwordOp(symp, [1])


def exponentiate_single_Pop(P: PauliwordOp) -> PauliwordOp:
    """exponentiation.py:6-24: e^{cP} = cosh(c) I + sinh(c) P for a single Pauli term with coefficient c
    (for e^{i theta P} pass c = i theta)."""
    assert (P.n_terms == 1), 'Can only exponentiate single Pauli terms'
    c = complex(P.coeff_vec[0])
    unit = PauliwordOp._from_device(P.device_rows, (P.device_coeffs * 0 + 1), P.n_qubits)
    return I(P.n_qubits).multiply_by_constant(np.cosh(c)) + unit.multiply_by_constant(np.sinh(c))


def trotter(op: PauliwordOp, trotnum: int = 1) -> PauliwordOp:
    """exponentiation.py:26-38: product of the single-term exponentials, repeated trotnum times with the
    coefficients divided by trotnum (exact when the terms commute)."""
    scaled = op.multiply_by_constant(1 / trotnum)
    factors = [exponentiate_single_Pop(scaled[i]) for i in range(scaled.n_terms)] * trotnum
    return reduce(lambda x, y: x * y, factors)


def truncated_exponential(op: PauliwordOp, truncate_at: int = 10) -> PauliwordOp:
    raise NotImplementedError                                                   # exponentiation.py:40-41


# ------------------------------------------------------------------ gate library (gate_library.py)
def I(n_qubits: int) -> PauliwordOp:
    return PauliwordOp(np.zeros((1, 2 * n_qubits), dtype=bool), [1])


def X(n_qubits: int, index: int) -> PauliwordOp:
    return _single(n_qubits, index, 'X')


def Y(n_qubits: int, index: int) -> PauliwordOp:
    return _single(n_qubits, index, 'Y')


def Z(n_qubits: int, index: int) -> PauliwordOp:
    return _single(n_qubits, index, 'Z')


def Had(n_qubits: int, index: int) -> PauliwordOp:
    """(Z + X)/sqrt(2), gate_library.py:63-77."""
    return (Z(n_qubits, index).multiply_by_constant(1 / np.sqrt(2))
            + X(n_qubits, index).multiply_by_constant(1 / np.sqrt(2)))


def CZ(n_qubits: int, control: int, target: int) -> PauliwordOp:
    """sqrt(i) exp(i pi/4 (Z_c Z_t - Z_t - Z_c)), gate_library.py:79-97."""
    ZI, IZ = Z(n_qubits, control), Z(n_qubits, target)
    exponent = (ZI * IZ - IZ - ZI).multiply_by_constant(np.pi / 4)
    return trotter(exponent.multiply_by_constant(1j), trotnum=1).multiply_by_constant(np.sqrt(1j))


def CX(n_qubits: int, control: int, target: int) -> PauliwordOp:
    h = Had(n_qubits, target)
    return h * CZ(n_qubits, control, target) * h


def CY(n_qubits: int, control: int, target: int) -> PauliwordOp:
    h, s = Had(n_qubits, target), S(n_qubits, target)
    return s * h * CZ(n_qubits, control, target) * h * s.dagger


def RX(n_qubits: int, index: int, angle: float) -> PauliwordOp:
    return trotter(X(n_qubits, index).multiply_by_constant(1j * angle / 2))


def RY(n_qubits: int, index: int, angle: float) -> PauliwordOp:
    return trotter(Y(n_qubits, index).multiply_by_constant(1j * angle / 2))


def RZ(n_qubits: int, index: int, angle: float) -> PauliwordOp:
    return trotter(Z(n_qubits, index).multiply_by_constant(1j * angle / 2))


def U1(n_qubits: int, index: int, angle: float) -> PauliwordOp:
    return RZ(n_qubits, index, angle).multiply_by_constant(np.exp(1j * angle / 2))


def S(n_qubits: int, index: int) -> PauliwordOp:
    return RZ(n_qubits, index, -np.pi / 2).multiply_by_constant(np.sqrt(1j))
