"""Circuit simulation in the Heisenberg picture (symmer/evolution/circuit_symmerlator.py:8-203) on the B200 engine:
every gate is a short sequence of Pauli rotations, the circuit is applied to an observable with
`PauliwordOp.perform_rotations` (device kernels `sym_rotate` + `sym_cleanup`; Clifford gates are pure relabels of the
packed rows), and <0|U^dagger O U|0> is the sum of the coefficients of the rotated terms that are diagonal — one
masked reduction on the device.
"""
import ast
import operator as _op
import re
from typing import List

import numpy as np
import torch

from .base import PauliwordOp

# gate -> [(Pauli letters on the gate's qubits, multiple of pi/2)], circuit_symmerlator.py:54-136
_CLIFFORD_TABLE = {
    'x': [('X', 2)], 'y': [('Y', 2)], 'z': [('Z', 2)],
    'h': [('Z', 2), ('Y', 1)],
    's': [('Z', 1)], 'sdg': [('Z', 3)],
    'sx': [('X', 1)], 'sy': [('Y', 1)], 'sz': [('Z', 1)],
    'cx': [('ZX', 1), ('ZI', 3), ('IX', 3)],
    'cy': [('ZY', 1), ('ZI', 3), ('IY', 3)],
    'cz': [('ZZ', 1), ('ZI', 3), ('IZ', 3)],
}
_ROTATION_AXIS = {'rx': 'X', 'ry': 'Y', 'rz': 'Z'}


def _arithmetic(expr: str) -> float:
    """Value of an angle expression such as '3*pi/2' or '-0.25' (numbers, pi, + - * / ** and parentheses only)."""
    binary = {ast.Add: _op.add, ast.Sub: _op.sub, ast.Mult: _op.mul, ast.Div: _op.truediv, ast.Pow: _op.pow}
    unary = {ast.UAdd: _op.pos, ast.USub: _op.neg}

    def walk(node):
        if isinstance(node, ast.Expression):
            return walk(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)):
            return node.value
        if isinstance(node, ast.Name) and node.id == 'pi':
            return np.pi
        if isinstance(node, ast.BinOp) and type(node.op) in binary:
            return binary[type(node.op)](walk(node.left), walk(node.right))
        if isinstance(node, ast.UnaryOp) and type(node.op) in unary:
            return unary[type(node.op)](walk(node.operand))
        raise ValueError(f'unsupported angle expression: {expr}')

    return float(walk(ast.parse(expr.strip(), mode='eval')))


class CircuitSymmerlator:
    """circuit_symmerlator.py:8-203. Clifford gates are exact up to a global phase that cancels in expectation
    values; rotation gates are general-angle rotations (the operator can grow by a factor 1.5 per gate)."""

    def __init__(self, n_qubits: int) -> None:
        self.n_qubits = n_qubits
        self.sequence = []
        self.gate_map = {
            'x': self.X, 'y': self.Y, 'z': self.Z, 'rx': self.RX, 'ry': self.RY, 'rz': self.RZ,
            'sx': self.sqrtX, 'sy': self.sqrtY, 'sz': self.sqrtZ, 'cx': self.CX, 'cy': self.CY, 'cz': self.CZ,
            'h': self.H, 's': self.S, 'sdg': self.Sdag, '': self.R, 't': self.T, 'ccx': self.Toffoli, 'swap': self.SWAP,
        }

    def get_rotation_string(self, pauli: str, indices: List[int]) -> PauliwordOp:
        pauli = list(pauli)
        assert len(pauli) == len(indices), 'Number of Paulis and indices do not match'
        assert set(pauli).issubset({'I', 'X', 'Y', 'Z'}), 'Pauli operators are either I, X, Y or Z.'
        symp = np.zeros((1, 2 * self.n_qubits), dtype=bool)
        for i, P in zip(indices, pauli):
            symp[0, i] = P in 'XY'
            symp[0, self.n_qubits + i] = P in 'ZY'
        return PauliwordOp(symp, [1])

    def pi_2_multiple(self, multiple: int) -> float:
        return np.pi / 2 * multiple

    def _clifford(self, name: str, qubits: List[int]) -> None:
        for letters, multiple in _CLIFFORD_TABLE[name]:
            self.sequence.append((self.get_rotation_string(letters, qubits), self.pi_2_multiple(multiple)))

    # ---- Clifford gates
    def X(self, index: int) -> None: self._clifford('x', [index])
    def Y(self, index: int) -> None: self._clifford('y', [index])
    def Z(self, index: int) -> None: self._clifford('z', [index])
    def H(self, index: int) -> None: self._clifford('h', [index])
    def S(self, index: int) -> None: self._clifford('s', [index])
    def Sdag(self, index: int) -> None: self._clifford('sdg', [index])
    def sqrtX(self, index: int) -> None: self._clifford('sx', [index])
    def sqrtY(self, index: int) -> None: self._clifford('sy', [index])
    def sqrtZ(self, index: int) -> None: self._clifford('sz', [index])
    def CX(self, control: int, target: int) -> None: self._clifford('cx', [control, target])
    def CY(self, control: int, target: int) -> None: self._clifford('cy', [control, target])
    def CZ(self, control: int, target: int) -> None: self._clifford('cz', [control, target])

    def SWAP(self, qubit_1: int, qubit_2: int) -> None:
        self.CX(qubit_1, qubit_2)
        self.CX(qubit_2, qubit_1)
        self.CX(qubit_1, qubit_2)

    # ---- non-Clifford gates
    def R(self, pauli: str, indices: List[int], angle: float) -> None:
        self.sequence.append((self.get_rotation_string(pauli, indices), -angle))

    def RX(self, index: int, angle: float) -> None: self.R('X', [index], angle)
    def RY(self, index: int, angle: float) -> None: self.R('Y', [index], angle)
    def RZ(self, index: int, angle: float) -> None: self.R('Z', [index], angle)

    def T(self, index: int, angle: float) -> None:
        raise NotImplementedError()

    def Toffoli(self, control_1: int, control_2: int, target: int) -> None:
        raise NotImplementedError()

    # ---- execution
    def apply_sequence(self, operator: PauliwordOp) -> PauliwordOp:
        assert operator.n_qubits == self.n_qubits, 'The operator is defined over a different number of qubits'
        return operator.perform_rotations(self.sequence[::-1])

    def evaluate(self, operator: PauliwordOp) -> complex:
        """<0|U^dagger O U|0>: the coefficients of the rotated terms made of I and Z only, summed on the device."""
        rotated = self.apply_sequence(operator).cleanup()
        rows, coeffs = rotated.device_rows, rotated.device_coeffs
        W = rows.shape[1] // 2
        diagonal = (rows[:, :W] == 0).all(dim=1)
        return complex(torch.sum(torch.where(diagonal, coeffs, torch.zeros_like(coeffs))).cpu().numpy())

    @classmethod
    def from_qasm(cls, qasm: str, angle_factor: int = 1) -> "CircuitSymmerlator":
        """circuit_symmerlator.py:169-199: OpenQASM text with one instruction per line (version, include and register
        lines first); angles are arithmetic in `pi`."""
        instructions = qasm.split(';\n')[:-1]
        registers = instructions[2]
        self = cls(int(re.findall(r'\d+', registers)[0]))
        for step in instructions[3:]:
            head, *rest = step.split(' ')
            qubits = [int(q[2:-1]) for q in ''.join(rest).split(',')]
            if '(' in head:
                gate, angle = head.split('(')
                self.gate_map[gate](*qubits, angle=angle_factor * _arithmetic(angle[:-1]))
            else:
                self.gate_map[head](*qubits)
        return self

    @classmethod
    def from_qiskit(cls, circuit) -> "CircuitSymmerlator":
        """circuit_symmerlator.py:201-203 (needs qiskit, an optional dependency)."""
        try:
            from qiskit import qasm3
        except ImportError as err:
            raise ImportError('CircuitSymmerlator.from_qiskit needs the qiskit package') from err
        return cls.from_qasm(qasm3.dumps(circuit.reverse_bits()), angle_factor=-1)
