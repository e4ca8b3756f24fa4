"""Stand-in for qiskit (test infrastructure only)."""
from . import qasm3  # noqa: F401


class QuantumCircuit:
    def __init__(self, *a, **k):
        raise NotImplementedError("qiskit stand-in")
