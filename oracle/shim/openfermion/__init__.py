"""Stand-in for openfermion (test infrastructure only)."""


class QubitOperator:
    def __init__(self, *a, **k):
        raise NotImplementedError("openfermion stand-in")


def count_qubits(op):
    raise NotImplementedError("openfermion stand-in")
