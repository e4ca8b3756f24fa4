class SparsePauliOp:
    def __init__(self, *a, **k):
        raise NotImplementedError("qiskit stand-in")


class Statevector:
    def __init__(self, *a, **k):
        raise NotImplementedError("qiskit stand-in")


def random_clifford(*a, **k):
    raise NotImplementedError("qiskit stand-in")
