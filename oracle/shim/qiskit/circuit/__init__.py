class QuantumCircuit:
    def __init__(self, *a, **k):
        raise NotImplementedError("qiskit stand-in")


class ParameterVector:
    def __init__(self, *a, **k):
        raise NotImplementedError("qiskit stand-in")
