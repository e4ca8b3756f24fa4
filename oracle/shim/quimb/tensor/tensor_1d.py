class MatrixProductOperator:  # placeholder, never instantiated by the oracle path
    pass


class MatrixProductState:
    pass
