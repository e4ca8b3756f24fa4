def ncon(*a, **k):
    raise NotImplementedError("ncon stand-in")
