def __getattr__(name):
    raise NotImplementedError("matplotlib stand-in: plotting is outside the hot path")
