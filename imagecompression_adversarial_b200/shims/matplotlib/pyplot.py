def __getattr__(name):
    def _noop(*a, **k):
        raise NotImplementedError(f"matplotlib.pyplot.{name}: plotting is outside the attack hot path")
    return _noop
