def __getattr__(name):
    def _noop(*a, **k):
        raise RuntimeError("matplotlib is stubbed out in the oracle shims")
    return _noop
