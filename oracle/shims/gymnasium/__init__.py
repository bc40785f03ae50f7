"""Minimal stand-in for gymnasium (reference: environment/env.py:7-8, :274, :286, :471)."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    render_mode = None

    def reset(self, *, seed=None, options=None):
        # gymnasium seeds its own np_random here; the reference draws from the
        # global np.random instead (env.py:291,595), so nothing to do.
        return None

    def close(self):
        pass
