"""Import the reference's ``environment/env.py`` VERBATIM on top of the shims.

Works only where ``/root/reference`` exists (this container, not the GPU box).
Used by ``tests/golden/make_golden.py`` and by the CPU tests that cross-check the
self-contained restatement (``oracle/kuramoto_oracle.py``) against the real
reference module.
"""
from __future__ import annotations

import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("DBSGYM_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "environment", "env.py"))


def load_reference():
    """Return (env_module, utils_module, {name: config_module}) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    # our own repo ships an `environment` package too: make sure the
    # reference's wins for this import and gets its own module names back.
    saved = {k: v for k, v in sys.modules.items()
             if k == "environment" or k.startswith("environment.")}
    for k in saved:
        del sys.modules[k]
    path0 = list(sys.path)
    import oracle.diffrax_restated  # noqa: F401  (resolved through sys.modules once the path changes)
    # the reference's `environment` has no __init__.py (namespace package): any regular package of
    # that name -- ours -- would win regardless of path order, so hide those path entries.
    sys.path[:] = [_SHIMS, REFERENCE_ROOT] + [
        q for q in path0 if not os.path.isfile(os.path.join(q or ".", "environment", "__init__.py"))]
    try:
        env = importlib.import_module("environment.env")
        utils = importlib.import_module("environment.utils")
        cfgs = {n: importlib.import_module(f"environment.env_configs.{n}")
                for n in ("env0", "env1", "env2")}
    finally:
        sys.path[:] = path0
        ref_mods = {k: v for k, v in sys.modules.items()
                    if k == "environment" or k.startswith("environment.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
        for k in ("gymnasium", "gymnasium.spaces", "jax", "jax.numpy", "diffrax",
                  "matplotlib", "matplotlib.pyplot", "seaborn", "imageio",
                  "mpl_toolkits", "mpl_toolkits.axes_grid1"):
            m = sys.modules.get(k)
            if m is not None and getattr(m, "__file__", "") and _SHIMS in (m.__file__ or ""):
                del sys.modules[k]
    return env, utils, cfgs
