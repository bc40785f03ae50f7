import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def make_params(cfg_name, seed, reward="bbpow_action", **over):
    """Build a full params_dict the way aDBS_RL/train_aDBS_RL.py:95-112 does, with the PRODUCT's host code."""
    import copy
    import importlib
    from dbsgym_b200 import utils
    cfg = importlib.import_module(f"dbsgym_b200.configs.{cfg_name}")
    base = cfg.params_dict_train
    np.random.seed(seed)
    w0, nc, ng, w0t, wl, lm = utils.generate_w0_with_locus(
        cfg.n_neurons, cfg.grid_size, cfg.coord_modif, locus_center=base["locus_center"],
        locus_size=base["locus_size"], wmuL=base["wmuL"], wsdL=base["wsdL"], show=False)
    d = copy.deepcopy(base)
    d.update(w0=w0, w0_without_locus=w0t, locus_without_w0=wl, locus_mask=lm, neur_coords=nc,
             neur_grid=ng, reward_func=reward, verbose=0)
    d.update(over)
    return d
