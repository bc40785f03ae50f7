"""The C-ABI library loads here (no GPU) and exports every symbol include/dbsgym.h declares."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from dbsgym_b200 import _capi


@pytest.fixture(scope="module")
def lib():
    _capi.build()
    return _capi.load()


def test_header_and_exports_agree(lib):
    hdr = open(os.path.join(ROOT, "include", "dbsgym.h")).read()
    declared = set(re.findall(r"^\s*(?:int|void|const char\*)\s+\*?(dbsgym_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    raw = C.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.dbsgym_abi_version() == _capi.ABI_VERSION == 3


def test_struct_layouts_match_header():
    # sizes are checked inside the library as well (struct_bytes); these guard the ctypes mirror
    assert C.sizeof(_capi.DbsGymConfig) == 4 * 12 + 8 * 9 + 4 * 4
    assert C.sizeof(_capi.DbsGymRewardSpec) == 4 * 4 + 8 * 5
    assert C.sizeof(_capi.DbsGymEvalSpec) == 4 * 2 + 8 * 14 + 4 * 2


def test_create_rejects_bad_config_without_touching_a_gpu(lib):
    cfg = _capi.DbsGymConfig()
    h = C.c_void_p()
    cfg.struct_bytes = 3
    assert lib.dbsgym_create(C.byref(cfg), C.byref(h)) == -1
    assert b"size mismatch" in lib.dbsgym_last_error(None)
    cfg.struct_bytes = C.sizeof(_capi.DbsGymConfig)
    assert lib.dbsgym_create(C.byref(cfg), C.byref(h)) == -1        # zero sizes
    assert not h.value

    def grid_cfg(n_osc, grid, precision=_capi.F32):
        c = _capi.DbsGymConfig()
        c.struct_bytes = C.sizeof(_capi.DbsGymConfig)
        c.device, c.n_envs, c.n_osc, c.window = 0, 4, n_osc, 2340
        c.grid[0], c.grid[1], c.grid[2] = grid
        c.precision, c.coupling, c.max_step_samples, c.max_steps = precision, _capi.COUPLING_GRID, 20, 4096
        c.K, c.rtol, c.atol, c.dt0 = 0.52, 1e-5, 1e-5, 0.05
        c.safety, c.factor_min, c.factor_max, c.action_lo, c.action_hi = 0.9, 0.2, 10.0, -5.0, 5.0
        return c

    # grid shapes the structured kernels cannot serve are refused with a message (before any device is touched)
    for c, msg in ((grid_cfg(8 * 24 * 8, (8, 24, 8)), b"gy"),                              # lines of 24
                   (grid_cfg(4096, (16, 16, 16), _capi.F64), b"gy = 16 supports fp32"),        # lines of 16 in fp64
                   (grid_cfg(16 * 16 * 3, (16, 16, 16)), b"even gx and gz"),                  # odd number of z-planes
                   (grid_cfg(500, (8, 8, 8)), b"whole z-planes")):                           # not whole planes
        assert lib.dbsgym_create(C.byref(c), C.byref(h)) == -1
        assert msg in lib.dbsgym_last_error(None), lib.dbsgym_last_error(None)
        assert not h.value


def test_no_cpu_fallback(lib):
    """Without a CUDA device the engine must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dbsgym_b200.engine import KuramotoEngine
    import numpy as np
    with pytest.raises(_capi.DbsGymError, match="no CPU path|no CUDA|CUDA"):
        KuramotoEngine(1, 512, [8, 8, 8], 2340, 0.52, coupling_table=np.ones(512))


def test_release_library_reads_no_environment_variables():
    """Tuning / A-B switches travel in DbsGymConfig (mw_mode, force_cluster, ctas_per_sm, debug_flags): the library
    itself must not consult the process environment."""
    csrc = os.path.join(ROOT, "dbsgym_b200", "csrc")
    files = [f for f in os.listdir(csrc) if f.endswith((".cu", ".cuh", ".h"))]
    assert "api.cu" in files and "step_kernel.cuh" in files
    for f in files:
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "dbsgym_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f
