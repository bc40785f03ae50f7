"""ctypes binding of ``include/dbsgym.h`` (the C-ABI of the CUDA step engine).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``)
into ``csrc/libdbsgym.so``.  There is no CPU fallback: if the library is missing or no
CUDA device is present every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# DBSGYM_LIB selects an alternative build of the same sources (tuning variants, scripts/build_variants.sh)
LIB_PATH = os.environ.get("DBSGYM_LIB") or os.path.join(CSRC, "libdbsgym.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "dbsgym.h")

ABI_VERSION = 3
F32, F64 = 0, 1
COUPLING_GRID, COUPLING_DENSE = 0, 1
REWARD_BBPOW, REWARD_TEMP_CONST, REWARD_BBPOW_THRESH = 0, 1, 2

EXPORTS = [
    "dbsgym_abi_version", "dbsgym_build_flags", "dbsgym_step_variant", "dbsgym_create", "dbsgym_destroy", "dbsgym_last_error",
    "dbsgym_set_coupling_grid", "dbsgym_set_coupling_dense", "dbsgym_set_coupling_spectral", "dbsgym_set_coupling_lowrank", "dbsgym_set_coupling_lowrank_sectors", "dbsgym_set_oscillator_order", "dbsgym_set_env_params",
    "dbsgym_set_recording", "dbsgym_set_schedule", "dbsgym_set_reward", "dbsgym_set_episode",
    "dbsgym_transient", "dbsgym_step", "dbsgym_step_host", "dbsgym_step_host_samples", "dbsgym_host_mirror", "dbsgym_step_host_mirror", "dbsgym_step_host_mirror_begin", "dbsgym_step_host_mirror_end", "dbsgym_get_obs_host",
    "dbsgym_get_lfp", "dbsgym_get_rewards", "dbsgym_get_phases", "dbsgym_get_window",
    "dbsgym_np_gauss", "dbsgym_np_choice", "dbsgym_np_reset_draws",
    "dbsgym_state_bytes", "dbsgym_get_state", "dbsgym_set_state", "dbsgym_launch_count", "dbsgym_measure_mufu_peak",
    "dbsgym_set_window", "dbsgym_get_episode", "dbsgym_counters", "dbsgym_last_step_ms",
    "dbsgym_set_timing", "dbsgym_measure_fp32_peak", "dbsgym_measure_fp32_peak_mode",
    "dbsgym_rhs_reused", "dbsgym_trace_begin", "dbsgym_trace_end", "dbsgym_trace_get", "dbsgym_eval_bbpow",
]


class DbsGymConfig(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("device", C.c_int32), ("n_envs", C.c_int32),
        ("n_osc", C.c_int32), ("grid", C.c_int32 * 3), ("window", C.c_int32),
        ("precision", C.c_int32), ("coupling", C.c_int32), ("max_step_samples", C.c_int32),
        ("max_steps", C.c_int32), ("K", C.c_double), ("rtol", C.c_double), ("atol", C.c_double),
        ("dt0", C.c_double), ("safety", C.c_double), ("factor_min", C.c_double),
        ("factor_max", C.c_double), ("action_lo", C.c_double), ("action_hi", C.c_double),
        ("mw_mode", C.c_int32), ("force_cluster", C.c_int32), ("ctas_per_sm", C.c_int32), ("debug_flags", C.c_uint32),
    ]


# DbsGymConfig.debug_flags
DBG_NO_GEO1, DBG_NO_SYM, DBG_NO_FSAL_REUSE, DBG_NO_FUSED_OBS, DBG_NO_FAST_OBS, DBG_NO_WARP_KERNEL = 1, 2, 4, 8, 16, 32


class DbsGymRewardSpec(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("kind", C.c_int32), ("bin_lo", C.c_int32),
        ("bin_hi", C.c_int32), ("power_scale", C.c_double), ("action_cost", C.c_double),
        ("threshold", C.c_double), ("threshold_penalty", C.c_double), ("temp_scale", C.c_double),
    ]


class DbsGymEvalSpec(C.Structure):
    _fields_ = [
        ("struct_bytes", C.c_uint32), ("padlen", C.c_int32), ("b", C.c_double * 5), ("a", C.c_double * 5),
        ("zi", C.c_double * 4), ("k_lo", C.c_int32), ("n_k", C.c_int32),
    ]


class DbsGymNpState(C.Structure):
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int32), ("has_gauss", C.c_int32), ("gauss", C.c_double)]


class DbsGymResetPlan(C.Structure):
    _fields_ = [("struct_bytes", C.c_uint32), ("n_envs", C.c_int32), ("n_osc", C.c_int32), ("walk_len", C.c_int32),
                ("coord_lo", C.c_int32), ("coord_hi", C.c_int32), ("table_len", C.c_int32),
                ("random_freq_update", C.c_int32), ("refix_cap_rows", C.c_int32), ("refix_cap_noise", C.c_int32),
                ("reserved", C.c_int32 * 2), ("init_mean", C.c_double), ("init_sd", C.c_double)]


RESET_ELECTRODE_MOVE, RESET_ENCAPSULATION, RESET_PLASTICITY, RESET_WALK_REGEN, RESET_SPATIAL = 1, 2, 4, 8, 16


class DbsGymError(RuntimeError):
    pass


UNITS = ("api", "host_rng", "step_f32_grid", "step_f32_sym", "step_f32_lines", "step_f32_dense", "step_f32_lowrank", "step_f32_mw", "step_f32_spectral", "step_f32_warp", "step_f32_warp1", "step_f32_oct",
         "step_f32_cluster", "step_f64_grid", "step_f64_sym", "step_f64_dense")
HEADERS = ("step_kernel.cuh", "warp_kernel.cuh", "warp1_kernel.cuh", "oct_kernel.cuh", "step_launch.h", "obs_kernel.cuh", "eval_kernel.cuh")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def sources():
    return [os.path.join(CSRC, u + ".cu") for u in UNITS] + [os.path.join(CSRC, h) for h in HEADERS] + [HEADER]


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile the translation units of ``csrc/`` for sm_100a (in parallel; cross-compiles without a GPU) and link them into
    ``csrc/libdbsgym.so``.  Objects go to ``csrc/build/`` and are reused while their sources are older."""
    if os.environ.get("DBSGYM_LIB"):
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise DbsGymError("nvcc not found: cannot build libdbsgym.so")
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + list(extra_flags) + (["-Xptxas=-v"] if verbose else [])
    stamp = os.path.join(objdir, "flags.txt")
    if not os.path.exists(stamp) or open(stamp).read() != " ".join(flags):
        force = True
    hdr_time = max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS)
    hdr_time = max(hdr_time, os.path.getmtime(HEADER))

    def compile_unit(u):
        src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(objdir, u + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
            return u, False, ""
        res = subprocess.run([nvcc] + flags + ["-c", "-o", obj + ".tmp.o", src], capture_output=True, text=True)
        if res.returncode != 0:
            raise DbsGymError(f"nvcc failed on {u}.cu:\n" + res.stdout + res.stderr)
        os.replace(obj + ".tmp.o", obj)
        return u, True, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_unit, UNITS))
    rebuilt = [u for u, did, _ in results if did]
    if verbose:
        for u, did, log in results:
            if did:
                print(f"---- {u}\n{log}")
    objs = [os.path.join(objdir, u + ".o") for u in UNITS]
    if rebuilt or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(o) for o in objs):
        res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH + ".tmp"] + objs,
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise DbsGymError("link failed:\n" + res.stdout + res.stderr)
        os.replace(LIB_PATH + ".tmp", LIB_PATH)
    with open(stamp, "w") as f:
        f.write(" ".join(flags))
    return LIB_PATH


_lib = None


def load():
    """Load the shared library and declare every prototype of ``include/dbsgym.h``."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DbsGymError(f"{LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, i32p, f32p, f64p, u8p, u64p = (C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                       C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64))
    P = {
        "dbsgym_abi_version": (C.c_int, []),
        "dbsgym_build_flags": (C.c_int, []),
        "dbsgym_step_variant": (C.c_int, [vp, C.c_int32]),
        "dbsgym_create": (C.c_int, [C.POINTER(DbsGymConfig), C.POINTER(vp)]),
        "dbsgym_destroy": (None, [vp]),
        "dbsgym_last_error": (C.c_char_p, [vp]),
        "dbsgym_set_coupling_grid": (C.c_int, [vp, vp]),
        "dbsgym_set_coupling_dense": (C.c_int, [vp, vp]),
        "dbsgym_set_coupling_spectral": (C.c_int, [vp, vp, C.c_int32, vp, vp]),
        "dbsgym_set_coupling_lowrank": (C.c_int, [vp, C.c_int32, vp, vp]),
        "dbsgym_set_coupling_lowrank_sectors": (C.c_int, [vp, vp, vp, vp]),
        "dbsgym_set_oscillator_order": (C.c_int, [vp, vp]),
        "dbsgym_set_env_params": (C.c_int, [vp, vp, C.c_int32, vp, vp, vp, vp]),
        "dbsgym_set_recording": (C.c_int, [vp, C.c_int32]),
        "dbsgym_set_schedule": (C.c_int, [vp, C.c_int32, vp, vp, vp, C.c_int32, vp, C.c_int32]),
        "dbsgym_set_reward": (C.c_int, [vp, C.POINTER(DbsGymRewardSpec), vp]),
        "dbsgym_set_episode": (C.c_int, [vp, vp, C.c_int32, vp, vp]),
        "dbsgym_transient": (C.c_int, [vp, vp, C.c_int32, vp, C.c_int32, vp, vp]),
        "dbsgym_step": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "dbsgym_step_host": (C.c_int, [vp, vp, vp, vp, vp]),
        "dbsgym_step_host_samples": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "dbsgym_host_mirror": (C.c_int, [vp, C.POINTER(f32p), i32p]),
        "dbsgym_step_host_mirror": (C.c_int, [vp, vp, i32p, i32p, vp, vp]),
        "dbsgym_step_host_mirror_begin": (C.c_int, [vp, vp]),
        "dbsgym_step_host_mirror_end": (C.c_int, [vp, i32p, i32p, vp, vp]),
        "dbsgym_get_obs_host": (C.c_int, [vp, vp]),
        "dbsgym_get_lfp": (C.c_int, [vp, vp, vp, vp]),
        "dbsgym_get_rewards": (C.c_int, [vp, vp, vp]),
        "dbsgym_get_phases": (C.c_int, [vp, vp, C.c_int32, vp]),
        "dbsgym_state_bytes": (C.c_int, [vp, u64p]),
        "dbsgym_get_state": (C.c_int, [vp, vp, C.c_uint64]),
        "dbsgym_set_state": (C.c_int, [vp, vp, C.c_uint64]),
        "dbsgym_launch_count": (C.c_int, [vp, u64p, C.c_int32]),
        "dbsgym_measure_mufu_peak": (C.c_int, [C.c_int32, C.c_double, f64p]),
        "dbsgym_np_gauss": (C.c_int, [C.POINTER(DbsGymNpState), C.c_int64, C.c_double, C.c_double, vp]),
        "dbsgym_np_choice": (C.c_int, [C.POINTER(DbsGymNpState), C.c_int64, C.c_uint32, vp]),
        "dbsgym_np_reset_draws": (C.c_int, [C.POINTER(DbsGymNpState), C.POINTER(DbsGymResetPlan)] + [vp] * 12),
        "dbsgym_get_window": (C.c_int, [vp, vp, C.c_int32, vp]),
        "dbsgym_set_window": (C.c_int, [vp, vp, C.c_int32, vp]),
        "dbsgym_get_episode": (C.c_int, [vp, vp, vp]),
        "dbsgym_counters": (C.c_int, [vp, u64p, u64p, u64p, i32p, C.c_int32]),
        "dbsgym_last_step_ms": (C.c_int, [vp, f32p]),
        "dbsgym_set_timing": (C.c_int, [vp, C.c_int32]),
        "dbsgym_rhs_reused": (C.c_int, [vp, u64p]),
        "dbsgym_trace_begin": (C.c_int, [vp, C.c_int32]),
        "dbsgym_trace_end": (C.c_int, [vp]),
        "dbsgym_trace_get": (C.c_int, [vp, vp, vp]),
        "dbsgym_eval_bbpow": (C.c_int, [vp, C.POINTER(DbsGymEvalSpec), vp, vp]),
        "dbsgym_measure_fp32_peak": (C.c_int, [C.c_int32, C.c_double, f64p]),
        "dbsgym_measure_fp32_peak_mode": (C.c_int, [C.c_int32, C.c_double, C.c_int32, f64p]),
    }
    assert sorted(P) == sorted(EXPORTS)
    for name, (res, args) in P.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.dbsgym_abi_version() != ABI_VERSION:
        raise DbsGymError("libdbsgym.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def check(lib, handle, rc):
    if rc != 0:
        msg = lib.dbsgym_last_error(handle)
        raise DbsGymError(f"dbsgym error {rc}: {msg.decode() if msg else '?'}")
