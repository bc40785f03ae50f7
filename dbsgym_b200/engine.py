"""Python face of the C-ABI step engine: one :class:`KuramotoEngine` == one ``DbsGymHandle`` == one GPU.

All compute happens in ``csrc/libdbsgym.so`` (hand-written sm_100a kernels).  This module only
marshals numpy / torch buffers into the plain-pointer calls of ``include/dbsgym.h``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .schedule import StepSchedule

_REWARD_KIND = {"bbpow_action": _capi.REWARD_BBPOW, "temp_const_action": _capi.REWARD_TEMP_CONST,
                "bbpow_threth_action": _capi.REWARD_BBPOW_THRESH}


_OWN_STREAM = C.c_void_p(-1)          # DBSGYM_OWN_STREAM


def _stream(stream):
    """None -> the handle's private stream; otherwise a raw cudaStream_t (0 = legacy default stream)."""
    return _OWN_STREAM if stream is None else C.c_void_p(int(stream))


def _ids(env_ids):
    if env_ids is None:
        return None
    return np.ascontiguousarray(env_ids, dtype=np.int32)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class KuramotoEngine:
    def __init__(self, n_envs, n_osc, grid_size, window, K, *, precision="f32", coupling_table=None,
                 alpha=None, lowrank=None, order=None, device=0, max_step_samples=20, rtol=1e-5, atol=1e-5, dt0=0.05,
                 action_bounds=(-5.0, 5.0), max_steps=4096, options=None):
        """``order``: order[d] = the caller's index of the oscillator at grid position d (dbsgym.h: dbsgym_set_oscillator_order).
        ``options``: tuning / diagnostic switches of DbsGymConfig (include/dbsgym.h) -- ``mw`` (None auto, False
        never, True always use the multi-worker step kernel), ``force_cluster``, ``ctas_per_sm`` and the boolean A/B
        switches ``no_geo1``, ``no_sym``, ``no_fsal_reuse``, ``no_fused_obs``, ``no_fast_obs``, ``no_warp_kernel``."""
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        if (coupling_table is None) == (alpha is None and lowrank is None):
            raise ValueError("give exactly one of coupling_table (GRID) or alpha / lowrank (DENSE)")
        self.lib = _capi.load()
        self.n_envs, self.n_osc, self.window = int(n_envs), int(n_osc), int(window)
        self.precision = precision
        self.max_step_samples = int(max_step_samples)
        self.coupling = "grid" if coupling_table is not None else "dense"
        cfg = _capi.DbsGymConfig()
        cfg.struct_bytes = C.sizeof(_capi.DbsGymConfig)
        cfg.device, cfg.n_envs, cfg.n_osc = int(device), self.n_envs, self.n_osc
        cfg.grid[0], cfg.grid[1], cfg.grid[2] = (int(g) for g in grid_size)
        cfg.window = self.window
        cfg.precision = _capi.F64 if precision == "f64" else _capi.F32
        cfg.coupling = _capi.COUPLING_GRID if coupling_table is not None else _capi.COUPLING_DENSE
        cfg.max_step_samples, cfg.max_steps = self.max_step_samples, int(max_steps)
        cfg.K, cfg.rtol, cfg.atol, cfg.dt0 = float(K), float(rtol), float(atol), float(dt0)
        cfg.safety, cfg.factor_min, cfg.factor_max = 0.9, 0.2, 10.0     # diffrax PIDController defaults
        cfg.action_lo, cfg.action_hi = float(action_bounds[0]), float(action_bounds[1])
        opt = dict(options or {})
        mw = opt.pop("mw", None)
        cfg.mw_mode = 0 if mw is None else (2 if mw else 1)
        cfg.force_cluster = int(opt.pop("force_cluster", 0) or 0)
        cfg.ctas_per_sm = int(opt.pop("ctas_per_sm", 0) or 0)
        flags = 0
        for name, bit in (("no_geo1", _capi.DBG_NO_GEO1), ("no_sym", _capi.DBG_NO_SYM),
                          ("no_fsal_reuse", _capi.DBG_NO_FSAL_REUSE), ("no_fused_obs", _capi.DBG_NO_FUSED_OBS),
                          ("no_fast_obs", _capi.DBG_NO_FAST_OBS), ("no_warp_kernel", _capi.DBG_NO_WARP_KERNEL)):
            if opt.pop(name, False):
                flags |= bit
        if opt:
            raise ValueError(f"unknown engine options: {sorted(opt)}")
        cfg.debug_flags = flags
        self.options = dict(options or {})
        self._h = C.c_void_p()
        rc = self.lib.dbsgym_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self.lib.dbsgym_last_error(None)
            self._h = None
            raise _capi.DbsGymError(f"dbsgym_create failed ({rc}): {msg.decode() if msg else '?'}")
        self.device = int(device)
        self.order = None
        if order is not None:                        # the caller's oscillators are a permutation of the regular grid
            o = np.ascontiguousarray(order, dtype=np.int32)
            assert o.shape == (self.n_osc,)
            self._ck(self.lib.dbsgym_set_oscillator_order(self._h, _capi.ptr(o)))
            self.order = o
        if coupling_table is not None:
            t = _f64(coupling_table)
            self._ck(self.lib.dbsgym_set_coupling_grid(self._h, _capi.ptr(t)))
        elif alpha is not None:
            a = _f64(alpha, (self.n_osc, self.n_osc))
            if not np.array_equal(a, a.T):
                raise ValueError("dense coupling must be symmetric")
            self._ck(self.lib.dbsgym_set_coupling_dense(self._h, _capi.ptr(a)))
        self.schedule = None
        self.spectral = None
        self.lowrank = None
        if lowrank is not None:                      # (vecs [r][N], vals [r], residual): the operator in low-rank form
            self.set_coupling_lowrank(lowrank[0], lowrank[1], lowrank[2] if len(lowrank) > 2 else None)

    def set_coupling_spectral(self, vecs, vals, ranks, residual=None):
        """Switch the GRID operator to its spectral form (dbsgym.h: dbsgym_set_coupling_spectral).  ``vecs`` [8][n_osc / 8][r_max]
        (8 x 8 x 8 grid: 64 octant points, 8 x 8 x 4: 32), ``vals`` [8][r_max], ``ranks`` [8] as returned by geometry.spectral_factors; ``ranks=None`` switches back."""
        if ranks is None:
            self._ck(self.lib.dbsgym_set_coupling_spectral(self._h, None, 0, None, None))
            self.spectral = None
            return
        v, w = _f64(vecs), _f64(vals)
        r8 = np.ascontiguousarray(ranks, dtype=np.int32)
        assert r8.shape == (8,) and v.shape[:2] == (8, self.n_osc // 8) and w.shape == (8, v.shape[2])
        self._ck(self.lib.dbsgym_set_coupling_spectral(self._h, _capi.ptr(r8), int(v.shape[2]), _capi.ptr(v), _capi.ptr(w)))
        self.spectral = {"ranks": [int(r) for r in ranks], "modes": int(sum(ranks)), "residual": residual}

    def set_coupling_lowrank(self, vecs, vals, residual=None):
        """Switch a DENSE fp32 handle to the low-rank form of its operator (dbsgym.h: dbsgym_set_coupling_lowrank).
        ``vecs`` [rank][n_osc], ``vals`` [rank] as returned by geometry.lowrank_factors; ``vecs=None`` switches back."""
        if vecs is None:
            self._ck(self.lib.dbsgym_set_coupling_lowrank(self._h, 0, None, None))
            self.lowrank = None
            return
        v, w = _f64(vecs), _f64(vals)
        assert v.ndim == 2 and v.shape[1] == self.n_osc and w.shape == (v.shape[0],)
        self._ck(self.lib.dbsgym_set_coupling_lowrank(self._h, int(v.shape[0]), _capi.ptr(v), _capi.ptr(w)))
        self.lowrank = {"rank": int(v.shape[0]), "residual": residual}

    def set_coupling_lowrank_sectors(self, soff, zvecs, vals, residual=None):
        """Sector form of the low-rank operator (dbsgym.h: dbsgym_set_coupling_lowrank_sectors; fp32 GRID handles, before any
        set_env_params): ``soff`` [9], ``zvecs`` [modes][n_osc / 8], ``vals`` [modes] as returned by
        geometry.grid_sector_factors."""
        o = np.ascontiguousarray(soff, dtype=np.int32)
        z, w = _f64(zvecs), _f64(vals)
        assert o.shape == (9,) and z.shape == (int(o[8]), self.n_osc // 8) and w.shape == (int(o[8]),)
        self._ck(self.lib.dbsgym_set_coupling_lowrank_sectors(self._h, _capi.ptr(o), _capi.ptr(z), _capi.ptr(w)))
        self.lowrank = {"rank": int(np.count_nonzero(w)), "padded_modes": int(o[8]), "sectors": True, "residual": residual}

    # ------------------------------------------------------------------ plumbing
    def _ck(self, rc):
        _capi.check(self.lib, self._h, rc)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.dbsgym_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ set-up
    def set_env_params(self, env_ids=None, w0=None, stim=None, rec=None, y0=None):
        ids = _ids(env_ids)
        n = self.n_envs if ids is None else len(ids)
        shape = (n, self.n_osc)
        w0, stim, rec, y0 = (_f64(v, shape) for v in (w0, stim, rec, y0))
        self._ck(self.lib.dbsgym_set_env_params(self._h, _capi.ptr(ids), n, _capi.ptr(w0), _capi.ptr(stim),
                                                _capi.ptr(rec), _capi.ptr(y0)))

    def set_recording(self, weighted: bool):
        self._ck(self.lib.dbsgym_set_recording(self._h, 1 if weighted else 0))

    def set_schedule(self, sched: StepSchedule):
        if sched.max_samples > self.max_step_samples:
            raise ValueError("schedule needs more samples per step than max_step_samples")
        nI = np.ascontiguousarray(sched.n_I, dtype=np.int32)
        nII = np.ascontiguousarray(sched.n_II, dtype=np.int32)
        oI, oII = _f64(sched.offs_I), _f64(sched.offs_II)
        self._ck(self.lib.dbsgym_set_schedule(self._h, sched.n_steps, _capi.ptr(nI), _capi.ptr(nII),
                                              _capi.ptr(oI), sched.max_I, _capi.ptr(oII), sched.max_II))
        self.schedule = sched

    def set_reward(self, reward_func, verbose_dt, lin_functional=None):
        """reward_func: one of params_dict['reward_func'] (env.py:323-330)."""
        from .utils import beta_bins, temp_const_functional, units2sec
        spec = _capi.DbsGymRewardSpec()
        spec.struct_bytes = C.sizeof(_capi.DbsGymRewardSpec)
        spec.kind = _REWARD_KIND[reward_func]
        dt_sec = units2sec(verbose_dt)
        spec.bin_lo, spec.bin_hi = beta_bins(self.window, dt_sec, 12.5, 21)      # env.py:644
        spec.power_scale, spec.temp_scale = 1e4, 1e3
        spec.threshold, spec.threshold_penalty = 20.0, 5.0
        spec.action_cost = 1.0 if reward_func == "bbpow_threth_action" else 1e-2
        g = None
        if reward_func == "temp_const_action":
            g = _f64(lin_functional if lin_functional is not None
                     else temp_const_functional(self.window, 1 / dt_sec, order=2), (self.window,))
        self._ck(self.lib.dbsgym_set_reward(self._h, C.byref(spec), _capi.ptr(g)))

    def set_episode(self, env_ids=None, step_idx=None, episode_len=None):
        ids = _ids(env_ids)
        n = self.n_envs if ids is None else len(ids)

        def arr(v):
            if v is None:
                return None
            return np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.int32), (n,)))
        s, e = arr(step_idx), arr(episode_len)
        self._ck(self.lib.dbsgym_set_episode(self._h, _capi.ptr(ids), n, _capi.ptr(s), _capi.ptr(e)))

    # ------------------------------------------------------------------ hot path
    def transient(self, ts, env_ids=None, obs_dev_ptr=None, stream=None):
        ids = _ids(env_ids)
        n = self.n_envs if ids is None else len(ids)
        offs = _f64(np.asarray(ts, dtype=np.float64) - float(ts[0]))
        self._ck(self.lib.dbsgym_transient(self._h, _capi.ptr(ids), n, _capi.ptr(offs), len(offs),
                                           obs_dev_ptr, _stream(stream)))

    def step_host(self, actions, obs=None, reward=None, done=None):
        """One batched env step through host buffers (pinned numpy / torch memory is fastest)."""
        a = np.ascontiguousarray(actions, dtype=np.float32).reshape(-1)
        if a.shape[0] != self.n_envs:
            raise ValueError(f"expected {self.n_envs} actions")
        if obs is None:
            obs = np.empty((self.n_envs, self.window), dtype=np.float32)
        if reward is None:
            reward = np.empty(self.n_envs, dtype=np.float32)
        if done is None:
            done = np.empty(self.n_envs, dtype=np.uint8)
        self._ck(self.lib.dbsgym_step_host(self._h, _capi.ptr(a), _capi.ptr(obs), _capi.ptr(reward),
                                           _capi.ptr(done)))
        return obs, reward, done

    def step_host_samples(self, actions, samples, n_samples, reward, done):
        """Delta-transfer step: only the new window samples come back ([B, max_step_samples] f32)."""
        self._ck(self.lib.dbsgym_step_host_samples(self._h, _capi.ptr(actions), _capi.ptr(samples),
                                                   _capi.ptr(n_samples), _capi.ptr(reward), _capi.ptr(done)))

    def host_mirror(self):
        """The CURRENT pinned [B, 2 (W + 256)] float32 host log the GPU appends the window samples to (dbsgym.h:
        dbsgym_host_mirror).  Ask again after every transient(): a reset moves on to the other buffer."""
        ptr, row = C.POINTER(C.c_float)(), C.c_int32(0)
        self._ck(self.lib.dbsgym_host_mirror(self._h, C.byref(ptr), C.byref(row)))
        return np.ctypeslib.as_array(ptr, shape=(self.n_envs, row.value))

    def step_host_mirror(self, actions, reward, done):
        """One step; the host mirror is updated by the GPU.  Returns (pos, n_new): the chronological window of
        every environment is mirror[:, pos:pos+W]; n_new == -1 means the environments are out of lockstep."""
        cpos, cn = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.dbsgym_step_host_mirror(self._h, _capi.ptr(actions), C.byref(cpos), C.byref(cn),
                                                  _capi.ptr(reward), _capi.ptr(done)))
        return cpos.value, cn.value

    def step_host_mirror_begin(self, actions):
        """Launch one host-mirror step and return immediately (finish it with step_host_mirror_end)."""
        self._ck(self.lib.dbsgym_step_host_mirror_begin(self._h, _capi.ptr(actions)))

    def step_host_mirror_end(self, reward, done):
        cpos, cn = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.dbsgym_step_host_mirror_end(self._h, C.byref(cpos), C.byref(cn), _capi.ptr(reward),
                                                      _capi.ptr(done)))
        return cpos.value, cn.value

    def step_device(self, actions_ptr, obs_ptr=None, reward_ptr=None, done_ptr=None, stream=None):
        """Asynchronous step on raw device pointers (e.g. ``tensor.data_ptr()``)."""
        self._ck(self.lib.dbsgym_step(self._h, actions_ptr, obs_ptr, reward_ptr, done_ptr, _stream(stream)))

    def obs_host(self, out=None):
        if out is None:
            out = np.empty((self.n_envs, self.window), dtype=np.float32)
        self._ck(self.lib.dbsgym_get_obs_host(self._h, _capi.ptr(out)))
        return out

    # ------------------------------------------------------------------ introspection
    def lfp(self):
        t = np.empty((self.n_envs, self.max_step_samples))
        r = np.empty((self.n_envs, self.max_step_samples))
        n = np.empty(self.n_envs, dtype=np.int32)
        self._ck(self.lib.dbsgym_get_lfp(self._h, _capi.ptr(t), _capi.ptr(r), _capi.ptr(n)))
        return t, r, n

    def rewards(self):
        r, u = np.empty(self.n_envs), np.empty(self.n_envs)
        self._ck(self.lib.dbsgym_get_rewards(self._h, _capi.ptr(r), _capi.ptr(u)))
        return r, u

    def state(self, env_ids=None):
        ids = _ids(env_ids)
        n = self.n_envs if ids is None else len(ids)
        y = np.empty((n, self.n_osc))
        self._ck(self.lib.dbsgym_get_phases(self._h, _capi.ptr(ids), n, _capi.ptr(y)))
        return y

    def get_state(self):
        """Opaque snapshot of the whole handle (dbsgym.h: dbsgym_get_state) as a uint8 array."""
        nb = C.c_uint64()
        self._ck(self.lib.dbsgym_state_bytes(self._h, C.byref(nb)))
        blob = np.empty(nb.value, dtype=np.uint8)
        self._ck(self.lib.dbsgym_get_state(self._h, _capi.ptr(blob), nb.value))
        return blob

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._ck(self.lib.dbsgym_set_state(self._h, _capi.ptr(blob), blob.size))

    def launch_count(self, reset=False):
        """Kernels launched through this handle since create / the last reset of the count."""
        v = C.c_uint64()
        self._ck(self.lib.dbsgym_launch_count(self._h, C.byref(v), 1 if reset else 0))
        return v.value

    def window_values(self, env_ids=None):
        ids = _ids(env_ids)
        n = self.n_envs if ids is None else len(ids)
        w = np.empty((n, self.window))
        self._ck(self.lib.dbsgym_get_window(self._h, _capi.ptr(ids), n, _capi.ptr(w)))
        return w

    def set_window(self, window, env_ids=None):
        ids = _ids(env_ids)
        n = self.n_envs if ids is None else len(ids)
        w = _f64(window, (n, self.window))
        self._ck(self.lib.dbsgym_set_window(self._h, _capi.ptr(ids), n, _capi.ptr(w)))

    def episode(self):
        s = np.empty(self.n_envs, dtype=np.int32)
        d = np.empty(self.n_envs, dtype=np.uint8)
        self._ck(self.lib.dbsgym_get_episode(self._h, _capi.ptr(s), _capi.ptr(d)))
        return s, d

    def counters(self, reset=False):
        a, r, f = C.c_uint64(), C.c_uint64(), C.c_uint64()
        st = C.c_int32()
        self._ck(self.lib.dbsgym_counters(self._h, C.byref(a), C.byref(r), C.byref(f), C.byref(st),
                                          1 if reset else 0))
        return {"accepted": a.value, "rejected": r.value, "rhs_evals": f.value, "status": st.value}

    def rhs_reused(self):
        """How many of the counted RHS evaluations were carried over instead of executed (dbsgym.h: dbsgym_rhs_reused);
        read it BEFORE counters(reset=True)."""
        v = C.c_uint64()
        self._ck(self.lib.dbsgym_rhs_reused(self._h, C.byref(v)))
        return v.value

    # ------------------------------------------------------------------ evaluation metric on the device
    def trace_begin(self, capacity):
        """Record the TRUE LFP (theta_mean) of every following step on the device, up to `capacity` samples per env."""
        self._trace_cap = int(capacity)
        self._ck(self.lib.dbsgym_trace_begin(self._h, self._trace_cap))

    def trace_end(self):
        self._ck(self.lib.dbsgym_trace_end(self._h))

    def trace(self):
        """(trace [B, capacity] float64, lengths [B] int32) copied to the host."""
        t = np.empty((self.n_envs, self._trace_cap))
        n = np.empty(self.n_envs, dtype=np.int32)
        self._ck(self.lib.dbsgym_trace_get(self._h, _capi.ptr(t), _capi.ptr(n)))
        return t, n

    def trace_lengths(self):
        n = np.empty(self.n_envs, dtype=np.int32)
        self._ck(self.lib.dbsgym_trace_get(self._h, None, _capi.ptr(n)))
        return n

    def eval_bbpow(self, b, a, zi, padlen, k_lo, weights):
        """Beta-band power of every environment's recorded trace (dbsgym.h: dbsgym_eval_bbpow)."""
        spec = _capi.DbsGymEvalSpec()
        spec.struct_bytes = C.sizeof(_capi.DbsGymEvalSpec)
        spec.padlen = int(padlen)
        for i in range(5):
            spec.b[i], spec.a[i] = float(b[i]), float(a[i])
        for i in range(4):
            spec.zi[i] = float(zi[i])
        w = _f64(weights)
        spec.k_lo, spec.n_k = int(k_lo), int(w.size)
        out = np.empty(self.n_envs)
        self._ck(self.lib.dbsgym_eval_bbpow(self._h, C.byref(spec), _capi.ptr(w), _capi.ptr(out)))
        return out

    def step_variant(self, n_envs=None):
        """Which step-kernel variant a launch over n_envs environments uses (dbsgym.h: dbsgym_step_variant)."""
        return int(self.lib.dbsgym_step_variant(self._h, self.n_envs if n_envs is None else int(n_envs)))

    def set_timing(self, enabled=True):
        self._ck(self.lib.dbsgym_set_timing(self._h, 1 if enabled else 0))

    def last_step_ms(self):
        ms = (C.c_float * 2)()
        self._ck(self.lib.dbsgym_last_step_ms(self._h, ms))
        return float(ms[0]), float(ms[1])


def measure_mufu_peak(device=0, ms_target=20.0):
    """MUFU (sin / cos special-function unit) peak in 1e12 operations per second."""
    lib = _capi.load()
    out = C.c_double()
    rc = lib.dbsgym_measure_mufu_peak(int(device), float(ms_target), C.byref(out))
    if rc != 0:
        raise _capi.DbsGymError(f"MUFU peak measurement failed ({rc})")
    return out.value


def measure_fp32_peak(device=0, ms_target=20.0, packed=None):
    """FP32 FMA peak in TFLOP/s: packed=None best of both, False scalar FFMA, True FFMA2."""
    lib = _capi.load()
    out = C.c_double()
    if packed is None:
        rc = lib.dbsgym_measure_fp32_peak(int(device), float(ms_target), C.byref(out))
    else:
        rc = lib.dbsgym_measure_fp32_peak_mode(int(device), float(ms_target), 1 if packed else 0, C.byref(out))
    if rc != 0:
        raise _capi.DbsGymError(f"fp32 peak measurement failed ({rc})")
    return out.value
