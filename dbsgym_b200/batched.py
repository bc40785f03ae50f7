"""A batch of reference-semantics environments on one GPU.

``BatchedKuramoto`` = B x (host bookkeeping of reference environment/env.py, see host_env.py)
+ one :class:`KuramotoEngine` that integrates all B environments per call.  It behaves like a
sequential ``DummyVecEnv`` of reference envs built from the same ``params_dict`` list: same
construction / reset order, same global-``np.random`` stream, same schedule, same outputs.
"""
from __future__ import annotations

import numpy as np

from .engine import KuramotoEngine
from .geometry import coupling_rows, coupling_table
from .host_env import HostEnvState
from .schedule import StepSchedule, transient_grid

# sector rank lists compiled into the half-grid kernel (csrc/step_f32_warp1.cu)
HALF_GRID_RANK_SETS = ((5, 4, 4, 2, 4, 4, 2, 1), (9, 4, 4, 3, 4, 4, 3, 1))
_SHARED_KEYS = ("num_oscillators", "grid_size", "K", "spatial_kernel", "wavelet_amp", "wavelet_steepness",
                "electrode_width", "electrode_pause", "verbose_dt", "observe_wind_counts",
                "transient_state_len", "dbs_action_bounds", "reward_func", "recording_kernel",
                # the coupling operator is built once, from the first dict: the neuron geometry must be common
                "neur_coords", "neur_grid")


def _pinned(shape, dtype):
    """Page-locked host buffer (through torch) when a GPU is present, plain numpy otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.float32: torch.float32, np.uint8: torch.uint8, np.float64: torch.float64}[dtype]
            t = torch.empty(shape, dtype=tdt, pin_memory=True)
            return t.numpy(), t          # caller keeps `t` alive: the numpy view borrows its memory
    except Exception:                                                    # noqa: BLE001
        pass
    return np.empty(shape, dtype=dtype), None


class BatchedKuramoto:
    def __init__(self, params_dicts, *, precision="f32", device=0, compat_env2=False, force_dense=False,
                 save_init=False, transfer="delta", engine_options=None, coupling_eval="auto", spectral_tol=1e-9,
                 batch_resets=True, prepare_resets_ahead=None):
        """``coupling_eval``: how the float32 kernels evaluate the coupling sum of env.py:252-256 on the 8 x 8 x 8 grid --
        ``"exact"``: the parity-sector block contraction (the same sum as the reference, reassociated); ``"spectral"``:
        the generalised mean-field identity over the eigenmodes of alpha above ``spectral_tol * |lambda_max|``
        (geometry.spectral_factors; 32 modes at the default 1e-9 -- ranks 7 + 4 x 6 + 1 over the eight parity sectors --
        truncation error 2.2e-7 in the spectral norm of alpha, i.e. ~1e-11 rad per time unit in d theta / dt, four orders
        below float32 rounding of the exact sum; 1e-10 gives 34 modes);
        ``"auto"`` = spectral where it applies (float32, regular 8 x 8 x 8 grid, ranks within the compiled range), else
        exact.  float64 (parity mode) always evaluates the exact sum.
        ``prepare_resets_ahead``: let the batched host reset (host_batch.HostBatch) prepare the next reset of all environments on
        a worker thread while the GPU steps (same numbers and same final np.random state as the synchronous path; None = for
        batches of 256 environments and more)."""
        if isinstance(params_dicts, dict):
            params_dicts = [params_dicts]
        self.params_dicts = list(params_dicts)
        self.num_envs = B = len(self.params_dicts)
        p0 = self.params_dicts[0]
        for k in _SHARED_KEYS:
            for d in self.params_dicts[1:]:
                if not _same(d[k], p0[k]):
                    raise ValueError(f"params_dict['{k}'] must be identical for all environments of one batch")
        self.precision = precision
        # --- host state, in index order: construct (seeds the global RNG) then first reset draws ---
        self.hosts, setups = [], []
        for d in self.params_dicts:
            h = HostEnvState(d, save_init=save_init, compat_env2=compat_env2)
            self.hosts.append(h)
            setups.append(h.begin_episode())
        h0 = self.hosts[0]
        self.n_osc = int(p0["num_oscillators"])
        self.window = h0.observe_wind_idxs
        self.t_transient = transient_grid(h0.transient_state_len, p0["verbose_dt"])
        max_len = max(h.total_episode_counts for h in self.hosts)
        self.schedule = StepSchedule(max_len + 1, self.t_transient[-1], p0["electrode_width"],
                                     p0["electrode_pause"], p0["verbose_dt"])
        # --- coupling operator (shared by every env: it depends only on the neuron grid) ---
        # lines of 16 (gy = 16) are handled by the fp32 structured kernel only: fp64 parity mode takes the DENSE path
        if int(p0["grid_size"][1]) != 8 and precision == "f64":
            force_dense = True
        table = None if force_dense else coupling_table(p0["neur_coords"], p0["neur_grid"], p0["grid_size"],
                                                        p0["spatial_kernel"], p0["wavelet_amp"],
                                                        p0["wavelet_steepness"])
        alpha = lowrank = order = None
        if coupling_eval not in ("auto", "exact", "spectral"):
            raise ValueError("coupling_eval must be 'auto', 'exact' or 'spectral'")
        if table is None and not force_dense and precision == "f32" and coupling_eval != "exact":
            # a shuffled regular grid (utils.py:490 shuffle=True) is a permutation of the regular one: the library stores it in
            # grid order (the permutation is applied at the ABI boundary) and runs the structured / spectral GRID kernels
            from .geometry import grid_permutation
            order = grid_permutation(p0["neur_grid"], p0["grid_size"])
            if order is not None:
                table = coupling_table(np.asarray(p0["neur_coords"])[order], np.asarray(p0["neur_grid"])[order], p0["grid_size"],
                                       p0["spatial_kernel"], p0["wavelet_amp"], p0["wavelet_steepness"])
                if table is None:
                    order = None
        if table is None:
            alpha = coupling_rows(p0["neur_coords"], np.arange(self.n_osc), p0["spatial_kernel"],
                                  p0["wavelet_amp"], p0["wavelet_steepness"])
            if coupling_eval != "exact" and precision == "f32":
                # any neuron ordering (utils.py:490 shuffle=True, partial grids): alpha in its truncated eigenbasis
                from .geometry import lowrank_factors
                lowrank = lowrank_factors(alpha, tol=spectral_tol)
        self.engine = KuramotoEngine(B, self.n_osc, p0["grid_size"], self.window, p0["K"],
                                     precision=precision, coupling_table=table, alpha=alpha, lowrank=lowrank, order=order, device=device,
                                     max_step_samples=max(self.schedule.max_samples, 20),
                                     action_bounds=p0["dbs_action_bounds"], options=engine_options)
        self.coupling_eval = "lowrank" if lowrank is not None else "exact"
        gs = [int(g) for g in p0["grid_size"]]
        if (table is not None and coupling_eval != "exact" and precision == "f32" and self.n_osc >= 1024 and
                self.n_osc == gs[0] * gs[1] * gs[2] and all(g % 2 == 0 for g in gs)):
            # large regular grids: the operator as sector-wise eigenpairs over the fundamental octant (O(N r / 8) per evaluation;
            # from N = 1024 up faster than the exact sector blocks -- profiles/r02_sweep_n_lowrank_sectors_1gpu.jsonl)
            from .geometry import grid_sector_factors
            f = grid_sector_factors(table, gs[0], gs[1], gs[2], tol=spectral_tol)
            if f is not None:
                self.engine.set_coupling_lowrank_sectors(*f)
                self.coupling_eval = "lowrank"
        grid888 = table is not None and [int(g) for g in p0["grid_size"]] == [8, 8, 8] and self.n_osc == 512
        if coupling_eval != "exact" and precision == "f32" and grid888 and not (engine_options or {}).get("no_sym"):
            from .geometry import spectral_factors
            vecs, vals, ranks, residual = spectral_factors(table, 8, 8, 8, tol=spectral_tol)
            if max(ranks) <= 9:
                self.engine.set_coupling_spectral(vecs, vals, ranks, residual)
                self.coupling_eval = "spectral"
        grid884 = table is not None and gs == [8, 8, 8] and self.n_osc == 256
        if coupling_eval != "exact" and precision == "f32" and grid884 and not (engine_options or {}).get("no_sym") \
                and not (engine_options or {}).get("no_warp_kernel"):
            # the half grid of BASELINE configs[4]'s smallest point: 32-point sector blocks, one octant point per lane
            from .geometry import spectral_factors
            vecs, vals, ranks, residual = spectral_factors(table, 8, 8, 4, tol=spectral_tol)
            if any(all(r <= c for r, c in zip(ranks, cs)) for cs in HALF_GRID_RANK_SETS):
                self.engine.set_coupling_spectral(vecs, vals, ranks, residual)
                self.coupling_eval = "spectral"
        if coupling_eval == "spectral" and self.coupling_eval == "exact":
            raise ValueError("coupling_eval='spectral' needs float32 and either the regular 8 x 8 x 8 grid with at most 9 modes "
                             "per sector, its 8 x 8 x 4 half with ranks inside a compiled list, or a dense operator with at most 256 modes above spectral_tol")
        self.engine.set_recording(p0["recording_kernel"] == "gaussian")
        self.engine.set_schedule(self.schedule)
        self.engine.set_reward(p0["reward_func"], p0["verbose_dt"])
        self.electrodes = [None] * B
        self.w0_model = [None] * B
        self._pin = []
        # two observation buffers used alternately: the array returned by step() stays valid until
        # the step after the next one, so callers need not copy 9.4 KB x B per step
        self._obs_bufs = []
        for _ in range(2):
            a, t = _pinned((B, self.window), np.float32); self._pin.append(t); self._obs_bufs.append(a)
        self._obs_flip = 0
        self.obs_buf = self._obs_bufs[0]
        self.rew_buf, t = _pinned((B,), np.float32); self._pin.append(t)
        self.done_buf, t = _pinned((B,), np.uint8); self._pin.append(t)
        self.act_buf, t = _pinned((B,), np.float32); self._pin.append(t)
        # delta transfer: the host mirrors the observation window and receives only the step's new
        # samples (B x 19 floats instead of B x 2340) -- see dbsgym_step_host_samples
        if transfer not in ("delta", "full"):
            raise ValueError("transfer must be 'delta' or 'full'")
        self.transfer = transfer
        smax = self.engine.max_step_samples
        self.samples_buf, t = _pinned((B, smax), np.float32); self._pin.append(t)
        self.nsamp_buf, t = _pinned((B,), np.int32); self._pin.append(t)
        # host mirror: a pinned sample log per environment, appended to by the step kernel itself (dbsgym.h:
        # dbsgym_host_mirror).  The window is always one contiguous slice of it and later steps only append BEHIND
        # a window already handed out, so step() returns views that stay intact for >= 13 further steps; a reset
        # switches to the second buffer (fetched again lazily).
        self._mirror = None
        self._mirror_used = False
        self._lfp_cache = None
        self.current_step = np.zeros(B, dtype=np.int64)
        self._episode_counts = np.array([h.total_episode_counts for h in self.hosts], dtype=np.int32)
        self._upload_and_run_transient(np.arange(B), setups)
        # configurations whose resets draw more than Gaussians (temporal drift events, spatial re-draws): the host side of
        # every later reset runs batched (host_batch.py: one native call replays numpy's global stream for all of them)
        self.host_batch = None
        if (p0["temporal_drift"] or p0["spatial_feature"]) and batch_resets:
            from .host_batch import HostBatch, HostList
            if HostBatch.supported(self.hosts):
                self.host_batch = HostBatch(self.hosts, speculate=prepare_resets_ahead)
                self.hosts = HostList(self.hosts, self.host_batch)

    # -------------------------------------------------------------------------------------
    def _upload_and_run_transient(self, ids, setups):
        ids = np.asarray(ids, dtype=np.int32)
        n, N = len(ids), self.n_osc
        # Per-oscillator vectors that did not change since this environment's last reset are not uploaded again:
        # without drift the natural frequencies and the electrode (stimulation / recording conductances) are the
        # very same objects from reset to reset (host_env.begin_episode_fast), only the initial phases are new.
        same_w0 = all(self.w0_model[i] is s.w0 for i, s in zip(ids, setups))
        same_el = all(self.electrodes[i] is s.electrode for i, s in zip(ids, setups))
        y0 = np.empty((n, N))
        for r, s in enumerate(setups):
            y0[r] = s.y0
        w0 = stim = rec = None
        if not same_w0:
            w0 = np.empty((n, N))
            for r, s in enumerate(setups):
                w0[r] = s.w0
        if not same_el:
            stim, rec = np.empty((n, N)), np.empty((n, N))
            for r, s in enumerate(setups):
                stim[r] = s.stim; rec[r] = s.rec
        self.engine.set_env_params(ids, w0=w0, stim=stim, rec=rec, y0=y0)
        self.engine.set_episode(ids, step_idx=0, episode_len=self._episode_counts[ids])
        for i, s in zip(ids, setups):
            self.electrodes[i] = s.electrode
            self.w0_model[i] = s.w0
            self.current_step[i] = 0
        self.engine.transient(self.t_transient, env_ids=ids)
        self._lfp_cache = None
        self._mirror = None            # the reset moved the mirror on to its other buffer

    def reset_envs(self, ids):
        """reset() of the listed environments, in the order given (reference env.py:467-614)."""
        ids = list(ids)
        if self.host_batch is not None:
            w0, stim, rec, y0, electrodes, el_ch, w0_ch = self.host_batch.begin_episodes(ids, changed_only=True)
            ids_a = np.asarray(ids, dtype=np.int32)
            # the initial phases are new every time; natural frequencies and electrode conductances only where they changed
            if w0_ch.all() and el_ch.all():
                self.engine.set_env_params(ids_a, w0=w0, stim=stim, rec=rec, y0=y0)
            else:
                self.engine.set_env_params(ids_a, y0=y0)
                if w0_ch.any():
                    self.engine.set_env_params(ids_a[w0_ch], w0=np.ascontiguousarray(w0[w0_ch]))
                if el_ch.any():
                    self.engine.set_env_params(ids_a[el_ch], stim=stim, rec=rec)
            self.engine.set_episode(ids_a, step_idx=0, episode_len=self._episode_counts[ids_a])
            for r, i in enumerate(ids):
                self.electrodes[i] = electrodes[r]
                self.w0_model[i] = w0[r]
            self.current_step[ids_a] = 0
            self.engine.transient(self.t_transient, env_ids=ids_a)
            self._lfp_cache = None
            self._mirror = None
            return
        setups = self._begin_episodes_fast(ids)
        if setups is None:
            setups = [self.hosts[i].begin_episode() for i in ids]
        self._upload_and_run_transient(ids, setups)

    def _begin_episodes_fast(self, ids):
        """Vectorised begin_episode() for environments whose reset only draws Gaussians (the fix of non-positive
        w0 entries, then the initial phases).  All of them come from ONE np.random.standard_normal call, which
        consumes the global legacy stream exactly like the sequential per-environment randn / normal calls
        (same Box-Muller generator, carried cache).  If an initial phase is <= 0 (remove_negative_w0 would draw
        again) the RNG is rewound and the sequential path runs instead."""
        hosts = [self.hosts[i] for i in ids]
        if not hosts:
            return None
        ks = [h.fast_reset_draws() for h in hosts]
        if any(k is None for k in ks):
            return None
        p0 = hosts[0].params_dict
        mean, sd = p0["init_state_mean"], p0["init_state_sd"]
        if any(h.params_dict["init_state_mean"] != mean or h.params_dict["init_state_sd"] != sd for h in hosts[1:]):
            return None
        N = self.n_osc
        state = np.random.get_state()
        z = np.random.standard_normal(sum(ks) + len(hosts) * N)
        y_all = mean + sd * z                    # loc + scale * gauss, exactly what np.random.normal computes
        if np.any(y_all <= 0.):                  # (also trips on a fix-noise slot, which only costs the slow path)
            np.random.set_state(state)
            return None
        setups, pos = [], 0
        for h, k in zip(hosts, ks):
            setups.append(h.begin_episode_fast(z[pos:pos + k], y_all[pos + k:pos + k + N]))
            pos += k + N
        return setups

    def observations(self):
        """Current observation windows [B, W] float32 (host), read back from the device into one of two pinned
        buffers used alternately: the returned array stays intact until the call after the next one."""
        self._obs_flip ^= 1
        self.obs_buf = self._obs_bufs[self._obs_flip]
        return self.engine.obs_host(self.obs_buf)

    def observations_after_reset(self):
        """Windows of all environments right after reset_envs(): with the delta transfer a view of the mirror buffer
        the reset switched to (every log restarts at column 0 there), otherwise a read-back."""
        if self.transfer == "delta" and self._mirror_used:
            self._mirror = self.engine.host_mirror()
            g = self._mirror.shape[1] // 2 - self.window
            return self._mirror[:, g:g + self.window]
        return self.observations()

    def step_begin(self, actions):
        """Launch one step of every environment and return at once (the GPU works); finish with step_end().
        Only the delta-transfer (host mirror) path is asynchronous; the full-observation path runs in step_end."""
        self.act_buf[:] = np.asarray(actions, dtype=np.float32).reshape(self.num_envs)
        self._lfp_cache = None
        self._pending = True
        if self.transfer == "delta":
            if self._mirror is None:
                self._mirror = self.engine.host_mirror()
                self._mirror_used = True
            self.engine.step_host_mirror_begin(self.act_buf)

    def step_end(self):
        if not getattr(self, "_pending", False):
            raise RuntimeError("step_end() without step_begin()")
        self._pending = False
        if self.transfer == "delta":
            pos, n = self.engine.step_host_mirror_end(self.rew_buf, self.done_buf)
            self.current_step += 1
            if n >= 0:
                return self._mirror[:, pos:pos + self.window], self.rew_buf, self.done_buf.view(np.bool_)
            return self.observations(), self.rew_buf, self.done_buf.view(np.bool_)   # out of lockstep
        self._obs_flip ^= 1
        self.obs_buf = self._obs_bufs[self._obs_flip]
        self.engine.step_host(self.act_buf, self.obs_buf, self.rew_buf, self.done_buf)
        self.current_step += 1
        return self.obs_buf, self.rew_buf, self.done_buf.view(np.bool_)

    def step(self, actions):
        """Advance every environment by one step.  Returns host views: obs [B,W] f32, reward [B] f32 and done [B]
        bool.  reward / done are overwritten by the next call; obs stays intact for at least one further step()
        (delta transfer: >= 13 steps, see dbsgym_host_mirror; full transfer: two alternating buffers)."""
        self.act_buf[:] = np.asarray(actions, dtype=np.float32).reshape(self.num_envs)
        self._lfp_cache = None
        if self.transfer == "delta":
            if self._mirror is None:
                self._mirror = self.engine.host_mirror()
                self._mirror_used = True
            pos, n = self.engine.step_host_mirror(self.act_buf, self.rew_buf, self.done_buf)
            self.current_step += 1
            if n >= 0:
                return self._mirror[:, pos:pos + self.window], self.rew_buf, self.done_buf.view(np.bool_)
            return self.observations(), self.rew_buf, self.done_buf.view(np.bool_)   # out of lockstep
        self._obs_flip ^= 1
        self.obs_buf = self._obs_bufs[self._obs_flip]
        self.engine.step_host(self.act_buf, self.obs_buf, self.rew_buf, self.done_buf)
        self.current_step += 1
        return self.obs_buf, self.rew_buf, self.done_buf.view(np.bool_)

    def step_tensor(self, actions, obs=None, reward=None, done=None):
        """On-device step for GPU-resident policies: torch CUDA tensors in, torch CUDA tensors out, no
        host copies at all; asynchronous on torch's current stream.  actions: float32 [B]."""
        import torch
        dev = torch.device("cuda", self.engine.device)
        B, W = self.num_envs, self.window
        actions = actions.to(device=dev, dtype=torch.float32).contiguous().view(B)
        if obs is None:
            obs = torch.empty((B, W), dtype=torch.float32, device=dev)
        if reward is None:
            reward = torch.empty(B, dtype=torch.float32, device=dev)
        if done is None:
            done = torch.empty(B, dtype=torch.uint8, device=dev)
        self.engine.step_device(actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), done.data_ptr(),
                                torch.cuda.current_stream(dev).cuda_stream)
        self.current_step += 1
        self._lfp_cache = None
        return obs, reward, done

    # -------------------------------------------------------------------------------------
    def _lfp(self):
        if self._lfp_cache is None:
            self._lfp_cache = self.engine.lfp()
        return self._lfp_cache

    def theta_mean(self, i):
        t, _, n = self._lfp()
        return t[i, :n[i]].copy()

    def theta_records(self, i):
        _, r, n = self._lfp()
        return r[i, :n[i]].copy()

    def u(self, i):
        return [float(self.engine.rewards()[1][i])]

    def current_time(self, i):
        k = int(self.current_step[i])
        return float(self.t_transient[-1]) if k == 0 else float(self.schedule.t_after[k - 1])

    def close(self):
        if self.host_batch is not None:
            self.host_batch.close()
        self.engine.close()


def _same(a, b):
    try:
        return bool(np.all(np.asarray(a) == np.asarray(b)))
    except Exception:                                                    # noqa: BLE001
        return a == b
