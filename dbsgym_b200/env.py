"""``SpatialKuramoto(params_dict)`` -- the reference's gymnasium environment, GPU-backed.

Same constructor, spaces, ``reset`` / ``step`` signatures, return values, attributes and error
behaviour as reference environment/env.py:274-688; the ODE integration, LFP, observation window
and reward run in the CUDA engine (one environment == a batch of one).

Batching without editing caller scripts: ``SpatialKuramoto(params_dict)`` returns a
:class:`~dbsgym_b200.vec_env.BatchedKuramotoVecEnv` instead when ``params_dict['num_envs'] > 1``
or the environment variable ``DBSGYM_NUM_ENVS`` is set (SB3 algorithms accept a VecEnv wherever
they accept an env).
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np

from . import utils
from ._compat import Box, GymEnv
from .batched import BatchedKuramoto
from .geometry import coupling_rows


def _default_precision(params_dict):
    return params_dict.get("precision", os.environ.get("DBSGYM_PRECISION", "f32"))


class SpatialKuramoto(GymEnv):
    metadata = {"render.modes": ["human"]}

    def __new__(cls, params_dict=None, save_init=False, **kw):
        n = 1
        if params_dict is not None:
            n = int(params_dict.get("num_envs", os.environ.get("DBSGYM_NUM_ENVS", 1)))
        if cls is SpatialKuramoto and n > 1:
            from .vec_env import BatchedKuramotoVecEnv
            return BatchedKuramotoVecEnv(params_dict, num_envs=n, **kw)
        return super().__new__(cls)

    def __init__(self, params_dict, save_init=False, precision=None, device=0, compat_env2=None):
        super().__init__()
        self.save_init = save_init
        self.params_dict = params_dict
        if compat_env2 is None:
            compat_env2 = bool(params_dict.get("compat_env2", os.environ.get("DBSGYM_COMPAT_ENV2", "") == "1"))
        self._core = BatchedKuramoto([params_dict], precision=precision or _default_precision(params_dict),
                                     device=device, compat_env2=compat_env2, save_init=save_init,
                                     coupling_eval=params_dict.get("coupling_eval", "auto"))
        host = self._core.hosts[0]
        self.verbose = host.verbose
        self.step_len = host.step_len
        self.observe_wind_len = host.observe_wind_len
        self.observe_wind_idxs = host.observe_wind_idxs
        self.total_episode_len = host.total_episode_len
        self.total_episode_counts = host.total_episode_counts
        self.transient_state_len = host.transient_state_len
        self.dim = 1
        self.dbs_action_bounds = host.dbs_action_bounds
        self.ppo_action_bounds = host.ppo_action_bounds
        self.action_space = Box(low=-1., high=1., shape=(1,), dtype=np.float32)           # env.py:310-312
        self.observation_space = Box(low=-1.5, high=1.5, shape=(1, self.observe_wind_idxs),
                                     dtype=np.float32)                                    # env.py:313-315
        self.K = host.K
        self.done = False
        self.u = [0.0]
        self.reward_ = None
        self._after_reset()

    # ---- attributes the reference exposes, served from the host state / device ----------
    def _after_reset(self):
        self.done = False
        self.theta_state = self._core.engine.window_values([0])
        self.t_eval_transient = self._core.t_transient
        self.theta_mean = None
        self.theta_records = None

    @property
    def _host(self):
        return self._core.hosts[0]       # (a HostList refreshes the object from the batched reset state on access)

    @property
    def reset_count(self):
        return self._host.reset_count

    @property
    def current_step(self):
        return int(self._core.current_step[0])

    def __getattr__(self, name):
        # reference attributes that are pure host bookkeeping (reset_count, elec_drift_episode, elec_encaps_episode,
        # plasticity_episode, spatial_var_episode, w0_without_locus, ...; env.py:339-386, :483-557) live on the
        # HostEnvState; only reached when normal lookup fails
        if name.startswith("__") or name in ("_host", "_core"):
            raise AttributeError(name)
        core = self.__dict__.get("_core")
        host = core.hosts[0] if core is not None else None
        if host is not None and hasattr(host, name):
            return getattr(host, name)
        raise AttributeError(f"'SpatialKuramoto' object has no attribute '{name}'")

    @property
    def current_time(self):
        return self._core.current_time(0)

    @property
    def w0(self):
        return self._host.w0

    @property
    def elec_coords(self):
        return self._host.elec_coords

    @property
    def rec_coords(self):
        return self._host.rec_coords

    @property
    def encapsulation_coeff(self):
        return self._host.encapsulation_coeff

    @property
    def spatial_events(self):
        return self._host.spatial_events

    @property
    def temporal_events(self):
        return getattr(self._host, "temporal_events", None)

    @property
    def init_state(self):
        return self._host.init_state

    @property
    def sol_state(self):
        """Phases at the end of the last segment, shape [1, N] (``sol_state[-1, :]`` as in env.py:429)."""
        return self._core.engine.state([0])

    @property
    def kuramoto(self):
        p = self.params_dict
        el = self._core.electrodes[0]
        return SimpleNamespace(
            K=p["K"], n_neurons=p["num_oscillators"], w0=self._core.w0_model[0], grid_size=p["grid_size"],
            neur_coords=p["neur_coords"], neur_grid=p["neur_grid"], dbs=el, spatial_kernel=p["spatial_kernel"],
            alpha=_LazyAlpha(p))

    @property
    def kw0(self):
        return self._core.w0_model[0]

    @property
    def kneur_grid(self):
        return self.params_dict["neur_grid"]

    @property
    def kgrid_size(self):
        return self.params_dict["grid_size"]

    # ---- gym API --------------------------------------------------------------------------
    def rescale_action(self, action):
        x, y = self.ppo_action_bounds
        z, k = self.dbs_action_bounds
        return z + ((k - z) * (action - x)) / (y - x)

    def step(self, action):
        """env.py:415-454."""
        a = [float(v) for v in action]
        self.u = [self.rescale_action(v) for v in a]
        obs, rew, done = self._core.step(np.asarray(a[:1], dtype=np.float32))
        self.theta_mean = self._core.theta_mean(0)
        self.theta_records = self._core.theta_records(0)
        self.theta_state = self._core.engine.window_values([0])
        self.done = bool(done[0])
        self.reward_ = float(self._core.engine.rewards()[0][0])
        return (obs[:1].copy(), self.reward_, self.done, False, {})

    def reset(self, seed=None, options=None):
        """env.py:467-614 (``seed`` is ignored by the draws, exactly like the reference)."""
        super().reset(seed=seed)
        self._core.reset_envs([0])
        self._after_reset()
        return self._core.observations()[:1].copy(), {}

    def render(self, mode="human", close=False):
        pass

    def close(self):
        self._core.close()

    # ---- reward / metric helpers callers use on host arrays (env.py:625-688) -------------
    def calc_naive_lfp(self, sig):
        return np.mean(np.cos(sig), axis=1)

    def calc_distance_lfp(self, sig):
        rec = np.zeros(sig.shape[0])
        for cond in self._core.electrodes[0].rec_conductances:
            rec += np.mean(np.cos(sig) * cond, axis=1)
        return rec

    def calculate_bbpow(self, solutions):
        sig = np.concatenate(solutions)
        return utils.calc_beta_band_power(sig, utils.units2sec(self.params_dict["verbose_dt"]), 12.5, 21)

    def calculate_energy(self, actions):
        return np.abs(actions).sum()

    def reward_bbpow_action(self, x_state, action_value, baseline=False):
        assert len(x_state.shape) == 1, "Incorrect dimension of theta_state"
        dt = utils.units2sec(self.params_dict["verbose_dt"])
        return -1e4 * utils.calc_beta_band_power(x_state, dt, 12.5, 21) - 1e-2 * np.abs(action_value[0])

    def reward_temp_const_lfp_betafilt_action(self, x_state, action_value, baseline=False):
        assert len(x_state.shape) == 1, "Incorrect dimension of theta_state"
        dt = utils.units2sec(self.params_dict["verbose_dt"])
        filt, _ = utils.band_pass_envelope(x_state, 1 / dt, order=2)
        return -1e3 * (filt[-1] - np.mean(filt)) ** 2 - 1e-2 * np.abs(action_value[0])

    def reward_bbpow_threth_action(self, x_state, action_value, baseline=False):
        assert len(x_state.shape) == 1, "Incorrect dimension of theta_state"
        dt = utils.units2sec(self.params_dict["verbose_dt"])
        bb = 1e4 * utils.calc_beta_band_power(x_state, dt, 12.5, 21)
        return -(5. if bb > 20 else 0) - np.abs(float(action_value[0]))


class _LazyAlpha:
    """``env.kuramoto.alpha`` materialised only when somebody indexes it (N x N float64)."""

    def __init__(self, p):
        self._p, self._a = p, None

    def _get(self):
        if self._a is None:
            p = self._p
            self._a = coupling_rows(p["neur_coords"], np.arange(p["num_oscillators"]), p["spatial_kernel"],
                                    p["wavelet_amp"], p["wavelet_steepness"])
        return self._a

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, k):
        return self._get()[k]

    @property
    def shape(self):
        n = self._p["num_oscillators"]
        return (n, n)
