"""Self-contained CPU restatement of the reference's environment step / reset.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED by the
reference's own tests (there are none); pinned instead against the reference's
``environment/env.py`` run verbatim under ``oracle/shims`` -- the fixtures in
``tests/golden`` -- by ``tests/test_oracle_vs_reference.py``.

Does not need ``/root/reference`` at run time, so it travels to the GPU box.
Every function cites the reference lines it restates (paths relative to
``/root/reference``).  Arithmetic is numpy float64 unless ``dtype`` says otherwise.
"""
from __future__ import annotations

import copy
import os

import numpy as np

from .diffrax_restated import (Dopri5, ODETerm, PIDController, SaveAt,
                               diffeqsolve)

# environment/env_configs/env1.py:4-20 (the one table all three env variants use,
# env.py:17-18): [stim contact, recording contact, locus] triples.
STIM_REC_LOCUS_ENV1 = [
    [[5, 2, 3], [3, 5, 1], [1, 2, 3]], [[4, 3, 1], [2, 5, 4], [2, 1, 4]],
    [[4, 3, 6], [2, 6, 4], [4, 3, 2]], [[5, 2, 1], [3, 5, 3], [5, 2, 5]],
    [[1, 3, 2], [4, 1, 4], [4, 5, 4]], [[6, 6, 4], [4, 4, 3], [3, 6, 5]],
    [[6, 5, 3], [1, 6, 4], [3, 2, 6]], [[6, 3, 5], [4, 1, 1], [5, 6, 1]],
    [[6, 5, 4], [1, 6, 3], [3, 2, 1]], [[4, 5, 3], [3, 3, 1], [6, 4, 1]],
    [[2, 3, 2], [4, 5, 3], [1, 5, 4]], [[5, 3, 2], [5, 5, 4], [5, 2, 5]],
    [[1, 6, 2], [6, 5, 1], [3, 2, 4]], [[2, 3, 3], [3, 3, 6], [1, 1, 5]],
    [[3, 5, 2], [1, 6, 4], [1, 3, 3]],
]


# ----------------------------------------------------------------- geometry
def neuron_grid_3d(gx, gy, gz, n_neurons, coord_modif=0.1):
    """utils.py:478-497 -- ``meshgrid(x,y,z).T.reshape(-1,3)``: row i is
    ``[x, y, z]`` with ``i = z*gx*gy + x*gy + y`` (y fastest)."""
    if n_neurons > gx * gy * gz:
        raise ValueError("Number of neurons should be less than grid size.")
    i = np.arange(gx * gy * gz)
    grid = np.stack([(i // gy) % gx, i % gy, i // (gx * gy)], axis=1)[:n_neurons]
    return grid * coord_modif, grid


def distance_rows(coords, rows=None):
    """utils.py:457-466 -- Euclidean distances; only the requested rows."""
    c = np.asarray(coords, dtype=np.float64)
    src = c if rows is None else c[np.atleast_1d(rows)]
    d = src[:, None, :] - c[None, :, :]
    # batched 3-vector dot: the same BLAS-style dot np.linalg.norm uses in the
    # reference's pair loop, so the distances come out bit-identical.
    return np.sqrt((d[..., None, :] @ d[..., :, None])[..., 0, 0])


def wavelet_kernel(distances, amplitude, steepness):
    """utils.py:469-475."""
    return (amplitude * (-steepness) * (12 * steepness ** 4 * distances ** 2 - 8 * steepness ** 2)
            * np.exp(-steepness * distances ** 2) / (2 * np.pi))


def coupling_matrix(neur_coords, spatial_kernel, wavelet_amp=1.0, wavelet_steepness=1.0):
    """env.py:219-229."""
    dm = distance_rows(neur_coords)
    if spatial_kernel == "cos":
        return np.cos(dm)
    if spatial_kernel == "wavelet":
        return wavelet_kernel(dm, wavelet_amp, wavelet_steepness)
    raise ValueError(f"Wrong distance matrix type: {dm}")


def contact_index(coord, grid_size):
    """env.py:94,97 (and utils.py:887)."""
    return coord[0] * grid_size[2] ** 2 + coord[1] * grid_size[1] + coord[2]


def directed_stim_masks(grid_points, center, center_idx):
    """utils.py:30-57 -- three 120-degree azimuth sectors around ``center``."""
    g = np.asarray(grid_points)
    theta = np.arctan2(g[:, 1] - center[1], g[:, 0] - center[0])
    m1 = (theta >= -np.pi / 3) & (theta < np.pi / 3)
    m2 = (theta >= np.pi / 3) & (theta <= np.pi)
    m3 = (theta >= -np.pi) & (theta < -np.pi / 3)
    for m in (m1, m2, m3):
        m[center_idx] = True
    return m1, m2, m3


def remove_negative_w0(w0):
    """utils.py:819-823 (in place, draws from the global np.random)."""
    idx = np.where(w0 <= 0.)[0]
    n = np.random.randn(len(idx)) * 0.05
    w0[idx] = np.abs(n) + np.mean(w0)
    return w0


def apply_locus_mask(w0, w_locus, lmask):
    """utils.py:902-906."""
    return w0 * (lmask * -1 + 1) + w_locus * lmask


def oscillation_locus(neur_grid, grid_size, locus_coord, locus_size):
    """utils.py:885-891."""
    l_idx = contact_index(locus_coord, grid_size)
    dist = distance_rows(np.asarray(neur_grid) * locus_size, l_idx)[0]
    return np.where(1 - dist < 0.0, 0., 1.)


def w0_samples(n):
    """utils.py:847-882 -- inverse-CDF sampling of a degree-10 polynomial PDF."""
    from scipy.integrate import quad
    from scipy.interpolate import interp1d
    y = [6, 7.7, 6, 7.7, 4, 3.5, 4, 5, 5.7, 10, 5.7, 4.9, 2.3, 1.2, 0.8, 0.75, 0.7, 0.7, 0.68]
    x = [0, 1.8, 2.5, 3.3, 4.5, 5.5, 8, 12.5, 18, 20, 22, 25, 30, 35, 40, 45, 50, 55, 60]
    poly = np.poly1d(np.polyfit(x, y, 10))
    x_range = np.linspace(np.min(x), 30, 1000)

    def pdf(v):
        return np.maximum(poly(v), 0)
    norm, _ = quad(pdf, np.min(x), np.max(x))
    cdf = np.cumsum(pdf(x_range) / norm)
    cdf /= cdf[-1]
    inv = interp1d(cdf, x_range, bounds_error=False, fill_value=(x_range[0], x_range[-1]))
    return inv(np.random.rand(n))


def w0_with_locus(n_neurons, grid_size, coord_modif, locus_center, locus_size, wmuL, wsdL):
    """utils.py:909-942 -- returns (w0, neur_coords, neur_grid, w0_without_locus,
    w_locus, locus_mask), frequencies in rad/unit (x0.065)."""
    w_deg = w0_samples(n_neurons)
    coords, grid = neuron_grid_3d(*grid_size, n_neurons, coord_modif=coord_modif)
    lmask = oscillation_locus(grid, grid_size, locus_center, locus_size)
    w_locus = np.random.uniform(low=wmuL - wsdL, high=wmuL + wsdL, size=(n_neurons))
    w_with = apply_locus_mask(w_deg, w_locus, lmask)
    return (w_with * 0.065, coords, grid, w_deg * 0.065, w_locus * 0.065, lmask)


def perturbation_walk(v0, M, step_scale):
    """env.py:21-57 -- plasticity random walk."""
    out = [v0.copy()]
    scale = np.std(v0.copy(), ddof=1)
    for _ in range(M):
        out.append(out[-1] + step_scale * scale * np.random.randn(len(out[-1])))
    return np.array(out)


# ----------------------------------------------------------------- electrode
class Electrode:
    """env.py:61-171 -- contact indices, stimulation / recording conductances."""

    def __init__(self, grid_size, neur_grid, conduct_modifier, elec_coords, rec_coords,
                 amplitudes, directed_stimulation=False, prc_type="I", naive=False):
        assert len(amplitudes) == len(elec_coords), \
            "Number of amplitudes is not equal to number of electrode coordinates!"
        scaled = np.asarray(neur_grid) * conduct_modifier            # env.py:232
        self.elec_idxs = [contact_index(c, grid_size) for c in elec_coords]
        self.rec_idxs = [contact_index(c, grid_size) for c in rec_coords]
        self.conductances = []
        for ei in self.elec_idxs:
            dv = distance_rows(scaled, ei)[0]
            cond = 1 - dv
            cond = np.where(cond < 0.0, 0, cond)
            if naive:
                cond = np.ones_like(dv)
            self.conductances.append(cond)
        self.directional_masks = []
        if directed_stimulation:
            stale_idx = self.elec_idxs[-1]                           # env.py:128-131 (stale loop var)
            for c in elec_coords:
                self.directional_masks.append(directed_stim_masks(neur_grid, np.asarray(c), stale_idx))
            self.conductances = [c * m[0] for c, m in zip(self.conductances, self.directional_masks)]
        self.rec_conductances = []
        for ri in self.rec_idxs:
            dv = distance_rows(scaled, ri)[0]
            cond = 1 - dv
            cond = np.where(cond < 0.0, 0, cond)
            if naive:
                cond = np.ones_like(dv)
            self.rec_conductances.append(cond)
        if prc_type not in ("I", "II", "Gaussian", "dummy"):
            raise ValueError("Wrong type of PRC function!")


# ----------------------------------------------------------------- ODE model
def kuramoto_rhs(y, w0, k_over_n, alpha, pulse, form="matvec"):
    """env.py:252-256.  ``form='as_written'`` materialises the N x N sine matrix
    like the reference; ``'matvec'`` uses sin(a-b)=sin a cos b - cos a sin b."""
    theta = np.fmod(y, 2 * np.pi)
    if form == "as_written":
        n = y.shape[0]
        return w0 + k_over_n * np.sum(alpha * np.sin(theta - np.tile(theta, (n, 1)).T), axis=1) + pulse
    s, c = np.sin(theta), np.cos(theta)
    return w0 + k_over_n * (c * (alpha @ s) - s * (alpha @ c)) + pulse


class KuramotoModel:
    """env.py:186-271."""

    def __init__(self, n, K, grid_size, w0, neur_coords, neur_grid, elec_coords, rec_coords,
                 conduct_modifier, spatial_kernel="cos", wavelet_amp=1.0, wavelet_steepness=1.0,
                 directed_stimulation=False, electrode_amps=(1.0,), prc_type="I", naive_dbs=False,
                 rhs_form="matvec", dtype=np.float64, alpha=None):
        self.n, self.K = n, K
        self.w0 = remove_negative_w0(w0)
        assert np.min(self.w0) >= 0, "Natural frequencies w0 must be positive!"
        self.alpha = coupling_matrix(neur_coords, spatial_kernel, wavelet_amp, wavelet_steepness) \
            if alpha is None else alpha
        self.dbs = Electrode(grid_size, neur_grid, conduct_modifier, elec_coords, rec_coords,
                             list(electrode_amps), directed_stimulation, prc_type, naive_dbs)
        self.pulse = np.zeros(n)
        self.rhs_form, self.dtype = rhs_form, np.dtype(dtype)
        self.controller = PIDController(rtol=1e-5, atol=1e-5)        # env.py:249
        self.stats = {"num_accepted_steps": 0, "num_rejected_steps": 0, "num_rhs_evals": 0}
        self.trace = None

    def forward(self, t_eval, state0):
        dt = self.dtype
        alpha = self.alpha.astype(dt, copy=False)
        w0 = self.w0.astype(dt, copy=False)
        pulse = self.pulse.astype(dt, copy=False)
        kn = dt.type(self.K / self.n)
        form = self.rhs_form
        sol = diffeqsolve(ODETerm(lambda t, y, a: kuramoto_rhs(y, w0, kn, alpha, pulse, form)),
                          Dopri5(), t0=t_eval[0], t1=t_eval[-1], dt0=0.05,
                          y0=np.asarray(state0, dtype=dt), saveat=SaveAt(ts=t_eval),
                          stepsize_controller=self.controller,
                          time_dtype=(np.float32 if dt == np.float32 else np.float64),
                          trace=self.trace)
        for k in self.stats:
            self.stats[k] += sol.stats[k]
        return sol.ys


# ----------------------------------------------------------------- rewards
def beta_band_power(sig, dt, beta_a, beta_b):
    """utils.py:21-27."""
    n = sig.shape[0]
    ft = np.abs(np.fft.rfft(sig) / n) ** 2 * 2
    freq = np.fft.rfftfreq(n, dt)
    return np.sum(ft[np.where((freq > beta_a) & (freq < beta_b))])


def band_pass(signal, fs, lowcut=12, highcut=30, order=5):
    """utils.py:794-816 (envelope dropped: no caller on the path uses it)."""
    from scipy.signal import butter, filtfilt
    nyq = 0.5 * fs
    b, a = butter(order, [lowcut / nyq, highcut / nyq], btype="band")
    return filtfilt(b, a, signal)


def reward_bbpow_action(x, u, verbose_dt):
    """env.py:638-650."""
    return -1e4 * beta_band_power(x, verbose_dt / 100, 12.5, 21) - 1e-2 * np.abs(u[0])


def reward_temp_const(x, u, verbose_dt):
    """env.py:653-666."""
    xf = band_pass(x, 1 / (verbose_dt / 100), order=2)
    return -1e3 * (xf[-1] - np.mean(xf)) ** 2 - 1e-2 * np.abs(u[0])


def reward_bbpow_threshold(x, u, verbose_dt):
    """env.py:669-688."""
    bb = 1e4 * beta_band_power(x, verbose_dt / 100, 12.5, 21)
    return -(5. if bb > 20 else 0) - np.abs(float(u[0]))


REWARDS = {"bbpow_action": reward_bbpow_action, "temp_const_action": reward_temp_const,
           "bbpow_threth_action": reward_bbpow_threshold}


# ----------------------------------------------------------------- the env
class OracleEnv:
    """env.py:274-688, same public surface (reset/step + the attributes callers read).

    ``compat_env2`` applies the two fixes SURVEY.md F7 lists so that the env2 config
    runs at all (the shipped code asserts / calls an undefined method): the
    plasticity assert is skipped and ``calc_next_temp_event`` := ``calc_next_event``.
    """

    def __init__(self, params_dict, save_init=False, rhs_form="matvec", dtype=np.float64,
                 compat_env2=False, share_alpha=True):
        p = self.params_dict = params_dict
        self.save_init, self.rhs_form, self.dtype = save_init, rhs_form, dtype
        self.compat_env2 = compat_env2
        self.reset_count = -1
        self.verbose = p["verbose"]
        np.random.seed(p["rand_seed"])                                       # env.py:291
        self.step_len = p["electrode_width"] + p["electrode_pause"]
        self.observe_wind_len = self.step_len * p["observe_wind_counts"]
        self.observe_wind_idxs = int(self.observe_wind_len / p["verbose_dt"])
        self.total_episode_counts = int(p["total_episode_len"] / self.step_len)
        self.transient_state_len = p["transient_state_len"]
        if self.transient_state_len < self.observe_wind_len:
            raise ValueError("Transient state should be longer than RL agent observation window!")
        self.dbs_action_bounds = p["dbs_action_bounds"]
        self.ppo_action_bounds = [-1., 1.]
        if p["reward_func"] not in REWARDS:
            raise ValueError("Wrong reward function!")
        self._reward = REWARDS[p["reward_func"]]
        if p["recording_kernel"] not in ("naive", "gaussian"):
            raise ValueError("Wrong recording kernel function!")
        self.w0 = p["w0"]
        self.w0_without_locus = p["w0_without_locus"]
        self.w0_without_locus_ = copy.deepcopy(p["w0_without_locus"])
        self.elec_coords, self.rec_coords = p["elec_coords"], p["rec_coords"]
        self.encapsulation_coeff = p["conduct_modifier"]
        self.temporal_events = {"electrode_drift": [], "encapsulation_drift": [],
                                "plasticity_drift": [], "mov_modulation_drift": []}
        if p["temporal_drift"]:
            self.random_freq_update = p["random_freq_update"]
            self.elec_drift_episode = p["electrode_drift_freq"]
            self.elec_encaps_episode = p["encapsulation_drift_freq"]
            self.encaps_percent = p["encapsulation_percent"]
            self.plasticity_episode = p["plasticity_drift_freq"]
            if not compat_env2:
                assert self.plasticity_episode >= 2, "Maybe set plasticity drift more rarely?"
            self.plasticity_percent = p["plasticity_percent"]
            self.reset_plasticity_episode = p["reset_plasticity_episode"]
            self.plasticity_process_count = 0
            self.w0_process = perturbation_walk(self.w0_without_locus,
                                                M=self.reset_plasticity_episode * 2,
                                                step_scale=self.plasticity_percent * 0.01)
        self.spatial_events = []
        self.spatial_var_freq = p["spatial_var_freq"]
        self.spatial_var_episode = self.spatial_var_freq
        self._alpha = None
        self._share_alpha = share_alpha
        self.current_step, self.current_time, self.done = 0, 0., False
        self.reset()

    # env.py:389-393
    def rescale_action(self, action):
        x, y = self.ppo_action_bounds
        z, k = self.dbs_action_bounds
        return z + ((k - z) * (action - x)) / (y - x)

    # env.py:396-412
    def calc_naive_lfp(self, sig):
        return np.mean(np.cos(sig), axis=1)

    def calc_lfp(self, sig):
        if self.params_dict["recording_kernel"] == "naive":
            return self.calc_naive_lfp(sig)
        rec = np.zeros(sig.shape[0])
        for cond in self.kuramoto.dbs.rec_conductances:
            rec += np.mean(np.cos(sig) * cond, axis=1)
        return rec

    # env.py:457-464
    def calc_next_event(self, f, deltas=(-1, 0, 1)):
        if self.random_freq_update:
            return np.random.choice([f + d for d in deltas])
        return f

    # env.py:415-454
    def step(self, action):
        p = self.params_dict
        self.u = [self.rescale_action(float(a)) for a in action]
        pulse = np.zeros(p["num_oscillators"])
        for amp, cond in zip(self.u, self.kuramoto.dbs.conductances):
            pulse += cond * amp
        self.kuramoto.pulse = pulse
        self.t_eval_step_I = np.arange(self.current_time, self.current_time + p["electrode_width"],
                                       p["verbose_dt"])
        self.sol_state = self.kuramoto.forward(self.t_eval_step_I, self.sol_state[-1, :])
        self.sol_state_ = self.sol_state
        self.current_time = self.t_eval_step_I[-1]
        self.kuramoto.pulse = np.zeros(p["num_oscillators"])
        self.t_eval_step_II = np.arange(self.current_time, self.current_time + p["electrode_pause"],
                                        p["verbose_dt"])
        self.sol_state = self.kuramoto.forward(self.t_eval_step_II, self.sol_state[-1, :])
        self.sol_state_ = np.concatenate([self.sol_state_, self.sol_state])
        self.current_time = self.t_eval_step_II[-1]
        self.theta_mean = self.calc_naive_lfp(self.sol_state_[:-1, :])
        self.theta_records = self.calc_lfp(self.sol_state_[:-1, :])
        self.theta_state = np.append(self.theta_state, self.theta_records[np.newaxis, ...], axis=1)
        self.theta_state = self.theta_state[:, -self.observe_wind_idxs:]
        self.current_step += 1
        self.done = self.current_step >= self.total_episode_counts
        self.reward_ = self._reward(self.theta_state[0], self.u, p["verbose_dt"])
        return (self.theta_state.astype(np.float32), self.reward_, self.done, False, {})

    # env.py:467-614
    def reset(self, seed=None, options=None):
        p = self.params_dict
        self.current_step, self.current_time, self.done = 0, 0., False
        self.reset_count += 1
        if p["temporal_drift"]:
            if self.elec_drift_episode == self.reset_count:
                self.elec_drift_episode += self.calc_next_event(p["electrode_drift_freq"], [-1, 0, 1])
                new = [[10000, 0, 0]]
                lo, hi = 1, min(p["grid_size"]) - 2
                while any(c < lo or c > hi for c in new[0]):
                    delta = np.empty(3)
                    for i in range(3):
                        delta[i] = np.random.choice([-1, 1]) * np.random.choice([0, 1])
                    new = np.asarray(self.elec_coords + delta).astype(int).tolist()
                self.elec_coords = new
                self.temporal_events["electrode_drift"].append([self.reset_count, self.elec_coords])
            if self.elec_encaps_episode == self.reset_count:
                self.elec_encaps_episode += self.calc_next_event(p["encapsulation_drift_freq"],
                                                                 [-2, -1, 0, 1, 2])
                self.encapsulation_coeff += self.encaps_percent
                self.temporal_events["encapsulation_drift"].append([self.reset_count, self.encaps_percent])
            if self.plasticity_episode == self.reset_count:
                if not self.compat_env2:
                    raise AttributeError("'SpatialKuramoto' object has no attribute "
                                         "'calc_next_temp_event'")              # env.py:520 (F7)
                self.plasticity_episode += self.calc_next_event(p["plasticity_drift_freq"], [0, 1])
                self.w0_without_locus = self.w0_process[self.plasticity_process_count]
                self.plasticity_process_count += 1
            if self.reset_count % self.reset_plasticity_episode == 0:
                self.plasticity_process_count = 0
                self.w0_without_locus = copy.deepcopy(self.w0_without_locus_)
                self.w0_process = perturbation_walk(self.w0_without_locus,
                                                    M=self.reset_plasticity_episode * 2,
                                                    step_scale=self.plasticity_percent * 0.01)
        if p["spatial_feature"]:
            if self.spatial_var_episode == self.reset_count and self.reset_count > 2:
                index = np.random.choice(len(STIM_REC_LOCUS_ENV1))
                self.elec_coords = [STIM_REC_LOCUS_ENV1[index][0]]
                self.rec_coords = [STIM_REC_LOCUS_ENV1[index][1]]
                self.spatial_var_episode += self.spatial_var_freq
                self.spatial_events.append([self.reset_count, STIM_REC_LOCUS_ENV1[index]])
        if p["save_events"] and p["log_path"] is not None and self.reset_count > 1:
            np.save(os.path.join(p["log_path"], f"temp_{self.reset_count}.npy"), self.temporal_events)
        self.w0 = apply_locus_mask(self.w0_without_locus, p["locus_without_w0"], p["locus_mask"])
        if self._share_alpha and self._alpha is None:
            self._alpha = coupling_matrix(p["neur_coords"], p["spatial_kernel"],
                                          p["wavelet_amp"], p["wavelet_steepness"])
        self.kuramoto = KuramotoModel(
            p["num_oscillators"], p["K"], p["grid_size"], self.w0, p["neur_coords"], p["neur_grid"],
            self.elec_coords, self.rec_coords, self.encapsulation_coeff, p["spatial_kernel"],
            p["wavelet_amp"], p["wavelet_steepness"], p["directed_stimulation"],
            p["electrode_amps"], p["electrode_prc_type"], p["naive_dbs"],
            rhs_form=self.rhs_form, dtype=self.dtype,
            alpha=self._alpha if self._share_alpha else None)
        if not self.save_init:
            self.init_state = np.random.normal(loc=p["init_state_mean"], scale=p["init_state_sd"],
                                               size=(p["num_oscillators"]))
            self.init_state = remove_negative_w0(self.init_state)
        self.t_eval_transient = np.arange(self.current_time, self.transient_state_len, p["verbose_dt"])
        self.current_time = self.t_eval_transient[-1]
        self.sol_state = self.kuramoto.forward(self.t_eval_transient, self.init_state)
        self.theta_record_transient = self.calc_lfp(self.sol_state[:-1, :])
        self.theta_state = self.theta_record_transient[-self.observe_wind_idxs:][np.newaxis, ...]
        return self.theta_state.astype(np.float32), {}


# ----------------------------------------------------------------- schedule helper
def step_schedule(n_steps, t_start, electrode_width, electrode_pause, verbose_dt):
    """env.py:426-441 replayed without integrating: per step the two ``np.arange`` grids.
    Returns a list of (t_eval_I, t_eval_II) float64 arrays."""
    out, ct = [], t_start
    for _ in range(n_steps):
        a = np.arange(ct, ct + electrode_width, verbose_dt)
        ct = a[-1]
        b = np.arange(ct, ct + electrode_pause, verbose_dt)
        ct = b[-1]
        out.append((a, b))
    return out
