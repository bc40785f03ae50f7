"""Per-environment HOST logic of the reference env: everything ``__init__`` / ``reset`` do besides
integrating (reference environment/env.py:277-386, :467-598).

It reproduces the reference's use of the *global* ``np.random`` stream draw for draw
(SURVEY.md Appendix C), so that a batch of these driven in index order is bit-identical to a
sequential ``DummyVecEnv`` of reference envs: drift events, spatial re-draws, the plasticity random
walk, ``remove_negative_w0`` and the N(pi, 0.6) initial phases.
"""
from __future__ import annotations

import copy
import os
from dataclasses import dataclass

import numpy as np

from .geometry import ElectrodeModel
from .utils import apply_locus_mask, remove_negative_w0

REWARD_NAMES = ("bbpow_action", "temp_const_action", "bbpow_threth_action")
RECORDING_KERNELS = ("naive", "gaussian")


def stim_rec_table():
    # env.py:17-18: the env1 table is used by all three variants
    from .configs.env1 import stim_rec_locus_coordinates
    return stim_rec_locus_coordinates


def generate_perturbations(initial_vector, M=10, step_scale=0.1, random_seed=None):
    """env.py:21-57 -- plasticity drift: Gaussian random walk of the w0 vector."""
    if random_seed is not None:
        np.random.seed(random_seed)
    sigma = step_scale * np.std(initial_vector.copy(), ddof=1)
    walk = [initial_vector.copy()]
    for _ in range(M):
        walk.append(walk[-1] + sigma * np.random.randn(len(walk[-1])))
    return np.array(walk)


_ELECTRODE_CACHE = {}


def cached_electrode(p, conduct_modifier, elec_coords, rec_coords, verbose):
    """Electrode geometry depends only on (grid, contacts, conduct_modifier, flags): thousands of
    batched environments share a handful of distinct electrodes, so build each one once."""
    grid = p["neur_grid"]
    key = (id(grid), tuple(p["grid_size"]), float(conduct_modifier), repr(elec_coords), repr(rec_coords),
           repr(p["electrode_amps"]), bool(p["directed_stimulation"]), p["electrode_prc_type"], bool(p["naive_dbs"]))
    hit = _ELECTRODE_CACHE.get(key)
    if hit is not None and hit[0] is grid and not verbose:
        return hit[1]
    el = ElectrodeModel(p["grid_size"], grid, conduct_modifier, elec_coords, rec_coords, p["electrode_amps"],
                        directed_stimulation=p["directed_stimulation"], prc_type=p["electrode_prc_type"],
                        naive=p["naive_dbs"], verbose=bool(verbose))
    if len(_ELECTRODE_CACHE) > 4096:
        _ELECTRODE_CACHE.clear()
    _ELECTRODE_CACHE[key] = (grid, el)          # holding `grid` keeps id(grid) unique while cached
    return el


@dataclass
class EpisodeSetup:
    """What one reset hands to the device."""
    w0: np.ndarray          # after apply_locus_mask + remove_negative_w0
    stim: np.ndarray        # conductance of the first stimulation contact
    rec: np.ndarray         # summed recording conductance
    y0: np.ndarray          # initial phases
    electrode: ElectrodeModel


class HostEnvState:
    """Episode bookkeeping of one environment (no integration, no device access)."""

    def __init__(self, params_dict, save_init=False, compat_env2=False):
        p = self.params_dict = params_dict
        self.save_init = save_init
        self.compat_env2 = compat_env2
        self.reset_count = -1
        self.verbose = p["verbose"]
        np.random.seed(p["rand_seed"])                                    # env.py:291

        self.step_len = p["electrode_width"] + p["electrode_pause"]
        self.observe_wind_len = self.step_len * p["observe_wind_counts"]
        self.observe_wind_idxs = int(self.observe_wind_len / p["verbose_dt"])
        self.total_episode_len = p["total_episode_len"]
        self.total_episode_counts = int(self.total_episode_len / self.step_len)
        self.transient_state_len = p["transient_state_len"]
        if self.transient_state_len < self.observe_wind_len:
            raise ValueError("Transient state should be longer than RL agent observation window!")
        self.dbs_action_bounds = p["dbs_action_bounds"]
        self.ppo_action_bounds = [-1., 1.]
        if p["reward_func"] not in REWARD_NAMES:
            raise ValueError("Wrong reward function!")
        if p["recording_kernel"] not in RECORDING_KERNELS:
            raise ValueError("Wrong recording kernel function!")
        if p["spatial_kernel"] not in ("cos", "wavelet"):
            raise ValueError(f"Wrong distance matrix type: {p['spatial_kernel']}")

        self.K = p["K"]
        self.w0 = p["w0"]
        self.w0_without_locus = p["w0_without_locus"]
        self.w0_without_locus_ = copy.deepcopy(p["w0_without_locus"])
        self.elec_coords = p["elec_coords"]
        self.rec_coords = p["rec_coords"]
        self.save_events = p["save_events"]
        self.encapsulation_coeff = p["conduct_modifier"]

        if p["temporal_drift"]:
            self.random_freq_update = p["random_freq_update"]
            if self.save_events:
                self.temporal_events = {"electrode_drift": [], "encapsulation_drift": [],
                                        "plasticity_drift": [], "mov_modulation_drift": []}
            self.elec_drift_episode = p["electrode_drift_freq"]
            self.elec_encaps_episode = p["encapsulation_drift_freq"]
            self.encaps_precent = p["encapsulation_percent"]
            self.mov_mod_episode = p["mov_modulation_drift_freq"]
            self.plasticity_episode = p["plasticity_drift_freq"]
            if not compat_env2:
                assert self.plasticity_episode >= 2, "Maybe set plasticity drift more rarely?"
            self.plasticity_percent = p["plasticity_percent"]
            self.reset_plasticity_episode = p["reset_plasticity_episode"]
            self.plasticity_process_count = 0
            self.rng = np.random.default_rng(seed=p["rand_seed"])         # created, never used (env.py:374)
            self.w0_process = generate_perturbations(self.w0_without_locus,
                                                     M=self.reset_plasticity_episode * 2,
                                                     step_scale=self.plasticity_percent * 0.01)
        elif self.verbose:
            print("temporal drift is off")
        self.spatial_events = []
        self.spatial_var_freq = p["spatial_var_freq"]
        self.spatial_var_episode = self.spatial_var_freq
        self.init_state = None
        self._fast_cache = None

    # env.py:457-464
    def calc_next_event(self, f, deltas=(-1, 0, 1)):
        if self.random_freq_update:
            return np.random.choice([f + d for d in deltas])
        return f

    def calc_next_temp_event(self, f, deltas=(0, 1)):
        """Undefined in the reference (env.py:520 raises AttributeError)."""
        if not self.compat_env2:
            raise AttributeError("'SpatialKuramoto' object has no attribute 'calc_next_temp_event'")
        return self.calc_next_event(f, deltas)

    # ---- batched fast path (used by BatchedKuramoto.reset_envs) ------------------------------------
    def fast_reset_draws(self):
        """Number of standard-normal draws the next reset makes BEFORE the initial phases (the
        remove_negative_w0 fix of non-positive natural frequencies, utils.py:819-823), or None when the next
        reset does anything else with the random stream or the configuration: temporal drift, a spatial
        re-draw due at the next reset count, event logging, save_init.  One ordinary begin_episode() must
        have run before (it fills the cache)."""
        p = self.params_dict
        if p["temporal_drift"] or self.save_init or self._fast_cache is None:
            return None
        nxt = self.reset_count + 1
        if p["spatial_feature"] and self.spatial_var_episode == nxt and nxt > 2:
            return None
        if p["save_events"] and p["log_path"] is not None and nxt > 1:
            return None
        return int(self._fast_cache[1].size)

    def begin_episode_fast(self, z_fix, y0_row) -> EpisodeSetup:
        """begin_episode() given this environment's slices of ONE batched standard-normal draw: ``z_fix``
        feeds remove_negative_w0(w0), ``y0_row`` are the initial phases."""
        self.reset_count += 1
        w0_raw, bad, electrode = self._fast_cache
        w0 = w0_raw
        if bad.size:
            w0 = w0_raw.copy()
            w0[bad] = np.abs(z_fix * 0.05) + np.mean(w0_raw)
        self.w0 = w0
        self.init_state = y0_row
        return EpisodeSetup(w0=w0, stim=electrode.stim_vector(), rec=electrode.rec_vector(), y0=y0_row,
                            electrode=electrode)

    def begin_episode(self) -> EpisodeSetup:
        """env.py:467-598: everything reset() does before the transient integration."""
        p = self.params_dict
        self.reset_count += 1
        if p["temporal_drift"]:
            if self.elec_drift_episode == self.reset_count:
                self.elec_drift_episode += self.calc_next_event(p["electrode_drift_freq"], [-1, 0, 1])
                lo, hi = 1, min(p["grid_size"]) - 2
                moved = [[10000, 0, 0]]
                while any(c < lo or c > hi for c in moved[0]):
                    delta = np.empty(3)
                    for axis in range(3):
                        delta[axis] = np.random.choice([-1, 1]) * np.random.choice([0, 1])
                    moved = np.asarray(self.elec_coords + delta).astype(int).tolist()
                self.elec_coords = moved
                if self.save_events:
                    self.temporal_events["electrode_drift"].append([self.reset_count, self.elec_coords])
                if self.verbose:
                    print(f"[reset {self.reset_count}] electrode moved to {self.elec_coords}")
            if self.elec_encaps_episode == self.reset_count:
                self.elec_encaps_episode += self.calc_next_event(p["encapsulation_drift_freq"], [-2, -1, 0, 1, 2])
                self.encapsulation_coeff += self.encaps_precent      # added as an absolute amount (F7)
                if self.save_events:
                    self.temporal_events["encapsulation_drift"].append([self.reset_count, self.encaps_precent])
                if self.verbose:
                    print(f"[reset {self.reset_count}] encapsulation: conduct_modifier is now {self.encapsulation_coeff}")
            if self.plasticity_episode == self.reset_count:
                self.plasticity_episode += self.calc_next_temp_event(p["plasticity_drift_freq"], [0, 1])
                self.w0_without_locus = self.w0_process[self.plasticity_process_count]
                self.plasticity_process_count += 1
                if self.save_events:
                    self.temporal_events["plasticity_drift"].append([self.reset_count, self.w0_without_locus])
                if self.verbose:
                    print(f"[reset {self.reset_count}] plasticity walk advanced to entry {self.plasticity_process_count}")
            if self.reset_count % self.reset_plasticity_episode == 0:
                if self.verbose:
                    print(f"[reset {self.reset_count}] plasticity walk regenerated")
                self.plasticity_process_count = 0
                self.w0_without_locus = copy.deepcopy(self.w0_without_locus_)
                self.w0_process = generate_perturbations(self.w0_without_locus,
                                                         M=self.reset_plasticity_episode * 2,
                                                         step_scale=self.plasticity_percent * 0.01)
        if p["spatial_feature"]:
            if self.spatial_var_episode == self.reset_count and self.reset_count > 2:
                table = stim_rec_table()
                pick = np.random.choice(len(table))
                self.elec_coords = [table[pick][0]]
                self.rec_coords = [table[pick][1]]
                self.spatial_var_episode += self.spatial_var_freq
                self.spatial_events.append([self.reset_count, table[pick]])
                if self.verbose:
                    print(f"[reset {self.reset_count}] new stimulation / recording contacts: {table[pick]}")
        if p["save_events"] and p["log_path"] is not None and self.reset_count > 1:
            np.save(os.path.join(p["log_path"], f"temp_{self.reset_count}.npy"), self.temporal_events)

        self.w0 = apply_locus_mask(self.w0_without_locus, p["locus_without_w0"], p["locus_mask"])
        w0_raw = None if p["temporal_drift"] else np.array(self.w0, dtype=np.float64)
        # KuramotoJAX.__init__ (env.py:211-243)
        self.w0 = remove_negative_w0(self.w0)
        assert np.min(self.w0) >= 0, "Natural frequencies w0 must be positive!"
        electrode = cached_electrode(p, self.encapsulation_coeff, self.elec_coords, self.rec_coords,
                                     self.verbose)
        if not self.save_init or self.init_state is None:
            self.init_state = np.random.normal(loc=p["init_state_mean"], scale=p["init_state_sd"],
                                               size=(p["num_oscillators"]))
            self.init_state = remove_negative_w0(self.init_state)
        w0_out = np.asarray(self.w0, dtype=np.float64)
        # without drift, w0 before the non-positive fix and the electrode never change -> reusable by the fast path
        self._fast_cache = None if w0_raw is None else (w0_raw, np.flatnonzero(w0_raw <= 0.), electrode)
        return EpisodeSetup(w0=w0_out, stim=electrode.stim_vector(), rec=electrode.rec_vector(),
                            y0=np.asarray(self.init_state, dtype=np.float64), electrode=electrode)
