"""Baseline controllers with the ``predict`` surface of reference aDBS_RL/agents/simple_dbs.py,
vectorised over the environments of a batched VecEnv.

* ``HFDBS``     constant action (simple_dbs.py:16-24); the reference returns ``[[a]]`` for one env.
* ``RandomDBS`` uniform(-m, m) per env (simple_dbs.py:27-40).
* ``PIDController`` (simple_dbs.py:43-95): error = -reward(window, last action); one PID state per env.
"""
from __future__ import annotations

import numpy as np

from . import utils


class HFDBS:
    def __init__(self, action: float):
        self.action = action

    def predict(self, observation, state=None, episode_start=None, deterministic=True):
        n = np.asarray(observation).shape[0]
        if n == 1:
            return [[self.action]], None
        return np.full((n, 1), self.action, dtype=np.float32), None


class RandomDBS:
    def __init__(self, action_magnitude: float):
        self.action_magnitude = action_magnitude
        assert self.action_magnitude > 0

    def predict(self, observation, state=None, episode_start=None, deterministic=True):
        n = np.asarray(observation).shape[0]
        a = np.random.uniform(-self.action_magnitude, self.action_magnitude, size=(n,)).astype(np.float32)
        return [a], None


class PIDController:
    """Per-environment PID on the reward-derived error.  ``env`` may be a SpatialKuramoto (its
    reward methods are used, as in the reference) or None (the built-in reward formulas are used)."""

    def __init__(self, Kp_init, Ki_init, Kd_init, dt, env=None, u_max=1., u_min=-1., reward="bbpow",
                 verbose_dt=0.05, dbs_action_bounds=(-5, 5)):
        if reward not in ("bbpow", "temp", "thr"):
            raise NotImplementedError()
        self.Kp, self.Ki, self.Kd, self.dt = Kp_init, Ki_init, Kd_init, dt
        self.u_max, self.u_min = u_max, u_min
        self.reward, self.env = reward, env
        self.verbose_dt = verbose_dt
        self.action = 0
        self.integral = 0
        self.prev_error = 1

    def _error(self, windows, actions):
        dt = utils.units2sec(self.verbose_dt)
        out = np.empty(len(windows))
        for i, (x, a) in enumerate(zip(windows, actions)):
            if self.reward == "bbpow":
                out[i] = 1e4 * utils.calc_beta_band_power(x, dt, 12.5, 21) + 1e-2 * np.abs(a)
            elif self.reward == "temp":
                f, _ = utils.band_pass_envelope(x, 1 / dt, order=2)
                out[i] = 1e3 * (f[-1] - np.mean(f)) ** 2 + 1e-2 * np.abs(a)
            else:
                bb = 1e4 * utils.calc_beta_band_power(x, dt, 12.5, 21)
                out[i] = (5. if bb > 20 else 0) + np.abs(float(a))
        return out

    def compute(self, error):
        self.integral = self.integral + error * self.dt
        derivative = (error - self.prev_error) / self.dt if self.dt != 0 else 0.0
        output = self.Kp * error + self.Ki * self.integral + self.Kd * derivative
        self.prev_error = error
        return np.clip(output, self.u_min, self.u_max)

    def predict(self, observation, state=None, episode_start=None, deterministic=True):
        obs = np.asarray(observation)
        n = obs.shape[0]
        windows = obs.reshape(n, -1)
        acts = np.broadcast_to(np.asarray(self.action, dtype=np.float64), (n,))
        e = self._error(windows, acts)
        self.action = self.compute(e if n > 1 else float(e[0]))
        a = np.broadcast_to(np.asarray(self.action, dtype=np.float32), (n,)).astype(np.float32)
        return [a], None


class BatchedPID:
    """PID controller for a whole batch, driven by the reward the GPU already computed.

    The reference's controller (simple_dbs.py:81-95) takes ``e = -reward_func(window, [a_prev])`` with the
    *normalised* previous action in the penalty term, while the environment's own reward uses the rescaled
    amplitude ``u = rescale(a_prev)``.  For the two beta-power rewards the band-power term is identical, so
    ``e = -reward_env - cost * (|u| - |a_prev|)`` and no FFT is needed on the host.  One PID state per env.
    """

    def __init__(self, Kp, Ki, Kd, dt, n_envs, u_max=1., u_min=-1., action_cost=1e-2, action_bounds=(-5., 5.)):
        self.Kp, self.Ki, self.Kd, self.dt = Kp, Ki, Kd, dt
        self.u_max, self.u_min = u_max, u_min
        self.cost, self.lo, self.hi = action_cost, action_bounds[0], action_bounds[1]
        self.action = np.zeros(n_envs)
        self.integral = np.zeros(n_envs)
        self.prev_error = np.ones(n_envs)

    def error_from_reward(self, reward):
        u = self.lo + (self.hi - self.lo) * (self.action + 1.0) / 2.0
        return -np.asarray(reward, dtype=np.float64) - self.cost * (np.abs(u) - np.abs(self.action))

    def act(self, error):
        self.integral = self.integral + error * self.dt
        derivative = (error - self.prev_error) / self.dt if self.dt != 0 else 0.0
        out = self.Kp * error + self.Ki * self.integral + self.Kd * derivative
        self.prev_error = error
        self.action = np.clip(out, self.u_min, self.u_max)
        return self.action.astype(np.float32)
