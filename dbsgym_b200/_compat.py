"""Resolve the third-party base classes lazily.

gymnasium and stable-baselines3 are what the reference's callers use (environment/env.py:7-8,
aDBS_RL/*.py), but neither is installed in the build image.  When they are importable the real
classes are used (SB3 checks ``isinstance(env, VecEnv)``, so duck typing is not enough); otherwise
minimal local mirrors with the same surface keep the package and its tests working.
"""
from __future__ import annotations

from abc import ABC, abstractmethod

import numpy as np

try:                                               # pragma: no cover - not installed in CI image
    import gymnasium as _gym
    from gymnasium.spaces import Box
    GymEnv = _gym.Env
    HAVE_GYMNASIUM = True
except Exception:                                  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:                                     # gymnasium.spaces.Box, the part env.py:310-315 uses
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            return np.random.uniform(self.low, self.high).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class GymEnv:
        metadata = {}
        render_mode = None
        spec = None

        def reset(self, *, seed=None, options=None):
            return None

        def close(self):
            pass

        @property
        def unwrapped(self):
            return self

try:                                               # pragma: no cover
    from stable_baselines3.common.vec_env import VecEnv as VecEnvBase
    HAVE_SB3 = True
except Exception:                                  # noqa: BLE001
    HAVE_SB3 = False

    class VecEnvBase(ABC):
        """Mirror of stable_baselines3.common.vec_env.VecEnv 2.6 (constructor + step plumbing)."""

        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs = num_envs
            self.observation_space = observation_space
            self.action_space = action_space
            self.reset_infos = [{} for _ in range(num_envs)]
            self._seeds = [None for _ in range(num_envs)]
            self._options = [{} for _ in range(num_envs)]
            self.render_mode = None

        @abstractmethod
        def reset(self): ...

        @abstractmethod
        def step_async(self, actions): ...

        @abstractmethod
        def step_wait(self): ...

        @abstractmethod
        def close(self): ...

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

        def seed(self, seed=None):
            if seed is None:
                seed = int(np.random.randint(0, np.iinfo(np.uint32).max, dtype=np.uint32))
            self._seeds = [seed + idx for idx in range(self.num_envs)]
            return self._seeds

        def set_options(self, options=None):
            if options is None:
                options = {}
            self._options = [options] * self.num_envs if isinstance(options, dict) else list(options)

        def _reset_seeds(self):
            self._seeds = [None for _ in range(self.num_envs)]

        def _reset_options(self):
            self._options = [{} for _ in range(self.num_envs)]

        def _get_indices(self, indices):
            if indices is None:
                return range(self.num_envs)
            if isinstance(indices, int):
                return [indices]
            return indices

        @property
        def unwrapped(self):
            return self
