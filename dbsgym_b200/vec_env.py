"""``BatchedKuramotoVecEnv`` -- a Stable-Baselines3 ``VecEnv`` over one GPU batch of environments.

Semantics are those of ``DummyVecEnv([lambda: Monitor(SpatialKuramoto(d)) for d in dicts])``
(what the reference's train / eval scripts build, aDBS_RL/train_aDBS_RL.py:48-56,181 and
evaluate_HF_DBS.py:53-54): observations ``[B, 1, W]`` float32, auto-reset inside ``step_wait`` with
``infos[i]['terminal_observation']``, ``infos[i]['TimeLimit.truncated'] = False`` and Monitor's
``infos[i]['episode'] = {'r', 'l', 't'}`` on the terminal step, ``get_attr`` for the names the
reference's callbacks read (``'params_dict'``, ``'u'``, ``'theta_mean'``; custom_callbacks.py:107-134).
It subclasses SB3's own ``VecEnv`` when SB3 is importable.
"""
from __future__ import annotations

import copy
import time

import numpy as np

from ._compat import Box, VecEnvBase
from .batched import BatchedKuramoto
from .env import _default_precision


class _LazyAttrList:
    """Sequence returned by ``get_attr``: element i is fetched only when indexed, so the
    callbacks' ``get_attr('theta_mean')[0]`` does not pull every environment off the GPU."""

    def __init__(self, getter, indices):
        self._getter, self._indices = getter, list(indices)

    def __len__(self):
        return len(self._indices)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._getter(i) for i in self._indices[k]]
        return self._getter(self._indices[k])

    def __iter__(self):
        return (self._getter(i) for i in self._indices)


class _InfoList(list):
    """``infos`` of a step: one dict per environment like DummyVecEnv returns, but the dicts are created only when
    somebody indexes them (building thousands of dicts -- and, at an episode end, thousands of ``terminal_observation``
    copies and Monitor records -- per step costs more than the GPU step).  Every environment gets its OWN dict, so a
    wrapper that mutates ``infos[i]`` touches nothing else."""

    def __init__(self, n):
        super().__init__([None] * n)
        self._terminal = None          # (done mask, terminal observations [B, 1, W], episode returns, lengths, wall time)

    def set_terminal(self, done, term_obs, ep_ret, ep_len, t, monitor):
        self._terminal = (done, term_obs, ep_ret, ep_len, t, monitor)

    def _fill(self, i):
        d = list.__getitem__(self, i)
        if d is None:
            d = {"TimeLimit.truncated": False}
            if self._terminal is not None and self._terminal[0][i]:
                done, term_obs, ep_ret, ep_len, t, monitor = self._terminal
                d["terminal_observation"] = term_obs[i]
                if monitor:
                    d["episode"] = {"r": round(float(ep_ret[i]), 6), "l": int(ep_len[i]), "t": round(t, 6)}
            list.__setitem__(self, i, d)
        return d

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._fill(i) for i in range(*k.indices(len(self)))]
        return self._fill(k if k >= 0 else k + len(self))

    def __iter__(self):
        return (self._fill(i) for i in range(len(self)))


class BatchedKuramotoVecEnv(VecEnvBase):
    """``copy_obs=False`` (default): ``step_wait`` returns a zero-copy view of the pinned host log the GPU appends the
    window samples to.  Later steps only append behind it and a reset switches to a second buffer, so the array
    stays intact for at least one further ``step()`` -- what SB3 needs: ``collect_rollouts`` stores
    ``self._last_obs`` right after the NEXT ``env.step()`` returns -- and in practice for 13 steps.  Keep an
    observation longer than that only as a copy, or pass ``copy_obs=True`` for arrays you own."""

    def __init__(self, params_dicts, num_envs=None, precision=None, device=0, compat_env2=False,
                 monitor=True, transfer="delta", copy_obs=False, engine_options=None, coupling_eval=None):
        if isinstance(params_dicts, dict):
            n = int(num_envs or params_dicts.get("num_envs", 1))
            # like DummyVecEnv([make_env(d)] * n): every env gets its own copy of the dict
            params_dicts = [copy.deepcopy(params_dicts) for _ in range(n)]
        p0 = params_dicts[0]
        self.core = BatchedKuramoto(params_dicts, precision=precision or _default_precision(p0),
                                    device=device, compat_env2=compat_env2, transfer=transfer,
                                    engine_options=engine_options,
                                    coupling_eval=coupling_eval or p0.get("coupling_eval", "auto"))
        self.copy_obs = bool(copy_obs)
        B, W = self.core.num_envs, self.core.window
        obs_space = Box(low=-1.5, high=1.5, shape=(1, W), dtype=np.float32)
        act_space = Box(low=-1., high=1., shape=(1,), dtype=np.float32)
        super().__init__(B, obs_space, act_space)
        self.monitor = monitor
        self._actions = None
        self._t0 = time.time()
        self._ep_ret = np.zeros(B, dtype=np.float64)
        self._ep_len = np.zeros(B, dtype=np.int64)
        self._extra = [dict() for _ in range(B)]      # set_attr storage

    # ---- VecEnv API --------------------------------------------------------------------------
    def reset(self):
        self.core.reset_envs(range(self.num_envs))
        self._ep_ret[:] = 0
        self._ep_len[:] = 0
        self.reset_infos = [{} for _ in range(self.num_envs)]
        self._reset_seeds()
        self._reset_options()
        return self.core.observations().reshape(self.num_envs, 1, -1).copy()

    def step_async(self, actions):
        """Really asynchronous: the step kernel is launched here; the bookkeeping that does not depend on its
        results runs while the GPU works, step_wait() only waits and reads the results."""
        self.core.step_begin(np.asarray(actions, dtype=np.float32).reshape(self.num_envs, -1)[:, 0])
        self._ep_len += 1
        self._infos = _InfoList(self.num_envs)        # per-environment dicts, created when indexed

    def step_wait(self):
        obs, rew, done = self.core.step_end()
        obs = obs.reshape(self.num_envs, 1, -1)       # a view (host window mirror / pinned buffer): no 38 MB copy
        if self.copy_obs:
            obs = obs.copy()
        rew = rew.copy()
        done = done.copy()
        self._ep_ret += rew                           # float64 accumulation of the float32 rewards SB3 sees (Monitor does the same)
        infos = self._infos
        finished = np.flatnonzero(done)
        if finished.size:
            # one copy of the finished environments' last observations (they are about to be replaced by the reset windows);
            # the per-environment info dicts (terminal_observation, Monitor's episode record) are built when indexed
            term = np.empty_like(obs) if finished.size == self.num_envs else None
            if term is not None:
                np.copyto(term, obs)
            else:
                term = {int(i): obs[i].copy() for i in finished}
            infos.set_terminal(done, term, self._ep_ret.copy(), self._ep_len.copy(), time.time() - self._t0, self.monitor)
            self.core.reset_envs(finished)            # index order == sequential DummyVecEnv order
            fresh = self.core.observations_after_reset()     # owned array / view of the NEW mirror buffer
            obs = fresh.reshape(self.num_envs, 1, -1)
            if self.copy_obs:
                obs = obs.copy()
            self._ep_ret[finished] = 0
            self._ep_len[finished] = 0
        return obs, rew, done, infos

    def close(self):
        self.core.close()

    def get_attr(self, attr_name, indices=None):
        idx = list(self._get_indices(indices))
        return _LazyAttrList(lambda i: self._attr(attr_name, i), idx)

    def set_attr(self, attr_name, value, indices=None):
        for i in self._get_indices(indices):
            self._extra[i][attr_name] = value

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        out = []
        for i in self._get_indices(indices):
            if method_name == "reset":
                self.core.reset_envs([i])
                out.append((self.core.observations()[i:i + 1].copy(), {}))
            elif method_name == "rescale_action":
                lo, hi = self.core.hosts[i].dbs_action_bounds
                out.append(lo + ((hi - lo) * (method_args[0] + 1.0)) / 2.0)
            else:
                raise AttributeError(f"env_method('{method_name}') is not available on the batched env")
        return out

    def env_is_wrapped(self, wrapper_class, indices=None):
        # the Monitor bookkeeping is built in: report "wrapped" only for a class literally called Monitor
        wrapped = self.monitor and getattr(wrapper_class, "__name__", "") == "Monitor"
        return [wrapped for _ in self._get_indices(indices)]

    def get_images(self):
        return [None for _ in range(self.num_envs)]

    def render(self, mode=None):
        return None

    # ---- per-env attribute view ----------------------------------------------------------------
    def _attr(self, name, i):
        if name in self._extra[i]:
            return self._extra[i][name]
        core, host = self.core, self.core.hosts[i]
        if name == "params_dict":
            return core.params_dicts[i]
        if name == "theta_mean":
            return core.theta_mean(i)
        if name == "theta_records":
            return core.theta_records(i)
        if name == "u":
            return core.u(i)
        if name == "current_step":
            return int(core.current_step[i])
        if name == "current_time":
            return core.current_time(i)
        if name == "theta_state":
            return core.engine.window_values([i])
        if name == "sol_state":
            return core.engine.state([i])
        if name == "render_mode":
            return None
        if hasattr(host, name):
            return getattr(host, name)
        raise AttributeError(name)
