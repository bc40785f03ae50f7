#!/usr/bin/env python
"""Soak: 4096 env1 environments through TWO full training episodes (5555 steps each, auto-reset in between) via the public
VecEnv API with uniform random actions; checks the solver status word, finiteness and the episode bookkeeping."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_params
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv

B = int(os.environ.get("SOAK_ENVS", 4096))
venv = BatchedKuramotoVecEnv(build_params(B))
obs = venv.reset()
L = int(venv.core.hosts[0].total_episode_counts)
rng = np.random.default_rng(0)
ret = np.zeros(B); ndone = 0; rmin, rmax = np.inf, -np.inf
t0 = time.perf_counter()
for k in range(2 * L + 10):
    obs, rew, done, infos = venv.step(rng.uniform(-1, 1, (B, 1)).astype(np.float32))
    assert np.all(np.isfinite(rew)) and np.all(np.isfinite(obs[::257]))
    ret += rew; ndone += int(done.sum()); rmin = min(rmin, float(rew.min())); rmax = max(rmax, float(rew.max()))
    if done.any():
        assert done.all() and "episode" in infos[0] and infos[0]["episode"]["l"] == L
dt = time.perf_counter() - t0
c = venv.core.engine.counters()
print(json.dumps({"envs": B, "episode_len": L, "steps": 2 * L + 10, "episodes_finished": ndone, "wall_s": dt,
                  "env_steps_per_s_incl_resets": B * (2 * L + 10) / dt, "solver_status": c["status"],
                  "rejected_fraction": c["rejected"] / max(1, c["accepted"] + c["rejected"]),
                  "reward_min": rmin, "reward_max": rmax, "mean_return_per_step": float(ret.mean() / (2 * L + 10))}))
assert c["status"] == 0 and ndone == 2 * B
venv.close()
