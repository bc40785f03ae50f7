#!/usr/bin/env python
"""Where does an auto-reset step of the env2 (drift) configuration spend its time? (run on the GPU box)"""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_params
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dicts = build_params(B, seed0=10, cfg_name="env2")
for d in dicts:
    d["total_episode_len"] = 90
venv = BatchedKuramotoVecEnv(dicts, compat_env2=True)
venv.reset()
L = int(venv.core.hosts[0].total_episode_counts); print("episode length", L)
rng = np.random.default_rng(0)
acts = rng.uniform(-1, 1, (L, B, 1)).astype(np.float32)
for ep in range(2):
    t0 = time.perf_counter()
    for k in range(L - 1):
        venv.step(acts[k])
    t_steps = time.perf_counter() - t0
    t0 = time.perf_counter(); venv.step(acts[L - 1]); t_reset = time.perf_counter() - t0
    print(f"episode {ep}: {L - 1} plain steps {1e3 * t_steps / (L - 1):.3f} ms each, auto-reset step {1e3 * t_reset:.1f} ms")
for k in range(L - 1):
    venv.step(acts[k])
venv.core.engine.set_timing(True)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); obs, rew, done, infos = venv.step(acts[L - 1]); dt = time.perf_counter() - t0
pr.disable()
print("auto-reset step s", dt, done.all())
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
