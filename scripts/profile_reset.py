import cProfile, pstats, sys, time, os
import numpy as np
sys.path.insert(0, '/root/repo')
from bench import build_params
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
B = 4096
dicts = build_params(B)
for d in dicts: d["total_episode_len"] = 4.5      # short episodes
venv = BatchedKuramotoVecEnv(dicts)
venv.reset()
L = int(venv.core.hosts[0].total_episode_counts); print("L", L)
a = np.zeros((B, 1), np.float32)
# first episode (slow-path reset at its end fills the fast cache), then profile the second auto-reset
for ep in range(2):
    for k in range(L - 1): venv.step(a)
    if ep == 0:
        t0 = time.perf_counter(); venv.step(a); print("first auto-reset step s", time.perf_counter() - t0)
pr = cProfile.Profile(); pr.enable()
t0 = time.perf_counter(); obs, rew, done, infos = venv.step(a); dt = time.perf_counter() - t0
pr.disable()
print("auto-reset step s", dt, done.all())
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
