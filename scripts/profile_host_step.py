#!/usr/bin/env python
"""Where does the host-side time of BatchedKuramotoVecEnv.step go? (run on the GPU box)"""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_params
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv

B = 4096
venv = BatchedKuramotoVecEnv(build_params(B))
venv.reset()
venv.core.engine.set_episode(None, step_idx=0, episode_len=2 ** 30)
acts = np.random.default_rng(0).uniform(-1, 1, (300, B, 1)).astype(np.float32)
for i in range(20):
    venv.step(acts[i])
t0 = time.perf_counter()
for i in range(20, 120):
    venv.step(acts[i])
print("ms per VecEnv.step:", (time.perf_counter() - t0) * 10)
core = venv.core
t0 = time.perf_counter()
for i in range(100):
    core.engine.step_host_mirror(core.act_buf, core.rew_buf, core.done_buf)
print("ms per raw dbsgym_step_host_mirror:", (time.perf_counter() - t0) * 10)
core.engine.set_timing(True)
core.engine.step_host_mirror(core.act_buf, core.rew_buf, core.done_buf)
print("kernel ms (step, obs):", core.engine.last_step_ms())
core.engine.set_timing(False)
t0 = time.perf_counter()
for i in range(100):
    core.engine.step_host_samples(core.act_buf, core.samples_buf, core.nsamp_buf, core.rew_buf, core.done_buf)
print("ms per raw dbsgym_step_host_samples:", (time.perf_counter() - t0) * 10)
pr = cProfile.Profile(); pr.enable()
for i in range(120, 220):
    venv.step(acts[i])
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
