import sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
from bench import build_params
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
import torch
B = 4096
venv = BatchedKuramotoVecEnv(build_params(B)); venv.reset()
core = venv.core; eng = core.engine
orig = eng.set_env_params
def timed(env_ids=None, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter(); orig(env_ids, **kw); torch.cuda.synchronize()
    print("  set_env_params", {k: (v.dtype, v.flags['C_CONTIGUOUS'], v.shape) for k, v in kw.items() if v is not None}, type(env_ids), getattr(env_ids, 'dtype', None), round((time.perf_counter() - t0) * 1e3, 1), "ms")
eng.set_env_params = timed
a = np.zeros((B, 1), np.float32)
for rep in range(4):
    for k in range(3): venv.step(a)
    torch.cuda.synchronize(); t0 = time.perf_counter(); core.reset_envs(range(B)); torch.cuda.synchronize()
    print("reset_envs", rep, round((time.perf_counter() - t0) * 1e3, 1), "ms")
