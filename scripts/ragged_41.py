#!/usr/bin/env python
"""BASELINE configs[4], the N = 65536 point as SURVEY.md 8d states it: the first 65536 rows of a 41 x 41 x 41 grid (odd extents,
38 full z-planes and 1658 neurons of the 39th) -- no grid structure the GRID kernels could use and a matrix (34 GB) that is
never formed: eigenpairs from the coordinates (geometry.lowrank_factors_points, on the GPU through torch), DENSE handle in
low-rank cluster mode (16 CTAs per environment).  B N = 2 097 152 as in the sweep: 32 environments."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dbsgym_b200.engine import KuramotoEngine  # noqa: E402
from dbsgym_b200.geometry import distances_from, lowrank_factors_points, neuron_grid  # noqa: E402
from dbsgym_b200.schedule import StepSchedule, transient_grid  # noqa: E402

G, N = 41, int(sys.argv[1]) if len(sys.argv) > 1 else 65536
B = max(2097152 // N, 8)
coords, grid = neuron_grid(G, G, G, N, 0.1)
t0 = time.perf_counter()
f = lowrank_factors_points(coords, "cos", tol=1e-9)
t_f = time.perf_counter() - t0
eng = KuramotoEngine(B, N, [G, G, G], 2340, 0.52, precision="f32", lowrank=f)
tt = transient_grid(200.0, 0.05)
eng.set_schedule(StepSchedule(400, tt[-1], 0.15, 0.75, 0.05)); eng.set_reward("bbpow_action", 0.05)
rng = np.random.default_rng(0)
stim = np.tile(np.maximum(0.0, 1.0 - distances_from(coords, [N // 2])[0]), (B, 1))
eng.set_env_params(None, w0=np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02, stim=stim, rec=stim, y0=rng.normal(np.pi, 0.6, (B, N)))
eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
act = torch.from_numpy(rng.uniform(-1, 1, (12, B)).astype(np.float32)).cuda()
st = torch.cuda.current_stream().cuda_stream
eng.set_timing(True)
ms = []
for i in range(10):
    eng.step_device(act[i].data_ptr(), None, None, None, st)
    torch.cuda.synchronize()
    if i >= 2:
        ms.append(eng.last_step_ms()[0])
c = eng.counters()
print(json.dumps({"N": N, "grid": "first N rows of 41 x 41 x 41 (ragged)", "envs": B, "variant": eng.step_variant(), "rank": int(f[0].shape[0]),
                  "residual_over_lambda_max": float(f[2] / abs(f[1][0])), "factorisation_s": t_f, "step_kernel_ms": float(np.mean(ms)),
                  "env_steps_per_s": B / np.mean(ms) * 1e3, "status": c["status"], "ctas_per_env": 16 if N > 32768 else None}))
