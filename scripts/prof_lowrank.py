#!/usr/bin/env python
"""A few steps of the sector low-rank kernel on a cubic grid (default 32 x 32 x 16, N = 16384) for ncu:
    ncu --set full --import-source on -k regex:step_kernel -s 3 -c 1 -o gpurun_out/x python scripts/prof_lowrank.py [gz [gx = gy]]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from dbsgym_b200.engine import KuramotoEngine  # noqa: E402
from dbsgym_b200.geometry import coupling_table, distances_from, grid_sector_factors, neuron_grid  # noqa: E402
from dbsgym_b200.schedule import StepSchedule, transient_grid  # noqa: E402

gx = gy = int(sys.argv[2]) if len(sys.argv) > 2 else 32           # (16 4 / 16 8 / 16 16: the register-resident kernel, variant 13)
gz = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = gx * gy * gz
B = 2097152 // N
coords, grid = neuron_grid(gx, gy, gz, N, 0.1)
table = coupling_table(coords, grid, [gx, gy, gz], "cos")
eng = KuramotoEngine(B, N, [gx, gy, gz], 2340, 0.52, precision="f32", coupling_table=table)
eng.set_coupling_lowrank_sectors(*grid_sector_factors(table, gx, gy, gz, tol=1e-9))
tt = transient_grid(200.0, 0.05)
eng.set_schedule(StepSchedule(400, tt[-1], 0.15, 0.75, 0.05)); eng.set_reward("bbpow_action", 0.05)
rng = np.random.default_rng(0)
stim = np.tile(np.maximum(0.0, 1.0 - distances_from(grid * 0.1, [N // 2])[0]), (B, 1))
eng.set_env_params(None, w0=np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02, stim=stim, rec=stim, y0=rng.normal(np.pi, 0.6, (B, N)))
eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
act = torch.from_numpy(rng.uniform(-1, 1, (6, B)).astype(np.float32)).cuda()
for i in range(6):
    eng.step_device(act[i].data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(eng.step_variant(), eng.counters())
