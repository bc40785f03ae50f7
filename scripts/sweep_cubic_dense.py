#!/usr/bin/env python
"""SURVEY.md 8d config 5 asks for cubic grids (16^3 for N = 1024 ... 4096).  The resident-state GRID kernels need
lines of 8 along y, so cubic grids with 16 neurons per line run on the generic DENSE path (alpha^T streamed from L2,
any symmetric alpha, N <= 8192).  This script measures that fallback on the first N rows of the 16^3 grid
(utils.py:478-497 takes the first n_neurons rows) for the record; the structured sweep is scripts/sweep_n.py."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dbsgym_b200.engine import KuramotoEngine
from dbsgym_b200.geometry import coupling_rows, distances_from, neuron_grid
from dbsgym_b200.schedule import StepSchedule, transient_grid

for N, B in ((1024, 512), (2048, 256), (4096, 128)):
    coords, grid = neuron_grid(16, 16, 16, N, 0.1)
    alpha = coupling_rows(coords, np.arange(N), "cos", 1.0, 1.0)
    eng = KuramotoEngine(B, N, [16, 16, 16], 2340, 0.52, precision="f32", alpha=alpha)
    tt = transient_grid(200.0, 0.05)
    eng.set_schedule(StepSchedule(400, tt[-1], 0.15, 0.75, 0.05)); eng.set_reward("bbpow_action", 0.05)
    rng = np.random.default_rng(N)
    centre = int(np.argmin(np.abs(grid - np.array([8, 7, grid[:, 2].max() // 2])).sum(axis=1)))
    stim = np.tile(np.maximum(0.0, 1.0 - distances_from(grid * 0.1, [centre])[0]), (B, 1))
    eng.set_env_params(None, w0=np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02, stim=stim, rec=stim,
                       y0=rng.normal(np.pi, 0.6, (B, N)))
    eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
    dev = torch.device("cuda", 0)
    act = torch.from_numpy(rng.uniform(-1, 1, (12, B)).astype(np.float32)).to(dev)
    rew = torch.empty(B, dtype=torch.float32, device=dev); done = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    eng.set_timing(True)
    for i in range(2):
        eng.step_device(act[i].data_ptr(), None, rew.data_ptr(), done.data_ptr(), st)
    torch.cuda.synchronize(); eng.counters(reset=True)
    ms = []
    for i in range(2, 8):
        eng.step_device(act[i].data_ptr(), None, rew.data_ptr(), done.data_ptr(), st)
        torch.cuda.synchronize(); ms.append(eng.last_step_ms()[0])
    c = eng.counters(); k = float(np.mean(ms))
    rhs = (c["rhs_evals"] - eng.rhs_reused()) / (6 * B)
    print(json.dumps({"N": N, "grid": "first N rows of 16x16x16", "coupling": "dense (alpha^T streamed from L2)", "envs": B,
                      "step_kernel_ms": k, "env_steps_per_s": B / (k * 1e-3), "rhs_executed_per_env_step": rhs,
                      "executed_tflops": rhs * 4 * N * N * B / (k * 1e-3) / 1e12, "status": c["status"],
                      "variant": eng.step_variant()}), flush=True)
    eng.close()
