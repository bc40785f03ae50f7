#!/usr/bin/env python
"""A/B timing of the coupling evaluations on the cubic grids of BASELINE configs[4] (first N rows of 16^3 / 32^3) and on a
shuffled 8 x 8 x 8 grid: structured exact kernel (GRID), full matrix (DENSE) and the low-rank form (variant 11)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dbsgym_b200.engine import KuramotoEngine  # noqa: E402
from dbsgym_b200.geometry import coupling_rows, coupling_table, distances_from, lowrank_factors, neuron_grid  # noqa: E402
from dbsgym_b200.schedule import StepSchedule, transient_grid  # noqa: E402

cases = [(512, 8, True), (1024, 16, False), (2048, 16, False), (4096, 16, False)]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1].split(",")]
for N, G, shuffle in cases:
    B = max(2097152 // N, 296)
    coords, grid = neuron_grid(G, G, G, N, 0.1)
    if shuffle:
        perm = np.random.default_rng(1).permutation(N)
        coords, grid = coords[perm], grid[perm]
    table = None if shuffle else coupling_table(coords, grid, [G, G, G], "cos")
    alpha = coupling_rows(coords, np.arange(N), "cos")
    t0 = time.time()
    lr = lowrank_factors(alpha, tol=1e-9)
    t_eig = time.time() - t0
    rng = np.random.default_rng(N)
    stim = np.tile(np.maximum(0.0, 1.0 - distances_from(coords, [N // 2])[0]), (B, 1))
    w0 = np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02
    y0 = rng.normal(np.pi, 0.6, (B, N))
    tt = transient_grid(50.0, 0.05)
    ref = None
    engines = ([("grid", dict(coupling_table=table))] if table is not None else []) + [("lowrank", dict(lowrank=lr)), ("dense", dict(alpha=alpha))]
    for name, kw in engines:
        eng = KuramotoEngine(B, N, [G, G, G], 2340, 0.52, precision="f32", **kw)
        eng.set_schedule(StepSchedule(400, tt[-1], 0.15, 0.75, 0.05)); eng.set_reward("bbpow_action", 0.05)
        eng.set_env_params(None, w0=w0, stim=stim, rec=stim, y0=y0)
        eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
        act = torch.from_numpy(np.random.default_rng(7).uniform(-1, 1, (12, B)).astype(np.float32)).cuda()
        st = torch.cuda.current_stream().cuda_stream
        eng.set_timing(True)
        n_steps = 4 if name == "dense" else 10
        ms = []
        for i in range(n_steps):
            eng.step_device(act[i].data_ptr(), None, None, None, st)
            torch.cuda.synchronize()
            if i >= 2:
                ms.append(eng.last_step_ms()[0])
            if i == 3:
                y = eng.state()[:8]                      # compared after 4 steps
        c = eng.counters()
        line = (f"N {N:5d} B {B:5d} {'shuffled ' if shuffle else ''}{name:8s} variant {eng.step_variant():2d} "
                f"step {np.mean(ms):8.3f} ms -> {B / np.mean(ms) * 1e3:10.0f} env-steps/s, status {c['status']}")
        if name == "lowrank":
            line += f", rank {lr[0].shape[0]}, residual {lr[2]:.1e}, eig {t_eig:.1f}s"
        if ref is not None:
            line += f", max |phase - {engines[0][0]}| {np.max(np.abs(y - ref)):.1e}"
        elif ref is None:
            ref = y
        print(line, flush=True)
        eng.close()
