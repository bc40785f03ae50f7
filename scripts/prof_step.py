#!/usr/bin/env python
"""A few device-resident steps of the benchmark configuration (env1, N = 512, 4096 environments, float32) -- the
workload the ncu captures in profiles/ are taken on:
    ncu --set full --clock-control none --import-source on -k regex:warp_step_kernel -s 5 -c 1 -o gpurun_out/x python scripts/prof_step.py
(``prof_step.py 8192 256``: the half grid, kernel warp1_step_kernel)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import build_params  # noqa: E402
from dbsgym_b200.batched import BatchedKuramoto  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 512            # 256: the half grid of configs[4] (step-kernel variant 12)
dicts = build_params(B)
if N == 256:
    from dbsgym_b200 import utils
    np.random.seed(3)
    w0, nc, ng, w0t, wl, lm = utils.generate_w0_with_locus(256, [8, 8, 8], 0.1, [2, 4, 4], 0.55, 17, 1, show=False)
    for e, d in enumerate(dicts):
        d.update(num_oscillators=256, w0=w0.copy(), w0_without_locus=w0t.copy(), locus_without_w0=wl, locus_mask=lm,
                 neur_coords=nc, neur_grid=ng, elec_coords=[[2, 3, 4]], rand_seed=100 + e)
core = BatchedKuramoto(dicts)
eng = core.engine
eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
acts = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (8, B)).astype(np.float32)).cuda()
for k in range(8):
    eng.step_device(acts[k].data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(eng.step_variant(), eng.counters())
