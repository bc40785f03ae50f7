#!/usr/bin/env python
"""A few device-resident steps of the benchmark configuration (env1, N = 512, 4096 environments, float32) -- the
workload the ncu captures in profiles/ are taken on:
    ncu --set full --clock-control none --import-source on -k regex:warp_step_kernel -s 5 -c 1 -o gpurun_out/x python scripts/prof_step.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from bench import build_params  # noqa: E402
from dbsgym_b200.batched import BatchedKuramoto  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
core = BatchedKuramoto(build_params(B))
eng = core.engine
eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
acts = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (8, B)).astype(np.float32)).cuda()
for k in range(8):
    eng.step_device(acts[k].data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(eng.step_variant(), eng.counters())
