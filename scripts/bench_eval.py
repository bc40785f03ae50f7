#!/usr/bin/env python
"""Evaluation metric at BASELINE size: 4096 env0 environments, one 1111-step evaluation episode under HF-DBS with the
TRUE-LFP trace recorded on the device, then calc_psd_for_simple_eval (aDBS_RL/evaluate_HF_DBS.py:122-135) on the
device (csrc/eval_kernel.cuh) and -- on a sample of the environments -- with scipy on the host."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import build_params
from dbsgym_b200.evaluation import calc_psd_for_simple_eval, device_bbpow
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv

B = int(os.environ.get("EVAL_ENVS", 4096)); STEPS = int(os.environ.get("EVAL_STEPS", 1111))
dicts = build_params(B, cfg_name="env0")
venv = BatchedKuramotoVecEnv(dicts)
venv.reset()
eng = venv.core.engine
eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
dev = torch.device("cuda", 0)
act = torch.ones(B, dtype=torch.float32, device=dev)
rew = torch.empty(B, dtype=torch.float32, device=dev); done = torch.empty(B, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
eng.trace_begin(STEPS * eng.max_step_samples)
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(STEPS):
    eng.step_device(act.data_ptr(), None, rew.data_ptr(), done.data_ptr(), st)
torch.cuda.synchronize(); t_episode = time.perf_counter() - t0
eng.trace_end()
device_bbpow(eng)                                                    # warm-up (allocations, weight construction)
t0 = time.perf_counter(); bb = device_bbpow(eng); t_dev = time.perf_counter() - t0
tr, ln = eng.trace()
n = int(ln[0]); sample = min(B, 256)
t0 = time.perf_counter(); ref = calc_psd_for_simple_eval(tr[:sample, :n], 0.0005); t_host = (time.perf_counter() - t0) * B / sample
print(json.dumps({"envs": B, "steps": STEPS, "trace_samples": n, "episode_s_device_resident": t_episode,
                  "eval_device_s": t_dev, "eval_host_scipy_s_extrapolated": t_host, "host_sample": sample,
                  "max_rel_diff_vs_scipy": float(np.max(np.abs(bb[:sample] - ref) / np.abs(ref))),
                  "bbpow_mean": float(bb.mean()), "bbpow_sd": float(bb.std(ddof=1)),
                  "paper_hf_dbs_row": "2.34e-3 +- 0.2e-3 (data/kur-table-metrics.xlsx)"}))
venv.close()
