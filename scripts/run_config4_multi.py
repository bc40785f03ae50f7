#!/usr/bin/env python
"""BASELINE configs[3]: env2 (temporal drift + electrode encapsulation / movement perturbations), 16384 environments
sharded over the GPUs of one box -- 16384 / world per rank, no collective on the step path, NCCL only to gather the
per-environment episode statistics at the end.  Launch like bench.py:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      scripts/run_config4_multi.py

Short episodes (100 steps) so that resets and drift events fall inside the timed window.
env2 as shipped cannot run: the two fixes are behind compat_env2=True."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
json_out = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)
import torch
import torch.distributed as dist
from bench import build_params
from dbsgym_b200.sharding import gather_episode_stats, shard_bounds
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
TOTAL = int(os.environ.get("CFG4_ENVS", 16384)); STEPS = int(os.environ.get("CFG4_STEPS", 300))
lo, hi = shard_bounds(TOTAL, rank, world)
B = hi - lo
dicts = build_params(B, seed0=10 + lo, cfg_name="env2")
for d in dicts:
    d["total_episode_len"] = 90                      # 100-step episodes
venv = BatchedKuramotoVecEnv(dicts, device=local_rank, compat_env2=True)
rng = np.random.default_rng(100 + rank)
obs = venv.reset()
for k in range(int(os.environ.get("CFG4_WARM", 110))):      # one full 100-step episode: the reset path (electrode cache) is warm
    obs, rew, done, _ = venv.step(rng.uniform(-1, 1, (B, 1)).astype(np.float32))
torch.cuda.synchronize()
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
rets = np.zeros(B); ndone = 0
t0 = time.perf_counter()
for k in range(STEPS):
    obs, rew, done, infos = venv.step(rng.uniform(-1, 1, (B, 1)).astype(np.float32))
    rets += rew; ndone += int(done.sum())
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.barrier()
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
c = venv.core.engine.counters()
stats = gather_episode_stats(np.stack([rets / STEPS, np.full(B, ndone / B)], axis=1), TOTAL, rank, world, device=dev)
if rank == 0:
    t = float(dt.cpu()[0])
    h0 = venv.core.hosts[0]
    json_out.write(json.dumps({
        "config": "env2 drift + encapsulation/movement perturbations, 100-step episodes (resets inside the timed window)",
        "n_gpus": world, "global_envs": TOTAL, "envs_per_gpu": B, "steps": STEPS,
        "env_steps_per_s": TOTAL * STEPS / t, "ms_per_batched_step": 1e3 * t / STEPS,
        "timing": "wall clock around the VecEnv loop (host resets included), barrier on both sides, max over ranks",
        "mean_reward": float(stats[:, 0].mean()), "episodes_finished_per_env": float(stats[:, 1].mean()),
        "solver_status_rank0": c["status"], "resets_env0": h0.reset_count, "elec_coords_env0": str(h0.elec_coords),
        "encapsulation_env0": h0.encapsulation_coeff}) + "\n")
    json_out.flush()
venv.close()
if world > 1:
    dist.destroy_process_group()
