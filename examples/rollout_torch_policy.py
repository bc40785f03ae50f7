#!/usr/bin/env python
"""GPU-resident rollout collection: a torch policy reads the observation tensor the environment kernel
wrote, the environment reads the action tensor the policy wrote -- nothing crosses PCIe.

    python examples/rollout_torch_policy.py --envs 4096 --steps 128

This is the data-collection half of what aDBS_RL/train_aDBS_RL.py does with SB3's PPO (n_steps = 128,
MlpPolicy), written against ``BatchedKuramoto.step_tensor``.
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_params                      # noqa: E402
from dbsgym_b200.batched import BatchedKuramoto     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=128)
    ap.add_argument("--cfg", default="env1")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    core = BatchedKuramoto(build_params(args.envs, cfg_name=args.cfg), transfer="full")
    core.engine.set_episode(None, step_idx=0, episode_len=2 ** 30)
    B, W = core.num_envs, core.window
    policy = torch.nn.Sequential(torch.nn.Linear(W, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 1), torch.nn.Tanh()).to(dev)
    obs_buf = torch.empty((args.steps + 1, B, W), dtype=torch.float32, device=dev)      # rollout storage
    act_buf = torch.empty((args.steps, B), dtype=torch.float32, device=dev)
    rew_buf = torch.empty((args.steps, B), dtype=torch.float32, device=dev)
    done = torch.empty(B, dtype=torch.uint8, device=dev)
    obs_buf[0].copy_(torch.from_numpy(core.observations()).to(dev))
    with torch.no_grad():                                      # warm-up: cuBLAS handles, lazy module loading
        for _ in range(10):
            a = policy(obs_buf[0]).squeeze(-1)
            core.step_tensor(a, obs=obs_buf[0], reward=rew_buf[0], done=done)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for k in range(args.steps):
            act_buf[k] = policy(obs_buf[k]).squeeze(-1)
            core.step_tensor(act_buf[k], obs=obs_buf[k + 1], reward=rew_buf[k], done=done)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{B} envs x {args.steps} steps with an MLP policy in the loop: {B * args.steps / dt:,.0f} env-steps/s, "
          f"mean reward {rew_buf.mean().item():.3f}")
    core.close()


if __name__ == "__main__":
    main()
