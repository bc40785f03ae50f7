#!/usr/bin/env python
"""Run the BASELINE.json configurations beyond the bench headline on one GPU and print a JSON summary
(throughput through the public VecEnv API with the controller in the loop + sanity statistics).

  configs[1]  env0, 4096 envs, HF-DBS and PID baseline controllers
  configs[2]  env1 with directed stimulation, 4096 envs, uniform random actions
  configs[3]  env2 (compat switch), 2048 envs (= the per-GPU share of 16384 / 8), short episodes
              so that resets + drift events fall inside the timed window
  configs[4]  smallest sweep point N = 256 on the GRID kernel, 8192 envs (B*N = 2 097 152)
"""
import copy, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import build_params
from dbsgym_b200 import utils
from dbsgym_b200.controllers import BatchedPID
from dbsgym_b200.vec_env import BatchedKuramotoVecEnv

out = {}
STEPS = int(os.environ.get("CFG_STEPS", 200))


def timed(venv, policy, steps, warm=5):
    obs = venv.reset()
    B = venv.num_envs
    rew = np.zeros(B, dtype=np.float32)
    rets, energy = np.zeros(B), np.zeros(B)
    for k in range(warm):
        obs, rew, done, _ = venv.step(policy(obs, rew, k))
    t0 = time.perf_counter()
    ndone = 0
    for k in range(steps):
        a = policy(obs, rew, k + warm)
        obs, rew, done, infos = venv.step(a)
        rets += rew; energy += np.abs(np.asarray(a).reshape(B))
        ndone += int(done.sum())
    dt = time.perf_counter() - t0
    c = venv.core.engine.counters()
    return {"env_steps_per_s": B * steps / dt, "ms_per_batched_step": 1e3 * dt / steps, "envs": B, "steps": steps,
            "mean_reward": float(rets.mean() / steps), "mean_abs_action": float(energy.mean() / steps),
            "episodes_finished": ndone, "solver_status": c["status"],
            "rejected_fraction": c["rejected"] / max(1, c["accepted"] + c["rejected"])}


# ---- configs[1]: env0, HF-DBS and PID -------------------------------------------------------------
B = 4096
dicts = build_params(B, cfg_name="env0")
venv = BatchedKuramotoVecEnv(dicts)
out["env0_hf_dbs_4096"] = timed(venv, lambda o, r, k: np.ones((B, 1), np.float32), STEPS)
pid = BatchedPID(7.3078, 3.7864, 5.1291, 0.05, B)          # "PID_R1" gains, evaluate_aDBS_RL_IQL.py:260
out["env0_pid_4096"] = timed(venv, lambda o, r, k: pid.act(pid.error_from_reward(r)).reshape(B, 1), STEPS)
out["env0_dbs_off_4096"] = timed(venv, lambda o, r, k: np.zeros((B, 1), np.float32), STEPS)
venv.close()

# ---- configs[2]: env1 with directed stimulation -----------------------------------------------------
dicts = build_params(B, cfg_name="env1")
for d in dicts:
    d["directed_stimulation"] = True
venv = BatchedKuramotoVecEnv(dicts)
rng = np.random.default_rng(0)
out["env1_directed_4096"] = timed(venv, lambda o, r, k: rng.uniform(-1, 1, (B, 1)).astype(np.float32), STEPS)
venv.close()

# ---- configs[3]: env2 with drift, short episodes ------------------------------------------------------
B2 = 2048
dicts = build_params(B2, cfg_name="env2")
for d in dicts:
    d["total_episode_len"] = 90          # 100-step episodes
venv = BatchedKuramotoVecEnv(dicts, compat_env2=True)
out["env2_drift_2048_100step_episodes"] = timed(venv, lambda o, r, k: rng.uniform(-1, 1, (B2, 1)).astype(np.float32), 300)
h0 = venv.core.hosts[0]
out["env2_drift_2048_100step_episodes"].update(resets=h0.reset_count, elec_coords_env0=str(h0.elec_coords),
                                                encapsulation_env0=h0.encapsulation_coeff)
venv.close()

# ---- configs[4]: N = 256 -----------------------------------------------------------------------------
B4 = 8192
np.random.seed(3)
w0, nc, ng, w0t, wl, lm = utils.generate_w0_with_locus(256, [8, 8, 8], 0.1, [2, 4, 4], 0.55, 17, 1, show=False)
base = build_params(1, cfg_name="env0")[0]
dicts = []
for e in range(B4):
    d = copy.copy(base)
    d.update(num_oscillators=256, w0=w0.copy(), w0_without_locus=w0t.copy(), locus_without_w0=wl, locus_mask=lm,
             neur_coords=nc, neur_grid=ng, elec_coords=[[2, 3, 4]], rand_seed=100 + e)
    dicts.append(d)
venv = BatchedKuramotoVecEnv(dicts)
out["sweep_N256_8192"] = timed(venv, lambda o, r, k: rng.uniform(-1, 1, (B4, 1)).astype(np.float32), STEPS)
out["sweep_N256_8192"]["coupling"] = venv.core.engine.coupling
venv.close()
print(json.dumps(out, indent=1))
