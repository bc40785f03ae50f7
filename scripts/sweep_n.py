#!/usr/bin/env python
"""Oscillator-count sweep (BASELINE configs[4]) on the resident-state GRID kernel: 8 x 8 x gz grids
(lines of 8 along y, gz z-planes), N = 64*gz, holding B*N = 2 097 152.  cos coupling, K/N scaling as in
env.py:264.  Engine-level (device-resident) timing.  N <= 4096: one CTA (N = 512: one worker) per environment;
N >= 8192: one thread-block cluster of N / 4096 CTAs per environment (cluster mode of the step kernel).
One GPU, or under torchrun one process per GPU (weak scaling: every rank runs the same B environments per N, the
per-N kernel time is the MAX over ranks, the reported rates are the whole job's):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 scripts/sweep_n.py"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
RANK, WORLD, LOCAL = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(LOCAL)
if WORLD > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", LOCAL))
from dbsgym_b200.engine import KuramotoEngine
from dbsgym_b200.geometry import coupling_table, neuron_grid, ElectrodeModel
from dbsgym_b200.schedule import StepSchedule, transient_grid

out = []
# 8 x 8 x gz grids (lines of 8), then the cubic grids of SURVEY.md 8d config 5 that fit one CTA: the first 4 / 8 / 16
# z-planes of the 16 x 16 x 16 grid (lines of 16, two threads per line)
GRIDS = ([(8, 8, gz) for gz in (4, 8, 16, 32, 64, 128, 256, 512, 1024)] + [(16, 16, gz) for gz in (4, 8, 16)] +
         [(32, 32, gz) for gz in (8, 16, 32, 64)])       # 32^3 (and 32 x 32 x 64 for N = 65536): lines of 32, cluster mode
if os.environ.get("SWEEP_CONFIG5"):                      # the grids SURVEY.md 8d names for BASELINE configs[4]: first-N rows of 8^3, 16^3, 32^3
    GRIDS = [(8, 8, 4), (8, 8, 8)] + [(16, 16, gz) for gz in (4, 8, 16)] + [(32, 32, gz) for gz in (8, 16, 32, 64)]
if os.environ.get("SWEEP_ONLY_CUBIC"):
    GRIDS = [g for g in GRIDS if g[1] != 8]
if os.environ.get("SWEEP_MAX_N"):
    GRIDS = [g for g in GRIDS if g[0] * g[1] * g[2] <= int(os.environ["SWEEP_MAX_N"])]
if os.environ.get("SWEEP_MIN_N"):
    GRIDS = [g for g in GRIDS if g[0] * g[1] * g[2] >= int(os.environ["SWEEP_MIN_N"])]
for gx, gy, gz in GRIDS:
    N = gx * gy * gz
    B = 2097152 // N
    coords, grid = neuron_grid(gx, gy, gz, N, 0.1)
    table = coupling_table(coords, grid, [gx, gy, gz], "cos")
    assert table is not None
    opts = None
    if os.environ.get("SWEEP_CTA_OSC"):                  # (tuning) oscillators per CTA: clusters of N / that many CTAs
        per = int(os.environ["SWEEP_CTA_OSC"])
        if N > per and N // per <= 16:
            opts = {"force_cluster": N // per}
    eng = KuramotoEngine(B, N, [gx, gy, gz], 2340, 0.52, precision="f32", coupling_table=table, device=LOCAL, options=opts)
    lowrank = None
    warp_grid = bool(os.environ.get("SWEEP_SPECTRAL")) and (gx, gy) == (8, 8) and gz in (4, 8)
    if os.environ.get("SWEEP_LOWRANK") and not warp_grid:  # the operator in its truncated eigenbasis (step-kernel variants 11 / 13)
        from dbsgym_b200.geometry import grid_lowrank_factors, grid_sector_factors
        t_f = time.perf_counter()
        tol = float(os.environ.get("SWEEP_LOWRANK_TOL", "1e-9"))
        if os.environ["SWEEP_LOWRANK"] == "sectors":      # sector form: eigenvectors over the fundamental octant only
            f = grid_sector_factors(table, gx, gy, gz, tol=tol)
            if f is None:
                continue
            eng.set_coupling_lowrank_sectors(*f)
            lowrank = {"rank": int(np.count_nonzero(f[2])), "padded_modes": int(f[0][8]), "form": "sectors",
                       "residual_over_lambda_max": float(f[3] / np.abs(f[2]).max()), "factorisation_s": time.perf_counter() - t_f}
        else:
            f = grid_lowrank_factors(table, gx, gy, gz, tol=tol)
            if f is None:
                continue
            eng.set_coupling_lowrank(*f)
            lowrank = {"rank": int(f[0].shape[0]), "form": "plain", "residual_over_lambda_max": float(f[2] / abs(f[1][0])),
                       "factorisation_s": time.perf_counter() - t_f}
    if warp_grid:                                          # the warp kernels (variants 12 / 10), as BatchedKuramoto selects them
        from dbsgym_b200.geometry import spectral_factors
        vecs, vals, ranks, residual = spectral_factors(table, gx, gy, gz, tol=float(os.environ.get("SWEEP_LOWRANK_TOL", "1e-9")))
        eng.set_coupling_spectral(vecs, vals, ranks, residual)
        lowrank = {"rank": int(sum(ranks)), "padded_modes": int(sum(ranks)), "form": "sectors", "kernel": "warp",
                   "residual_over_lambda_max": float(residual / np.abs(vals).max())}
    tt = transient_grid(200.0, 0.05)
    sched = StepSchedule(400, tt[-1], 0.15, 0.75, 0.05)
    eng.set_schedule(sched); eng.set_reward("bbpow_action", 0.05)
    rng = np.random.default_rng(gz + 1000 * RANK)
    # env.py:94's contact-index formula assumes a cubic grid; for the elongated sweep grids the contact is
    # simply the neuron nearest to the grid centre, with the reference's conductance law max(0, 1 - 0.1 d)
    from dbsgym_b200.geometry import distances_from
    centre = int(np.argmin(np.abs(grid - np.array([gx // 2, gy // 2 - 1, gz // 2])).sum(axis=1)))
    stim = np.tile(np.maximum(0.0, 1.0 - distances_from(grid * 0.1, [centre])[0]), (B, 1))
    w0 = np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02
    y0 = rng.normal(np.pi, 0.6, (B, N))
    eng.set_env_params(None, w0=w0, stim=stim, rec=stim, y0=y0)
    eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
    t0 = time.perf_counter(); eng.transient(tt); torch.cuda.synchronize(); t_tr = time.perf_counter() - t0
    dev = torch.device("cuda", LOCAL)
    act = torch.from_numpy(rng.uniform(-1, 1, (40, B)).astype(np.float32)).to(dev)
    obs = torch.empty((B, 2340), dtype=torch.float32, device=dev)
    rew = torch.empty(B, dtype=torch.float32, device=dev); done = torch.empty(B, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    eng.set_timing(True); eng.counters(reset=True)
    for i in range(5 if N <= 8192 else 2):     # (warm-up)
        eng.step_device(act[i].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), st)
    torch.cuda.synchronize(); eng.counters(reset=True)
    ms = []
    n_timed = 30 if N <= 8192 else 6
    for i in range(5, 5 + n_timed):
        eng.step_device(act[i].data_ptr(), obs.data_ptr(), rew.data_ptr(), done.data_ptr(), st)
        torch.cuda.synchronize(); ms.append(eng.last_step_ms())
    c = eng.counters()
    k_step, k_obs = np.mean([m[0] for m in ms]), np.mean([m[1] for m in ms])
    if WORLD > 1:                                    # slowest rank
        tmax = torch.tensor([k_step, k_obs], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        k_step, k_obs = (float(v) for v in tmax.cpu())
    rhs = (c["rhs_evals"] - eng.rhs_reused()) / (n_timed * B)       # executed evaluations (the reference's count is 32)
    ypar = bool(eng.lib.dbsgym_build_flags() & 1)          # own op count of the sector contraction, see bench.py
    variant = eng.step_variant(B)
    SYM_FLOP, SYM_LIN = ((148 if variant == 4 else 196 if ypar else 304) / 256.0), ((160 if ypar else 128) / 8.0)
    if gy != 8:               # lines of 16 / 32: per (zj,xj) block and thread 24 / 48 FFMA2 (sector row) + 2 / 4 x 64 FFMA2,
        SYM_FLOP, SYM_LIN = (4 * 152) / 512.0, 128 / 8.0          # N^2/512 resp. N^2/1024 of those: 1.1875 N^2 flop per RHS
    out.append({"N": N, "grid": [gx, gy, gz], "n_gpus": WORLD, "envs_per_gpu": B, "envs": B * WORLD, "step_kernel_ms": float(k_step), "obs_kernel_ms": float(k_obs),
                "env_steps_per_s": WORLD * B / ((k_step + k_obs) * 1e-3), "oscillator_updates_per_s": WORLD * B * N * (c["accepted"] + c["rejected"]) / (n_timed * B) / ((k_step + k_obs) * 1e-3),
                "rhs_per_env_step": rhs, "executed_tflops": WORLD * rhs * (SYM_FLOP * N * N + SYM_LIN * N) * B / (k_step * 1e-3) / 1e12,
                "dense_equivalent_tflops": WORLD * rhs * 4 * N * N * B / (k_step * 1e-3) / 1e12,
                "transient_s": t_tr, "status": c["status"], "ctas_per_env": max(1, N // 4096), "variant": variant})
    if lowrank is not None:                               # own op count of the low-rank contraction: 2 x 4 flop per mode and oscillator
        out[-1].update(lowrank=lowrank, executed_tflops=WORLD * rhs * (16.0 * lowrank.get("padded_modes", lowrank["rank"]) * N / (8 if lowrank["form"] == "sectors" else 1) + 150.0 * N) * B / (k_step * 1e-3) / 1e12)
    if RANK == 0:
        print(json.dumps(out[-1]), flush=True)
    eng.close()
if WORLD > 1:
    dist.destroy_process_group()
