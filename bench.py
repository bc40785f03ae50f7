#!/usr/bin/env python
"""Throughput benchmark of the DBS-Gym environment step (BASELINE.json metric: env-steps/s and
oscillator-updates/s at 1/2/4/8 B200 beside the reference's CPU step()).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU step() (oracle port)

Workload (``config.workload``): BASELINE.json configs[2] -- env1 (distance-weighted recording,
spatial electrode variation), N = 512 oscillators, 4096 environments PER GPU (weak scaling), one
uniform(-1,1) action per environment per step, synthetic inputs (w0 drawn by the reference's own
sampler, N(pi,0.6) initial phases, 200-unit transient).  One "step" = one batched VecEnv step =
4096 env-steps per GPU.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_OSC = 512
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


# ------------------------------------------------------------------------------------------
def build_params(n_envs, seed0=10, distinct_w0=32, cfg_name="env1", reward="bbpow_action"):
    """params_dict list the way aDBS_RL/train_aDBS_RL.py:95-112 builds one, with per-env seeds.
    The w0 sampler (polyfit + quad) costs ~5 ms, so `distinct_w0` frequency sets are cycled."""
    import importlib
    from dbsgym_b200 import utils
    cfg = importlib.import_module(f"dbsgym_b200.configs.{cfg_name}")
    base = cfg.params_dict_train
    sets = []
    for s in range(distinct_w0):
        np.random.seed(seed0 + s)
        sets.append(utils.generate_w0_with_locus(
            cfg.n_neurons, cfg.grid_size, cfg.coord_modif, locus_center=base["locus_center"],
            locus_size=base["locus_size"], wmuL=base["wmuL"], wsdL=base["wsdL"], show=False))
    dicts = []
    for e in range(n_envs):
        w0, nc, ng, w0t, wl, lm = sets[e % distinct_w0]
        d = copy.deepcopy(base)
        d.update(w0=w0.copy(), w0_without_locus=w0t.copy(), locus_without_w0=wl, locus_mask=lm,
                 neur_coords=sets[0][1], neur_grid=sets[0][2], reward_func=reward, verbose=0,
                 rand_seed=seed0 + e)
        dicts.append(d)
    return dicts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------------------------------
def cpu_port_steps_per_sec(n_steps, rhs_form, dtype, seed=10):
    """The oracle port of the reference step(), one process, one env (env1)."""
    from oracle import kuramoto_oracle as ko
    d = build_params(1, seed0=seed, distinct_w0=1)[0]
    env = ko.OracleEnv(d, rhs_form="matvec", dtype=dtype)         # transient with the fast RHS (not timed)
    env.kuramoto.rhs_form = rhs_form
    acts = np.random.default_rng(0).uniform(-1, 1, n_steps).astype(np.float32)
    env.step(np.array([acts[0]], dtype=np.float32))
    t0 = time.perf_counter()
    for a in acts:
        env.step(np.array([a], dtype=np.float32))
    return n_steps / (time.perf_counter() - t0)


def _ref_worker(args):
    seed, n_steps, barrier_t = args
    from oracle import kuramoto_oracle as ko
    d = build_params(1, seed0=seed, distinct_w0=1)[0]
    env = ko.OracleEnv(d, rhs_form="matvec", dtype=np.float32)
    env.kuramoto.rhs_form = "as_written"
    rng = np.random.default_rng(seed)
    out = []
    for k in range(len(n_steps)):
        t0 = time.perf_counter()
        for _ in range(n_steps[k]):
            env.step(rng.uniform(-1, 1, 1).astype(np.float32))
        out.append(time.perf_counter() - t0)
    return out


def run_reference_arm(args):
    """--impl reference: the reference's CPU formulation of step() (N x N sine RHS as written at
    environment/env.py:252-256, float32 like the JAX default) through the oracle port, one
    environment per process on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    # env-steps per process per bench step, sized so that the whole run stays within a few minutes
    per_step = 2 if args.steps <= 300 else 1
    plan = [per_step] * (args.warmup + args.steps)
    with mp.get_context("spawn").Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(100 + i, plan, 0.0) for i in range(cores)])
        wall = time.perf_counter() - t0
    per_proc = np.array(res)                                  # [cores, warmup+steps] seconds
    timed = per_proc[:, args.warmup:].sum(axis=1).max()       # slowest process over the timed steps
    total_env_steps = cores * per_step * args.steps
    value = total_env_steps / timed
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * timed / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "env1 N=512, one env per host process, reference RHS as written (N x N sin), float32",
                   "env_steps_per_bench_step": cores * per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x {per_step} env-steps x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "oscillator_updates_per_sec": value * N_OSC * 5,
        "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def verbatim_reference_steps_per_sec(n_steps=16):
    """the reference's OWN environment/env.py imported verbatim (numpy for jax.numpy, the
    restated diffrax) -- env0 params_dict_train, np.random.seed(10), actions default_rng(0).uniform(-1, 1); a bounded
    sample of the 2048-step rollout.  Only where /root/reference exists (not on the GPU box)."""
    from oracle import run_reference
    if not run_reference.reference_available():
        return None
    env_mod, utils_mod, cfgs = run_reference.load_reference()
    cfg = cfgs["env0"]
    np.random.seed(10)
    base = copy.deepcopy(cfg.params_dict_train)
    w0, nc, ng, w0t, wl, lm = utils_mod.generate_w0_with_locus(
        cfg.n_neurons, cfg.grid_size, cfg.coord_modif, locus_center=base["locus_center"], locus_size=base["locus_size"],
        wmuL=base["wmuL"], wsdL=base["wsdL"], show=False)
    base.update(w0=w0, w0_without_locus=w0t, locus_without_w0=wl, locus_mask=lm, neur_coords=nc, neur_grid=ng,
                reward_func="bbpow_action", verbose=0)
    env = env_mod.SpatialKuramoto(base)
    acts = np.random.default_rng(0).uniform(-1, 1, 2048).astype(np.float32)
    env.step(acts[:1])
    t0 = time.perf_counter()
    for a in acts[1:1 + n_steps]:
        env.step(np.array([a], dtype=np.float32))
    return n_steps / (time.perf_counter() - t0)


def step_kernel_model(variant, rhs_exec, ypar, n_osc=N_OSC):
    """Executed FP32 flop of one env-step of the step kernel (FFMA = 2, FFMA2 = 4, FADD / FMUL = 1, FADD2 / FMUL2 = 2 per
    lane), from the instruction counts of the contraction per thread; DESIGN.md section 3 derives them and
    profiles/*_flop_count.json holds the ncu-measured count of the same kernel for comparison."""
    N = n_osc
    other = 700 * N                                      # tableau combinations, error norm, dense output, LFP, reward
    if variant == 10:
        # spectral kernel, one warp per environment, 32 modes; per LANE (16 oscillators) and RHS evaluation: stage argument
        # (20 tableau rows per 6 evaluations x 16 FFMA + the y1 add) 109, range reduction of the sincos 16 x 6 = 96, the two
        # sector transforms 2 x (24 FADD2 + 24 FFMA2) = 288, projection 32 x (FMUL2 + FFMA2) = 192, mode sums
        # 2 x 17 packed adds / multiplies = 68, expansion 16 FMUL2 + 48 FFMA2 = 224, k = c0 + c A s - s A c 64;
        # per lane and RK sub-step: error estimate 288, y1 - y0 160, dense-output coefficients 384, 3.6 LFP samples x 16 x 16,
        # accept 64 (profiles/r02_step_kernel_variant10_counts.json: the SASS execution counts give 1.37 Mflop per env-step)
        per_lane_rhs = 109 + 96 + 288 + 192 + 68 + 224 + 64
        per_lane_substep = 288 + 160 + 384 + 920 + 64
        return (rhs_exec * per_lane_rhs + 5 * per_lane_substep) * 32, "spectral, one warp per environment (32 modes, 7 + 4 x 6 + 1 over the parity sectors)"
    if variant == 9:
        # spectral kernel, compiled ranks (9 even, 4 odd), per thread (8 oscillators) and RHS evaluation:
        # projection 13 x (FMUL2 + 3 FFMA2) = 182, row sums 52 x (15 FADD2 + FMUL2) / 64 = 26, expansion 4 x (FMUL2 + 8 FFMA2)
        # + 4 x (FMUL2 + 3 FFMA2) = 192, two sector butterflies 2 x 16 FFMA2 = 128, y folds 32, k = c0 + c A s - s A c 32,
        # range reduction of the sincos 40, stage combination (mean 27 FFMA + 16) 70
        per_thread = 182 + 26 + 192 + 128 + 32 + 32 + 40 + 70
        return rhs_exec * per_thread * (N // 8) + other, "spectral (34 modes in 9 + 4 padded slots per sector pair)"
    blk = 148 if variant == 4 else (196 if (ypar and variant in (3, 6)) else 304)
    per_rhs = (blk / 256.0) * N * N + ((160 if ypar else 128) / 8.0) * N
    return rhs_exec * per_rhs + other, "exact parity-sector blocks"


def measured_kernel_counts(variant):
    """ncu-measured per-launch numbers of the step kernel committed with the round (scripts/ncu_counts.py):
    dram bytes and executed FP32 flop at 4096 environments; None when no capture of this variant is committed."""
    path = os.path.join(ROOT, "profiles", f"r02_step_kernel_variant{variant}_counts.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:  # noqa: BLE001
        return None


def run_gpu_arm(args):
    # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from dbsgym_b200 import _capi
    from dbsgym_b200.engine import measure_fp32_peak, measure_mufu_peak
    from dbsgym_b200.sharding import gather_episode_stats, shard_bounds
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv

    if args.strong:                                      # fixed total work: args.envs environments over all ranks
        lo, hi = shard_bounds(args.envs, rank, world)
        B, seed_off = hi - lo, lo
        total_envs = args.envs
    else:
        B, seed_off = args.envs, rank * args.envs
        total_envs = B * world
    t_setup = time.perf_counter()
    dicts = build_params(B, seed0=10 + seed_off)
    for d in dicts:
        d["precision"] = args.precision
    venv = BatchedKuramotoVecEnv(dicts, device=local_rank, coupling_eval=args.coupling_eval)
    venv.reset()
    eng = venv.core.engine
    setup_s = time.perf_counter() - t_setup
    K, Wm = args.steps, args.warmup
    rng = np.random.default_rng(1234 + rank)
    actions = rng.uniform(-1, 1, (Wm + K, B)).astype(np.float32)

    # ---------- (1) device-resident timing: inputs already in HBM, CUDA events per step ----------
    act_dev = torch.from_numpy(actions).to(dev)
    obs_dev = torch.empty((B, venv.core.window), dtype=torch.float32, device=dev)
    rew_dev = torch.empty(B, dtype=torch.float32, device=dev)
    done_dev = torch.empty(B, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2
    flush_rd = torch.zeros(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def flush_l2():
        # write a buffer larger than L2 (evicts everything the previous step left), then read another one: the timed
        # kernels start on a cold L2 that holds CLEAN lines -- after the write alone, their first loads would also pay
        # for writing the flush buffer's dirty lines back to HBM (traffic that belongs to the flush, not to the step)
        flush.zero_()
        flush_rd.max()
    stream = torch.cuda.current_stream()
    eng.set_timing(True)
    eng.set_episode(None, step_idx=0, episode_len=2 ** 30)          # no resets inside the timed region
    eng.counters(reset=True)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.start()                                                  # nvidia-smi needs ~1 s to start sampling
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kern_ms = []
    for i in range(Wm):
        flush_l2()
        eng.step_device(act_dev[i].data_ptr(), obs_dev.data_ptr(), rew_dev.data_ptr(), done_dev.data_ptr(),
                        stream.cuda_stream)
    sync_all()
    eng.counters(reset=True)
    eng.launch_count(reset=True)
    n_clk0 = len(clocks.rows)
    t_wall = time.perf_counter()
    for i in range(K):
        flush_l2()                                                  # L2 flush between timed iterations
        ev[i][0].record(stream)
        eng.step_device(act_dev[Wm + i].data_ptr(), obs_dev.data_ptr(), rew_dev.data_ptr(),
                        done_dev.data_ptr(), stream.cuda_stream)
        ev[i][1].record(stream)
        if i % 8 == 7 or i == K - 1:
            torch.cuda.synchronize()
            kern_ms.append(eng.last_step_ms())
    sync_all()
    wall_dev = time.perf_counter() - t_wall
    n_clk1 = len(clocks.rows)
    launches_dev = eng.launch_count(reset=True)                     # kernels launched by the library in the timed region
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    dev_time_s = float(step_ms.sum()) * 1e-3
    counters = eng.counters()
    rhs_reused = eng.rhs_reused()

    # ---------- (1b) sustained: back-to-back steps for >= args.sustain seconds (no flush, two events), clocks sampled ----------
    sustained = None
    if args.sustain > 0:
        n_sus = max(K, int(args.sustain / max(dev_time_s / K, 1e-6) * 1.15))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        c0 = len(clocks.rows)
        e0.record(stream)
        for i in range(n_sus):
            eng.step_device(act_dev[Wm + i % K].data_ptr(), obs_dev.data_ptr(), rew_dev.data_ptr(), done_dev.data_ptr(),
                            stream.cuda_stream)
        e1.record(stream)
        sync_all()
        sus_s = e0.elapsed_time(e1) * 1e-3
        sm = [float(r[1]) for r in clocks.rows[c0:] if len(r) > 1 and r[1].replace(".", "").isdigit()]
        sustained = {"steps": n_sus, "seconds": sus_s, "value_this_rank": B * n_sus / sus_s,
                     "sm_mhz_median": float(np.median(sm)) if sm else None, "clock_samples": len(sm),
                     "note": "back-to-back launches, no L2 flush, one CUDA-event pair around the whole run"}

    # ---------- (2) end to end through the public VecEnv API with host buffers ----------
    eng.set_timing(False)
    for i in range(Wm):
        venv.step_async(actions[i].reshape(B, 1)); venv.step_wait()
    sync_all()
    eng.launch_count(reset=True)
    n_new_total = 0
    t0 = time.perf_counter()
    for i in range(K):
        venv.step_async(actions[Wm + i].reshape(B, 1))
        obs, rew, done, infos = venv.step_wait()
    sync_all()
    e2e_s = time.perf_counter() - t0
    launches_e2e = eng.launch_count(reset=True)
    # bytes that crossed PCIe per step: every new window sample is stored twice in the pinned host log (zero-copy stores),
    # plus the control block (reward f32, sample count i32, log position i32, done u8); counted from the step counters
    n_new_total = int(eng.lfp()[2].astype(np.int64).sum())          # samples of the LAST step, every environment
    d2h_e2e = n_new_total * 2 * 4 + B * (4 + 4 + 4 + 1)
    assert isinstance(obs, np.ndarray) and obs.shape == (B, 1, venv.core.window)
    # the C-ABI host call that returns the FULL [B, W] observation every step (no host-side window mirror)
    core = venv.core
    for i in range(Wm):
        eng.step_host(actions[i], core._obs_bufs[0], core.rew_buf, core.done_buf)
    sync_all()
    t0 = time.perf_counter()
    for i in range(K):
        eng.step_host(actions[Wm + i], core._obs_bufs[0], core.rew_buf, core.done_buf)
    sync_all()
    capi_s = time.perf_counter() - t0
    clk = clocks.stop()
    clk["samples_in_device_loop"] = n_clk1 - n_clk0

    # ---------- max over ranks ----------
    sus_s = sustained["seconds"] if sustained else 0.0
    times = torch.tensor([dev_time_s, e2e_s, capi_s, sus_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        # NCCL is used only off the step path: gather per-env episode statistics
        if not args.strong:
            stats = gather_episode_stats(np.stack([rew.astype(np.float64), done.astype(np.float64)], axis=1),
                                         B * world, rank, world, device=dev)
            assert stats.shape == (B * world, 2)
    dev_time_s, e2e_s, capi_s, sus_s = (float(v) for v in times.cpu())

    if rank == 0:
        total_env_steps = total_envs * K
        value = total_env_steps / dev_time_s
        rhs_per_env_step = counters["rhs_evals"] / (B * K)               # the reference's count (32)
        rhs_exec = (counters["rhs_evals"] - rhs_reused) / (B * K)        # executed: the first stage of a segment is carried over
        substeps_per_env_step = (counters["accepted"] + counters["rejected"]) / (B * K)
        ypar = bool(_capi.load().dbsgym_build_flags() & 1)
        variant = eng.step_variant(B)
        dense_flop_per_env_step = rhs_per_env_step * 4 * N_OSC * N_OSC + 700 * N_OSC
        flop_per_env_step, flop_model = step_kernel_model(variant, rhs_exec, ypar)
        kernel_name = {0: "step_kernel<float,GRID>", 1: "step_kernel<DENSE>", 2: "step_kernel<GRID_SYM>",
                       3: "step_kernel<float,GRID_SYM,8x8x8" + (",y-parity>" if ypar else ">"),
                       4: "step_kernel<float,GRID_SYM,8x8x8,y-parity,multi-worker (8 envs per CTA, precomputed sector coefficients)>",
                       5: "step_kernel<cluster>", 6: "step_kernel<float,GRID_SYM,gx=8>",
                       9: "step_kernel<float,SPECTRAL,8x8x8 (generalised mean-field identity, 8 envs per CTA)>",
                       10: "warp_step_kernel<RankSet<7,4,4,4,4,4,4,1>> (generalised mean-field identity, one warp per environment, "
                           "octant ownership, 8 environments per SM)",
                       11: "step_kernel<float,LOWRANK> (any operator in its truncated eigenbasis)"}.get(variant, "step_kernel")
        k_step = float(np.mean([m[0] for m in kern_ms])) * 1e-3
        k_obs = float(np.mean([m[1] for m in kern_ms])) * 1e-3
        peaks, peak_src = measured_peaks()
        peak_ffma = measure_fp32_peak(local_rank, packed=False)
        peak_ffma2 = measure_fp32_peak(local_rank, packed=True)
        peak_mufu = measure_mufu_peak(local_rank)
        fp32_peak = max(peak_ffma, peak_ffma2)
        achieved_tf = flop_per_env_step * B / k_step / 1e12
        counts = measured_kernel_counts(variant)
        traffic = None
        if counts and counts.get("n_envs"):
            traffic = (counts["dram_bytes_read"] + counts["dram_bytes_write"]) * B / counts["n_envs"]
        fused_obs = not venv.core.engine.options.get("no_fused_obs", False)
        obs_bytes = B * (2 * 2340 * 4)                              # ring read + chronological observation write
        cpu_n = 12
        cpu_as_written = cpu_port_steps_per_sec(cpu_n, "as_written", np.float32)
        cpu_matvec = cpu_port_steps_per_sec(200, "matvec", np.float64)
        verbatim = verbatim_reference_steps_per_sec(16) if args.cpu_reference else None
        mufu_per_env_step = (rhs_exec * 2 * N_OSC + 19 * N_OSC)       # sin + cos per oscillator and RHS, cos per LFP sample
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 * dev_time_s / K, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f32" if args.precision == "f32" else "f64", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: env1, N=512 oscillators, 4096 envs per GPU, uniform(-1,1) actions"
                                   if not args.strong else
                                   f"BASELINE configs[2] (env1, N=512), STRONG scaling: {total_envs} environments in total over all GPUs",
                       "envs_per_gpu": B, "global_envs": total_envs, "precision": args.precision,
                       "coupling": eng.coupling, "coupling_eval": venv.core.coupling_eval,
                       "spectral": eng.spectral,
                       "l2": "flush between timed iterations: a 256 MB buffer written, then another 256 MB buffer read (cold, clean L2)",
                       "timing": "sum of per-step CUDA-event intervals on the launch stream, max over ranks"},
            "oscillator_updates_per_sec": value * N_OSC * substeps_per_env_step,
            "oscillator_rhs_evals_per_sec": value * N_OSC * rhs_exec,          # executed evaluations
            "rk_substeps_per_env_step": substeps_per_env_step, "rhs_evals_per_env_step": rhs_per_env_step,
            "rhs_evals_executed_per_env_step": rhs_exec,
            "solver_status": counters["status"],
            "e2e": {"value": total_envs * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 4,
                    "d2h_bytes_per_step": d2h_e2e, "gpu_launches": launches_e2e,
                    "api": "BatchedKuramotoVecEnv.step_async/step_wait (copy_obs=False), numpy actions in, numpy obs/reward/done out. "
                           "Observations are views of a pinned host sample log the step kernel appends to through mapped memory "
                           "(each new sample stored twice, ~18 samples/step); later steps never overwrite a window already handed out "
                           "(>= 13 steps, resets switch buffers), which is what SB3's rollout buffers need (dbsgym_step_host_mirror)",
                    "full_obs_d2h": {"value": total_envs * K / capi_s, "d2h_bytes_per_step": B * (2340 * 4 + 4 + 1),
                                     "api": "dbsgym_step_host (whole [B,2340] f32 observation copied to pinned host memory every step)"}},
            "gpu_launches": launches_dev,      # counted by the library: step kernel + observation copy kernel per device-timed step
            "roofline": {"bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak if fp32_peak else None,
                         "traffic": traffic,
                         "traffic_source": (counts or {}).get("source"),
                         "kernel": kernel_name,
                         "kernel_ms": k_step * 1e3, "flop_per_env_step": flop_per_env_step, "flop_model": flop_model,
                         "flop_per_env_step_ncu": (counts or {}).get("fp32_flop_per_env_step"),
                         "dense_formulation_flop_per_env_step": dense_flop_per_env_step,
                         "dense_equivalent_tflops": dense_flop_per_env_step * B / k_step / 1e12,
                         "peak_source": "best of the FFMA and FFMA2 micro-benchmarks run in this process (dbsgym_measure_fp32_peak_mode); nominal 74.4",
                         "peak_ffma": peak_ffma, "peak_ffma2": peak_ffma2,
                         "ncu_pipes": {k: (counts or {}).get(k) for k in ("issue_slots_busy_pct", "fma_pipe_cycles_active_pct", "xu_pipe_inst_pct",
                                                                          "lsu_pipe_inst_pct", "warps_active_pct", "registers_per_thread")},
                         "issue": {"warp_instructions_per_env_step_ncu": (counts or {}).get("warp_inst_per_env_step"),
                                   "frac_of_issue_slots": ((counts or {}).get("warp_inst_per_env_step") or 0) * B / k_step /
                                                          (4 * 148 * (clk.get("sm_mhz") or 1965.0) * 1e6) or None,
                                   "note": "share of the 4 x 148 warp-issue slots per cycle the kernel fills: what actually bounds the spectral kernel"},
                         "mufu": {"achieved_tops": mufu_per_env_step * B / k_step / 1e12, "peak_tops": peak_mufu,
                                  "frac": mufu_per_env_step * B / k_step / 1e12 / peak_mufu if peak_mufu else None,
                                  "peak_source": "dbsgym_measure_mufu_peak in this process (nominal 148 x 16 x 1.965 GHz = 4.65)"}},
            "roofline_obs": {"bound": "hbm", "achieved": obs_bytes / k_obs / 1e9, "peak": peaks.get("hbm_gbs"),
                             "unit": "GB/s", "frac": obs_bytes / k_obs / 1e9 / peaks.get("hbm_gbs", 1.0),
                             "kernel": "obs_copy_kernel (ring -> chronological [B,W] f32; append/reward fused into the step kernel)" if fused_obs else "obs_kernel",
                             "kernel_ms": k_obs * 1e3, "peak_source": peak_src},
            "cpu_baseline": {"value": cpu_as_written, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{cpu_n} env1 steps, one env, oracle port with the reference's N x N sine RHS in float32",
                             "matvec_f64_value": cpu_matvec, "published_reference": "17-20 it/s (notebook tqdm, unknown CPU)",
                             "reference_verbatim": None if verbatim is None else {
                                 "value": verbatim, "kind": "reference", "cores": 1,
                                 "sample": "16 of the 2048 steps of BASELINE configs[0] (env0, seed 10): the reference's own "
                                           "environment/env.py imported verbatim over numpy + the restated diffrax, float64"}},
            "sustained": None if sustained is None else dict(sustained, value=total_envs * sustained["steps"] / sus_s),
            "clocks": clk, "setup_s": setup_s, "wall_s_device_loop": wall_dev,
        }
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    venv.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--precision", default="f32", choices=["f32", "f64"])
    ap.add_argument("--coupling-eval", default="auto", choices=["auto", "exact", "spectral"],
                    help="float32 coupling evaluation: spectral (default where it applies) or the exact sector blocks")
    ap.add_argument("--strong", action="store_true", help="strong scaling: --envs is the TOTAL over all GPUs")
    ap.add_argument("--sustain", type=float, default=5.0, help="seconds of back-to-back steps for the sustained figure (0 = skip)")
    ap.add_argument("--no-cpu-reference", dest="cpu_reference", action="store_false",
                    help="skip timing the verbatim reference env.py (only possible where /root/reference exists)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
