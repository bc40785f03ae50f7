"""Generate the golden fixtures in this directory from the REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    PYTHONPATH=/root/repo python tests/golden/make_golden.py

The reference's ``environment/env.py`` / ``utils.py`` / ``env_configs`` are imported
verbatim on top of ``oracle/shims`` (numpy float64 instead of JAX float32, the
diffrax solver restated in ``oracle/diffrax_restated.py`` -- see that file's
"parity unpinned" note).  Everything else in the fixtures -- geometry, electrode
indices and conductances, the np.arange time grids, RNG draw order, LFP, window,
rewards -- is produced by reference code, not by a restatement.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import diffrax_restated as dr            # noqa: E402
from oracle.run_reference import load_reference      # noqa: E402

ref_env, ref_utils, ref_cfgs = load_reference()


def build_params(cfg, base, seed, reward, **over):
    np.random.seed(seed)
    w0, nc, ng, w0t, wl, lm = ref_utils.generate_w0_with_locus(
        cfg.n_neurons, cfg.grid_size, cfg.coord_modif, locus_center=base["locus_center"],
        locus_size=base["locus_size"], wmuL=base["wmuL"], wsdL=base["wsdL"], show=False)
    d = copy.deepcopy(base)
    d.update(w0=w0, w0_without_locus=w0t, locus_without_w0=wl, locus_mask=lm,
             neur_coords=nc, neur_grid=ng, reward_func=reward, verbose=0)
    d.update(over)
    return d


def record_steps(name, cfg_name, seed, n_steps, reward="bbpow_action", action_seed=0, **over):
    cfg = ref_cfgs[cfg_name]
    d = build_params(cfg, cfg.params_dict_train, seed, reward, **over)
    inputs = {k: np.array(d[k]) for k in ("w0", "w0_without_locus", "locus_without_w0", "locus_mask")}
    for k in dr.GLOBAL_STATS:
        dr.GLOBAL_STATS[k] = 0
    env = ref_env.SpatialKuramoto(d)
    stats_reset = dict(dr.GLOBAL_STATS)
    out = dict(inputs)
    out["init_state"] = np.array(env.init_state)
    out["w0_model"] = np.array(env.kuramoto.w0)
    out["stim_cond"] = np.array(env.kuramoto.dbs.conductances)
    out["rec_cond"] = np.array(env.kuramoto.dbs.rec_conductances)
    out["elec_idxs"] = np.array(env.kuramoto.dbs.elec_idxs)
    out["rec_idxs"] = np.array(env.kuramoto.dbs.rec_idxs)
    out["y_after_transient"] = np.array(env.sol_state[-1])
    out["window0"] = np.array(env.theta_state[0])
    out["t_after_transient"] = np.float64(env.current_time)
    out["reset_stats"] = np.array([stats_reset[k] for k in sorted(stats_reset)])
    acts = np.random.default_rng(action_seed).uniform(-1, 1, n_steps).astype(np.float32)
    y_end = np.empty((n_steps, cfg.n_neurons))
    y_seam = np.empty((n_steps, cfg.n_neurons))
    lfp_true = np.full((n_steps, 20), np.nan)
    lfp_rec = np.full((n_steps, 20), np.nan)
    nI = np.empty(n_steps, np.int32)
    nII = np.empty(n_steps, np.int32)
    offs_I = np.full((n_steps, 4), np.nan)
    offs_II = np.full((n_steps, 16), np.nan)
    rew = np.empty(n_steps)
    u = np.empty(n_steps)
    tcur = np.empty(n_steps)
    for k in range(n_steps):
        obs, r, done, trunc, info = env.step(np.array([acts[k]], dtype=np.float32))
        nI[k], nII[k] = len(env.t_eval_step_I), len(env.t_eval_step_II)
        offs_I[k, :nI[k]] = env.t_eval_step_I - env.t_eval_step_I[0]
        offs_II[k, :nII[k]] = env.t_eval_step_II - env.t_eval_step_II[0]
        y_end[k] = env.sol_state[-1]
        y_seam[k] = env.sol_state_[nI[k] - 1]
        s = nI[k] + nII[k] - 1
        lfp_true[k, :s] = env.theta_mean
        lfp_rec[k, :s] = env.theta_records
        rew[k], u[k], tcur[k] = r, env.u[0], env.current_time
    out.update(actions=acts, y_end=y_end, y_seam=y_seam, lfp_true=lfp_true, lfp_rec=lfp_rec,
               nI=nI, nII=nII, offs_I=offs_I, offs_II=offs_II, reward=rew, u=u, t_cur=tcur,
               window_last=np.array(env.theta_state[0]), obs_last=obs[0].copy())
    total = dict(dr.GLOBAL_STATS)
    out["step_stats"] = np.array([total[k] - stats_reset[k] for k in sorted(total)])
    out["stats_keys"] = np.array(sorted(total))
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, "done; per-step stats", out["step_stats"] / n_steps, "reset", out["reset_stats"])
    return env


def record_schedule():
    """env.py:426-441 replayed with the reference's own expressions (no integration)."""
    out = {}
    for tag, n in (("train", 5555), ("eval", 1111)):
        ct = np.arange(0., 200., 0.05)[-1]
        nI, nII, tc = [], [], []
        for _ in range(n):
            a = np.arange(ct, ct + 0.15, 0.05)
            ct = a[-1]
            b = np.arange(ct, ct + 0.75, 0.05)
            ct = b[-1]
            nI.append(len(a)), nII.append(len(b)), tc.append(ct)
        out[f"nI_{tag}"] = np.array(nI, np.int8)
        out[f"nII_{tag}"] = np.array(nII, np.int8)
        out[f"t_final_{tag}"] = np.float64(ct)
        out[f"t_cur_every100_{tag}"] = np.array(tc[::100])
    np.savez_compressed(os.path.join(HERE, "schedule.npz"), **out)
    print("schedule done", {k: int(v.sum()) for k, v in out.items() if k.startswith("nI")})


def record_rewards(env):
    """All three reward functions of one reference env on a few windows."""
    rng = np.random.default_rng(5)
    t = np.arange(2340) * 0.0005
    windows = [np.array(env.theta_state[0]),
               0.3 * np.sin(2 * np.pi * 17.0 * t) + 0.05 * rng.standard_normal(2340),
               0.4 * np.cos(2 * np.pi * 14.1 * t + 0.3) + 0.2 * np.cos(2 * np.pi * 40 * t),
               rng.uniform(-1, 1, 2340)]
    us = [0.0, -3.5, 5.0, 1.25]
    r = np.array([[env.reward_bbpow_action(w, [u]),
                   env.reward_temp_const_lfp_betafilt_action(w, [u]),
                   env.reward_bbpow_threth_action(w, [u])] for w, u in zip(windows, us)])
    bb = np.array([ref_utils.calc_beta_band_power(w, 0.0005, 12.5, 21) for w in windows])
    np.savez_compressed(os.path.join(HERE, "rewards.npz"), windows=np.array(windows), u=np.array(us),
                        rewards=r, bbpow=bb)
    print("rewards", r)


def record_geometry():
    out = {}
    nc, ng = ref_utils.generate_neuron_grid_3D(8, 8, 8, 512, coord_modif=0.1)
    out["neur_grid"] = ng.astype(np.int16)
    out["alpha_row0"] = np.cos(ref_utils.create_distance_matrix(nc))[0]
    out["alpha_row284"] = np.cos(ref_utils.create_distance_matrix(nc))[284]
    out["locus_mask_444_055"] = ref_utils.create_oscillation_locus(ng, [8, 8, 8], [4, 4, 4], 0.55)
    cases = [([[4, 3, 4]], [[1, 1, 1]], 0.1, False), ([[4, 3, 4]], [[1, 1, 1]], 0.15, False),
             ([[5, 2, 3]], [[3, 5, 1]], 0.1, True), ([[1, 6, 2]], [[6, 5, 1]], 2.1, True),
             ([[6, 6, 4]], [[4, 4, 3]], 0.12, True)]
    for i, (ec, rc, cm, directed) in enumerate(cases):
        dbs = ref_env.SimpleDBS([8, 8, 8], ref_utils.create_distance_matrix(ng * cm), ec, rc, ng,
                                amplitudes=[0.], directed_stimulation=directed, prc_type="dummy")
        out[f"case{i}_elec"], out[f"case{i}_rec"] = np.array(ec), np.array(rc)
        out[f"case{i}_cm"], out[f"case{i}_directed"] = np.float64(cm), np.bool_(directed)
        out[f"case{i}_elec_idx"], out[f"case{i}_rec_idx"] = np.array(dbs.elec_idxs), np.array(dbs.rec_idxs)
        out[f"case{i}_cond"] = np.array(dbs.conductances[0])
        out[f"case{i}_rec_cond"] = np.array(dbs.rec_conductances[0])
        if directed:
            out[f"case{i}_masks"] = np.array(dbs.directional_masks_list[0])
    # smaller / non-cubic neuron counts (first-n rows of the grid), utils.py:483-494
    for n, g in ((256, (8, 8, 8)), (100, (5, 5, 5))):
        _, gg = ref_utils.generate_neuron_grid_3D(*g, n, coord_modif=0.1)
        out[f"grid_{n}_{g[0]}"] = gg.astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **out)
    print("geometry done")


def record_env2_events(n_resets=14):
    """Host reset logic under temporal drift (env.py:483-557).  env2 as shipped cannot
    run: the plasticity event is pushed out of range so the reference's own
    code path executes; electrode drift, encapsulation, plasticity-walk regeneration and the
    spatial re-draw all fire.  Transient shortened so this stays a fixture generator."""
    cfg = ref_cfgs["env2"]
    d = build_params(cfg, cfg.params_dict_train, 21, "bbpow_action",
                     plasticity_drift_freq=10 ** 6, transient_state_len=117.5,
                     total_episode_len=9., spatial_var_freq=4)
    env = ref_env.SpatialKuramoto(d)
    rec = {"elec": [], "rec": [], "encaps": [], "w0": [], "init_state": [], "window_head": [],
           "elec_drift_episode": [], "encaps_episode": []}
    for r in range(n_resets):
        if r > 0:
            env.reset()
        rec["elec"].append(np.array(env.elec_coords)[0])
        rec["rec"].append(np.array(env.rec_coords)[0])
        rec["encaps"].append(env.encapsulation_coeff)
        rec["w0"].append(np.array(env.kuramoto.w0))
        rec["init_state"].append(np.array(env.init_state))
        rec["window_head"].append(np.array(env.theta_state[0][:8]))
        rec["elec_drift_episode"].append(env.elec_drift_episode)
        rec["encaps_episode"].append(env.elec_encaps_episode)
    out = {k: np.array(v) for k, v in rec.items()}
    for k in ("w0", "w0_without_locus", "locus_without_w0", "locus_mask"):
        out["in_" + k] = np.array(d[k]) if k != "w0_without_locus" else np.array(env.w0_without_locus_)
    out["overrides"] = np.array(["plasticity_drift_freq=1000000", "transient_state_len=117.5",
                                 "total_episode_len=9.0", "spatial_var_freq=4", "rand_seed=21(w0 seed)"])
    np.savez_compressed(os.path.join(HERE, "env2_events.npz"), **out)
    print("env2 events: elec", out["elec"].tolist(), "encaps", out["encaps"].tolist())


if __name__ == "__main__":
    which = sys.argv[1:] or ["schedule", "geometry", "env0", "env1", "env1d", "env2"]
    if "schedule" in which:
        record_schedule()
    if "geometry" in which:
        record_geometry()
    if "env0" in which:
        e0 = record_steps("step_env0.npz", "env0", seed=10, n_steps=70)
        record_rewards(e0)
    if "env1" in which:
        record_steps("step_env1.npz", "env1", seed=11, n_steps=12, action_seed=1)
    if "env1d" in which:
        record_steps("step_env1_directed.npz", "env1", seed=12, n_steps=6, action_seed=2,
                     directed_stimulation=True, elec_coords=[[5, 2, 3]], rec_coords=[[3, 5, 1]],
                     reward_func="temp_const_action")
    if "env2" in which:
        record_env2_events()
