"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes -> libdbsgym.so),
against the reference-generated golden fixtures and the CPU oracle.

Tolerances (BASELINE.json north_star): per-step phases within 1e-9 rad in fp64 mode and 1e-5 rad in
fp32 mode, teacher-forced (every step restarted from the reference state); integer results
(sample counts, schedule, done flags, RK step / RHS counters) bit-exact.
"""
import copy

import numpy as np
import pytest

from conftest import load_golden, make_params

pytestmark = pytest.mark.gpu

PHASE_TOL = {"f64": 1e-9, "f32": 1e-5}
LFP_TOL = {"f64": 1e-11, "f32": 2e-6}


def _engine_from_params(d, B, precision, force_dense=False, transfer="full", engine_options=None):
    # transfer="full": these tests overwrite device state between steps (teacher forcing), which the
    # host-side window mirror of the delta-transfer mode cannot see
    from dbsgym_b200.batched import BatchedKuramoto
    return BatchedKuramoto([copy.deepcopy(d) for _ in range(B)], precision=precision, force_dense=force_dense,
                           transfer=transfer, engine_options=engine_options)


def _teacher_forced(core, g, precision, n_steps, reward_tol_rel):
    eng = core.engine
    B = core.num_envs
    K = min(n_steps, len(g["actions"]))
    y_prev = g["y_after_transient"]
    win_prev = g["window0"]
    for k in range(K):
        eng.set_env_params(None, y0=np.tile(y_prev, (B, 1)))
        eng.set_window(np.tile(win_prev, (B, 1)))
        eng.set_episode(None, step_idx=k)
        obs, rew, done = core.step(np.full(B, g["actions"][k], dtype=np.float32))
        y = eng.state()
        t, r, n = eng.lfp()
        s = g["nI"][k] + g["nII"][k] - 1
        assert np.all(n == s)                                        # schedule / sample count: bit-exact
        err = np.max(np.abs(y - g["y_end"][k][None, :]))
        assert err < PHASE_TOL[precision], (k, err)
        assert np.max(np.abs(t[:, :s] - g["lfp_true"][k, :s])) < LFP_TOL[precision]
        assert np.max(np.abs(r[:, :s] - g["lfp_rec"][k, :s])) < LFP_TOL[precision]
        rr, uu = eng.rewards()
        assert np.all(uu == g["u"][k])                               # action rescale: exact in float64
        np.testing.assert_allclose(rr, g["reward"][k], rtol=reward_tol_rel, atol=1e-7)
        assert not done.any()
        y_prev = g["y_end"][k]
        # window after this step (reference): previous window shifted by s + new recorded samples
        win_prev = np.concatenate([win_prev, g["lfp_rec"][k, :s]])[-2340:]
        np.testing.assert_allclose(obs[0], win_prev.astype(np.float32), rtol=0,
                                   atol=1e-7 if precision == "f64" else 3e-6)
    c = eng.counters()
    assert c["status"] == 0
    return c


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_env0_teacher_forced_per_step_parity(precision):
    g = load_golden("step_env0.npz")
    d = make_params("env0", 10)
    core = _engine_from_params(d, 2, precision)
    core.engine.counters(reset=True)
    c = _teacher_forced(core, g, precision, 70, 1e-9 if precision == "f64" else 2e-4)
    # exactly 5 accepted sub-steps and 32 RHS evaluations per step, no rejections
    assert (c["accepted"], c["rejected"], c["rhs_evals"]) == (2 * 70 * 5, 0, 2 * 70 * 32)
    core.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_env1_weighted_recording_parity(precision):
    g = load_golden("step_env1.npz")
    d = make_params("env1", 11)
    core = _engine_from_params(d, 1, precision)
    _teacher_forced(core, g, precision, 12, 1e-9 if precision == "f64" else 2e-4)
    core.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_env1_directed_stimulation_and_temp_const_reward(precision):
    g = load_golden("step_env1_directed.npz")
    d = make_params("env1", 12, reward="temp_const_action", directed_stimulation=True,
                    elec_coords=[[5, 2, 3]], rec_coords=[[3, 5, 1]])
    core = _engine_from_params(d, 1, precision)
    _teacher_forced(core, g, precision, 6, 1e-7 if precision == "f64" else 5e-3)
    core.close()


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_reset_transient_matches_reference(precision):
    """Device transient (env.py:605-612): adaptive steps WITH rejections, 4000 dense-output samples."""
    g = load_golden("step_env0.npz")
    d = make_params("env0", 10)
    core = _engine_from_params(d, 1, precision)          # constructor == reference __init__ -> reset()
    assert np.array_equal(core.hosts[0].init_state, g["init_state"])
    c = core.engine.counters()
    keys = list(g["stats_keys"])
    ref = dict(zip(keys, g["reset_stats"]))
    assert c["status"] == 0
    if precision == "f64":
        assert (c["accepted"], c["rejected"], c["rhs_evals"]) == (
            ref["num_accepted_steps"], ref["num_rejected_steps"], ref["num_rhs_evals"])
    y = core.engine.state()[0]
    w = core.engine.window_values()[0]
    tol_y, tol_w = (1e-7, 1e-9) if precision == "f64" else (5e-3, 5e-4)
    assert np.max(np.abs(y - g["y_after_transient"])) < tol_y
    assert np.max(np.abs(w - g["window0"])) < tol_w
    obs = core.observations()
    np.testing.assert_allclose(obs[0], g["window0"].astype(np.float32), rtol=0, atol=tol_w + 1e-7)
    core.close()


def test_free_running_episode_matches_oracle_f64():
    """No teacher forcing: 40 consecutive steps from the transient on, against the CPU oracle."""
    from oracle import kuramoto_oracle as ko
    d = make_params("env1", 7)
    orc = ko.OracleEnv(copy.deepcopy(d))
    core = _engine_from_params(d, 1, "f64")
    acts = np.random.default_rng(3).uniform(-1, 1, 40).astype(np.float32)
    for k, a in enumerate(acts):
        o_ref, r_ref, d_ref, _, _ = orc.step(np.array([a], dtype=np.float32))
        obs, rew, done = core.step(np.array([a]))
        assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 1e-7, k
        assert abs(core.engine.rewards()[0][0] - r_ref) < 1e-6 * max(1.0, abs(r_ref))
        np.testing.assert_allclose(obs[0], o_ref[0], rtol=0, atol=1e-6)
    core.close()


def test_dense_fallback_agrees_with_grid_kernel():
    g = load_golden("step_env0.npz")
    d = make_params("env0", 10)
    outs = []
    for dense in (False, True):
        core = _engine_from_params(d, 1, "f64", force_dense=dense)
        assert core.engine.coupling == ("dense" if dense else "grid")
        core.engine.set_env_params(None, y0=g["y_after_transient"][None, :])
        core.step(np.array([g["actions"][0]]))
        outs.append(core.engine.state()[0])
        core.close()
    assert np.max(np.abs(outs[0] - g["y_end"][0])) < 1e-9
    assert np.max(np.abs(outs[1] - g["y_end"][0])) < 1e-9


def test_shuffled_grid_uses_dense_path_and_matches_oracle():
    """Coordinates that are not the regular grid (utils.py:490 shuffle=True) -> DENSE coupling."""
    from oracle import kuramoto_oracle as ko
    d = make_params("env0", 4, transient_state_len=118.0)
    perm = np.random.default_rng(1).permutation(512)
    d["neur_coords"] = d["neur_coords"][perm]
    d["neur_grid"] = d["neur_grid"][perm]
    orc = ko.OracleEnv(copy.deepcopy(d))
    core = _engine_from_params(d, 1, "f64")
    assert core.engine.coupling == "dense"
    assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 1e-7
    o_ref, r_ref, *_ = orc.step(np.array([0.4], dtype=np.float32))
    obs, rew, done = core.step(np.array([0.4]))
    assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 1e-7
    core.close()


def test_shuffled_grid_f32_runs_the_grid_kernels_through_the_order_map_and_the_low_rank_kernel():
    """Shuffled coordinates in float32 (utils.py:490 shuffle=True on the whole grid).  'auto': the library stores the
    oscillators in grid order (dbsgym_set_oscillator_order; the permutation is applied at the ABI boundary) and runs the
    spectral warp kernel of the regular grid (variant 10).  force_dense=True: the operator as a matrix, in its truncated
    eigenbasis (variant 11, 32 modes: a permutation does not change the spectrum).  coupling_eval='exact': the full matrix
    (variant 1).  All three: teacher-forced steps against the float64 oracle at the float32 tolerance, exact counters."""
    from oracle import kuramoto_oracle as ko
    d = make_params("env1", 4, transient_state_len=118.0)
    perm = np.random.default_rng(1).permutation(512)
    d["neur_coords"] = d["neur_coords"][perm]
    d["neur_grid"] = d["neur_grid"][perm]
    orc = ko.OracleEnv(copy.deepcopy(d))
    from dbsgym_b200.batched import BatchedKuramoto
    cores = {"grid_order": BatchedKuramoto([copy.deepcopy(d)] * 2, precision="f32", transfer="full"),
             "lowrank": BatchedKuramoto([copy.deepcopy(d)] * 2, precision="f32", transfer="full", force_dense=True),
             "dense": BatchedKuramoto([copy.deepcopy(d)] * 2, precision="f32", transfer="full", coupling_eval="exact")}
    assert cores["grid_order"].engine.step_variant() == 10 and cores["grid_order"].engine.coupling == "grid"
    assert np.array_equal(np.asarray(d["neur_grid"])[cores["grid_order"].engine.order], np.asarray(make_params("env1", 4)["neur_grid"]))
    assert cores["lowrank"].engine.step_variant() == 11 and cores["lowrank"].coupling_eval == "lowrank"
    assert cores["lowrank"].engine.lowrank["rank"] == 32
    assert cores["dense"].engine.step_variant() == 1
    for core in cores.values():
        assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 5e-3       # 118-unit transient with rejections
        core.engine.counters(reset=True)
    for k, a in enumerate((0.4, -0.7, 0.1, 0.9)):
        y_before = orc.sol_state[-1].copy()
        o_ref, r_ref, *_ = orc.step(np.array([a], dtype=np.float32))
        for name, core in cores.items():
            core.engine.set_env_params(None, y0=np.tile(y_before, (2, 1)))        # teacher-forced
            obs, rew, done = core.step(np.array([a, a], dtype=np.float32))
            err = np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1]))
            assert err < 1e-5, (name, k, err)
            assert np.max(np.abs(core.theta_records(0) - orc.theta_records)) < 2e-6
    for core in cores.values():
        c = core.engine.counters()
        assert (c["accepted"], c["rejected"], c["rhs_evals"], c["status"]) == (2 * 4 * 5, 0, 2 * 4 * 32, 0)
        core.close()


def test_large_regular_grid_selects_the_sector_low_rank_form_and_matches_oracle():
    """N = 2048 on a 16 x 16 x 8 grid through the public BatchedKuramoto: float32 'auto' takes the operator as sector-wise
    eigenpairs over the fundamental octant (step-kernel variants 13 / 11, oscillators stored in octant order inside the library) --
    reset transient and teacher-forced steps against the float64 oracle; coupling_eval='exact' keeps the structured kernel."""
    from oracle import kuramoto_oracle as ko
    import dbsgym_b200.utils as U
    from dbsgym_b200.batched import BatchedKuramoto
    np.random.seed(5)
    w0, nc, ng, w0t, wl, lm = U.generate_w0_with_locus(2048, [16, 16, 8], 0.1, [7, 7, 4], 0.55, 17, 1, show=False)
    d = make_params("env1", 5, transient_state_len=118.0, num_oscillators=2048, grid_size=[16, 16, 8], elec_coords=[[7, 6, 4]],
                    rec_coords=[[5, 8, 3]])
    d.update(w0=w0, w0_without_locus=w0t, locus_without_w0=wl, locus_mask=lm, neur_coords=nc, neur_grid=ng)
    orc = ko.OracleEnv(copy.deepcopy(d))
    core = BatchedKuramoto([copy.deepcopy(d)] * 2, precision="f32", transfer="full")
    assert core.engine.coupling == "grid" and core.coupling_eval == "lowrank"
    assert core.engine.step_variant() == 13 and core.engine.lowrank["sectors"]       # eigenvectors in registers (oct_kernel.cuh)
    block = BatchedKuramoto([copy.deepcopy(d)] * 2, precision="f32", transfer="full", engine_options={"no_warp_kernel": True})
    assert block.engine.step_variant() == 11 and block.engine.lowrank["sectors"]     # the block kernel streaming them from L2
    exact = BatchedKuramoto([copy.deepcopy(d)] * 2, precision="f32", transfer="full", coupling_eval="exact")
    assert exact.engine.step_variant() == 7
    for c in (core, block, exact):
        assert np.max(np.abs(c.engine.state()[0] - orc.sol_state[-1])) < 2e-3          # 118-unit transient with rejections
        c.engine.counters(reset=True)
    for k, a in enumerate((0.6, -0.3, 0.9)):
        y_before = orc.sol_state[-1].copy()
        o_ref, r_ref, *_ = orc.step(np.array([a], dtype=np.float32))
        for name, c in (("lowrank", core), ("lowrank_block", block), ("exact", exact)):
            c.engine.set_env_params(None, y0=np.tile(y_before, (2, 1)))                # teacher-forced
            obs, rew, done = c.step(np.array([a, a], dtype=np.float32))
            err = np.max(np.abs(c.engine.state()[0] - orc.sol_state[-1]))
            assert err < 1e-5, (name, k, err)
            assert np.max(np.abs(c.theta_records(0) - orc.theta_records)) < 2e-6
            assert np.max(np.abs(c.theta_mean(0) - orc.theta_mean)) < 2e-6
    for c in (core, block, exact):
        cc = c.engine.counters()
        assert (cc["accepted"], cc["rejected"], cc["rhs_evals"], cc["status"]) == (2 * 3 * 5, 0, 2 * 3 * 32, 0)
        c.close()
    # the same grid with its neurons shuffled (utils.py:490): order map (caller -> grid) composed with the octant order of the
    # sector form inside the library; everything the caller sees stays in the caller's order
    perm = np.random.default_rng(9).permutation(2048)
    ds = copy.deepcopy(d)
    for k in ("w0", "w0_without_locus", "locus_without_w0", "locus_mask", "neur_coords", "neur_grid"):
        ds[k] = np.asarray(ds[k])[perm]
    orc_s = ko.OracleEnv(copy.deepcopy(ds))
    shuf = BatchedKuramoto([copy.deepcopy(ds)] * 2, precision="f32", transfer="full")
    assert shuf.engine.coupling == "grid" and shuf.engine.order is not None and shuf.engine.step_variant() == 13
    assert shuf.engine.lowrank["sectors"]
    y_before = orc_s.sol_state[-1].copy()
    orc_s.step(np.array([0.5], dtype=np.float32))
    shuf.engine.set_env_params(None, y0=np.tile(y_before, (2, 1)))
    assert np.max(np.abs(shuf.engine.state()[0] - y_before)) < 4e-6                     # round trip through both permutations
    shuf.step(np.array([0.5, 0.5], dtype=np.float32))
    assert np.max(np.abs(shuf.engine.state()[0] - orc_s.sol_state[-1])) < 1e-5
    assert np.max(np.abs(shuf.theta_records(0) - orc_s.theta_records)) < 2e-6
    shuf.close()


def test_half_grid_256_oscillators():
    """N = 256 (first four z-planes, BASELINE config 5's smallest point) on the GRID kernel."""
    from oracle import kuramoto_oracle as ko
    import dbsgym_b200.utils as U
    np.random.seed(2)
    w0, nc, ng, w0t, wl, lm = U.generate_w0_with_locus(256, [8, 8, 8], 0.1, [2, 4, 4], 0.55, 17, 1, show=False)
    d = make_params("env0", 2, transient_state_len=118.0, num_oscillators=256, elec_coords=[[2, 3, 4]])
    d.update(w0=w0, w0_without_locus=w0t, locus_without_w0=wl, locus_mask=lm, neur_coords=nc, neur_grid=ng)
    orc = ko.OracleEnv(copy.deepcopy(d))
    core = _engine_from_params(d, 1, "f64")
    assert core.engine.coupling == "grid"
    assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 1e-7
    orc.step(np.array([-0.6], dtype=np.float32))
    core.step(np.array([-0.6]))
    assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 1e-7
    np.testing.assert_allclose(core.theta_mean(0), orc.theta_mean, rtol=0, atol=1e-10)
    core.close()
    # float32: the half-grid spectral kernel (one warp per environment, one octant point per lane; step-kernel variant 12, 26
    # modes) and the exact sector contraction, teacher-forced against the oracle; 35 environments = 3 CTAs with idle warps
    from dbsgym_b200.batched import BatchedKuramoto
    B = 35
    cores = {}
    for name, ceval, variant in (("spectral", "auto", 12), ("exact", "exact", 6)):
        c = BatchedKuramoto([copy.deepcopy(d) for _ in range(B)], precision="f32", transfer="full", coupling_eval=ceval)
        assert c.engine.step_variant() == variant, (name, c.engine.step_variant())
        c.engine.counters(reset=True)
        cores[name] = c
    assert cores["spectral"].engine.spectral["ranks"] == [5, 4, 4, 2, 4, 4, 2, 1]
    sample = [0, 15, 16, 34]
    for name, c in cores.items():
        orcs = {e: copy.deepcopy(orc) for e in sample}
        for k, a in enumerate((0.7, -0.2, 0.95, -1.0, 0.1)):
            y_before, w_before = c.engine.state(sample), c.engine.window_values(sample)
            for j, e in enumerate(sample):          # every environment ran its own float32 transient: restart the oracle from its state
                o = orcs[e]
                o.sol_state, o.theta_state = y_before[j][None, :].copy(), w_before[j][None, :].copy()
                o.current_step, o.current_time = int(c.current_step[e]), c.current_time(e)
            obs, rew, done = c.step(np.full(B, a, dtype=np.float32))
            y = c.engine.state(sample)
            for j, e in enumerate(sample):
                o = orcs[e]
                o_ref, r_ref, *_ = o.step(np.array([a], dtype=np.float32))
                err = np.max(np.abs(y[j] - o.sol_state[-1]))
                assert err < 1e-5, (name, k, e, err)
                assert np.max(np.abs(c.theta_records(e) - o.theta_records)) < 2e-6
                assert np.max(np.abs(c.theta_mean(e) - o.theta_mean)) < 2e-6
                assert rew[e] == pytest.approx(r_ref, rel=2e-4, abs=1e-7)
                np.testing.assert_allclose(obs[e], o_ref[0], rtol=0, atol=3e-6)
    for c in cores.values():
        cc = c.engine.counters()
        assert (cc["accepted"], cc["rejected"], cc["rhs_evals"], cc["status"]) == (B * 5 * 5, 0, B * 5 * 32, 0)
        c.close()


def test_plain_toeplitz_contraction_matches_symmetric_one():
    """CPL_GRID (engine option no_sym) and the default CPL_GRID_SYM kernel evaluate the same sum."""
    g = load_golden("step_env0.npz")
    d = make_params("env0", 10)
    outs = {}
    for nosym in ("0", "1"):
        for prec in ("f64", "f32"):
            core = _engine_from_params(d, 1, prec, engine_options={"no_sym": nosym == "1"})
            core.engine.set_env_params(None, y0=g["y_after_transient"][None, :])
            core.step(np.array([g["actions"][0]]))
            outs[(nosym, prec)] = core.engine.state()[0]
            core.close()
    for prec, tol in (("f64", 1e-9), ("f32", 1e-5)):
        assert np.max(np.abs(outs[("0", prec)] - g["y_end"][0])) < tol
        assert np.max(np.abs(outs[("1", prec)] - g["y_end"][0])) < tol
    assert np.max(np.abs(outs[("0", "f64")] - outs[("1", "f64")])) < 1e-11


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_threshold_reward_and_wavelet_kernel(precision):
    """R3 (env.py:669-688) and spatial_kernel='wavelet' (env.py:224-227) against the oracle."""
    from oracle import kuramoto_oracle as ko
    d = make_params("env1", 17, reward="bbpow_threth_action", spatial_kernel="wavelet", wavelet_amp=0.8,
                    wavelet_steepness=0.6, transient_state_len=118.0)
    orc = ko.OracleEnv(copy.deepcopy(d))
    core = _engine_from_params(d, 1, precision)
    assert core.engine.coupling == "grid"
    tol = 1e-7 if precision == "f64" else 5e-3
    assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < tol
    for a in (0.9, -0.3, 0.0):
        o_ref, r_ref, *_ = orc.step(np.array([a], dtype=np.float32))
        obs, rew, done = core.step(np.array([a]))
        assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < tol
        assert core.engine.rewards()[0][0] == pytest.approx(r_ref, abs=1e-9 if precision == "f64" else 1e-5)
    core.close()


def test_small_non_grid_system_dense_padding():
    """N = 100 oscillators on a 5x5x5 grid (not a multiple of 8: padded, DENSE coupling), naive DBS."""
    from oracle import kuramoto_oracle as ko
    import dbsgym_b200.utils as U
    np.random.seed(8)
    w0, nc, ng, w0t, wl, lm = U.generate_w0_with_locus(100, [5, 5, 5], 0.1, [2, 2, 2], 0.55, 17, 1, show=False)
    d = make_params("env0", 8, transient_state_len=118.0, num_oscillators=100, grid_size=[5, 5, 5],
                    elec_coords=[[2, 1, 2]], rec_coords=[[1, 1, 1]], naive_dbs=True, recording_kernel="gaussian")
    d.update(w0=w0, w0_without_locus=w0t, locus_without_w0=wl, locus_mask=lm, neur_coords=nc, neur_grid=ng)
    orc = ko.OracleEnv(copy.deepcopy(d))
    core = _engine_from_params(d, 2, "f64")
    assert core.engine.coupling == "dense"
    assert np.max(np.abs(core.engine.state()[0] - orc.sol_state[-1])) < 1e-7
    o_ref, r_ref, *_ = orc.step(np.array([0.5], dtype=np.float32))
    obs, rew, done = core.step(np.array([0.5, 0.5]))
    assert np.max(np.abs(core.engine.state() - orc.sol_state[-1][None, :])) < 1e-7
    np.testing.assert_allclose(core.theta_records(1), orc.theta_records, rtol=0, atol=1e-10)
    assert core.engine.rewards()[0][1] == pytest.approx(r_ref, rel=1e-8)
    core.close()
