"""The oracle (oracle/kuramoto_oracle.py) against the fixtures produced by the reference's own
environment/env.py (tests/golden/make_golden.py).  CPU only."""
import copy
import os

import numpy as np
import pytest

from conftest import load_golden, make_params
from oracle import kuramoto_oracle as ko
from oracle.diffrax_restated import Dopri5, ODETerm, PIDController, SaveAt, diffeqsolve


def _oracle_env(cfg_name, seed, g, **over):
    d = make_params(cfg_name, seed, **over)
    # the product's w0 generator must reproduce the golden inputs bit for bit
    for k in ("w0", "w0_without_locus", "locus_without_w0", "locus_mask"):
        assert np.array_equal(d[k], g[k]), k
    return ko.OracleEnv(d)


@pytest.mark.parametrize("name,cfg,seed,aseed,over", [
    ("step_env0.npz", "env0", 10, 0, {}),
    ("step_env1.npz", "env1", 11, 1, {}),
    ("step_env1_directed.npz", "env1", 12, 2, dict(directed_stimulation=True, elec_coords=[[5, 2, 3]],
                                                   rec_coords=[[3, 5, 1]], reward="temp_const_action")),
])
def test_oracle_reproduces_reference_steps(name, cfg, seed, aseed, over):
    g = load_golden(name)
    reward = over.pop("reward", "bbpow_action") if "reward" in over else "bbpow_action"
    env = _oracle_env(cfg, seed, g, reward_func=reward, **over)
    assert np.array_equal(env.init_state, g["init_state"])          # RNG draw order (Appendix C)
    assert np.array_equal(env.kuramoto.w0, g["w0_model"])
    assert np.array_equal(np.array(env.kuramoto.dbs.elec_idxs), g["elec_idxs"])
    assert np.array_equal(np.array(env.kuramoto.dbs.rec_idxs), g["rec_idxs"])
    assert np.array_equal(np.array(env.kuramoto.dbs.conductances), g["stim_cond"])
    assert np.array_equal(np.array(env.kuramoto.dbs.rec_conductances), g["rec_cond"])
    assert env.current_time == g["t_after_transient"]
    np.testing.assert_allclose(env.sol_state[-1], g["y_after_transient"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(env.theta_state[0], g["window0"], rtol=0, atol=1e-11)
    stats = env.kuramoto.stats
    assert [stats[k] for k in sorted(stats)] == list(g["reset_stats"])
    n = min(10, len(g["actions"]))
    for k in range(n):
        obs, r, done, trunc, info = env.step(np.array([g["actions"][k]], dtype=np.float32))
        assert (len(env.t_eval_step_I), len(env.t_eval_step_II)) == (g["nI"][k], g["nII"][k])
        assert env.current_time == g["t_cur"][k]
        s = g["nI"][k] + g["nII"][k] - 1
        np.testing.assert_allclose(env.sol_state[-1], g["y_end"][k], rtol=0, atol=1e-9)
        np.testing.assert_allclose(env.theta_mean, g["lfp_true"][k, :s], rtol=0, atol=1e-11)
        np.testing.assert_allclose(env.theta_records, g["lfp_rec"][k, :s], rtol=0, atol=1e-11)
        assert r == pytest.approx(g["reward"][k], rel=1e-8, abs=1e-9)
        assert env.u[0] == g["u"][k]
        assert not done and trunc is False and info == {}
        assert obs.dtype == np.float32 and obs.shape == (1, 2340)


def test_oracle_rewards_match_reference():
    g = load_golden("rewards.npz")
    for w, u, r, bb in zip(g["windows"], g["u"], g["rewards"], g["bbpow"]):
        assert ko.beta_band_power(w, 0.0005, 12.5, 21) == pytest.approx(bb, rel=1e-13)
        assert ko.reward_bbpow_action(w, [u], 0.05) == pytest.approx(r[0], rel=1e-12)
        assert ko.reward_temp_const(w, [u], 0.05) == pytest.approx(r[1], rel=1e-10)
        assert ko.reward_bbpow_threshold(w, [u], 0.05) == pytest.approx(r[2], rel=1e-12)


def test_oracle_schedule_matches_reference():
    g = load_golden("schedule.npz")
    t0 = np.arange(0., 200., 0.05)[-1]
    for tag, n in (("train", 5555), ("eval", 1111)):
        sch = ko.step_schedule(n, t0, 0.15, 0.75, 0.05)
        assert np.array_equal([len(a) for a, _ in sch], g[f"nI_{tag}"])
        assert np.array_equal([len(b) for _, b in sch], g[f"nII_{tag}"])
        assert sch[-1][1][-1] == g[f"t_final_{tag}"]
    # SURVEY.md Appendix B known answers
    nI, nII = g["nI_train"], g["nII_train"]
    assert ((nI == 4) & (nII == 15)).sum() == 3678 and ((nI == 3) & (nII == 15)).sum() == 1875
    assert ((nI == 4) & (nII == 16)).sum() == 2 and nII[65] == 16


def test_oracle_geometry_matches_reference():
    g = load_golden("geometry.npz")
    _, grid = ko.neuron_grid_3d(8, 8, 8, 512, 0.1)
    assert np.array_equal(grid, g["neur_grid"])
    coords = grid * 0.1
    assert np.array_equal(ko.coupling_matrix(coords, "cos")[0], g["alpha_row0"])
    assert np.array_equal(ko.coupling_matrix(coords, "cos")[284], g["alpha_row284"])
    assert np.array_equal(ko.oscillation_locus(grid, [8, 8, 8], [4, 4, 4], 0.55), g["locus_mask_444_055"])
    assert g["locus_mask_444_055"].sum() == 27
    for i in range(5):
        el = ko.Electrode([8, 8, 8], grid, float(g[f"case{i}_cm"]), g[f"case{i}_elec"].tolist(),
                          g[f"case{i}_rec"].tolist(), [0.], bool(g[f"case{i}_directed"]), "dummy")
        assert np.array_equal(el.elec_idxs, g[f"case{i}_elec_idx"])
        assert np.array_equal(el.rec_idxs, g[f"case{i}_rec_idx"])
        assert np.array_equal(el.conductances[0], g[f"case{i}_cond"])
        assert np.array_equal(el.rec_conductances[0], g[f"case{i}_rec_cond"])
        if g[f"case{i}_directed"]:
            assert np.array_equal(np.array(el.directional_masks[0]), g[f"case{i}_masks"])
    assert contact_known_answer()
    for n, gs in ((256, (8, 8, 8)), (100, (5, 5, 5))):
        assert np.array_equal(ko.neuron_grid_3d(*gs, n, 0.1)[1], g[f"grid_{n}_{gs[0]}"])


def contact_known_answer():
    # contact [4,3,4] -> index 284 -> neuron [3,4,4]
    _, grid = ko.neuron_grid_3d(8, 8, 8, 512, 0.1)
    idx = ko.contact_index([4, 3, 4], [8, 8, 8])
    return idx == 284 and grid[idx].tolist() == [3, 4, 4]


def test_oracle_env2_event_sequence():
    """Drift logic + RNG order under temporal drift (env.py:483-557), against the reference run."""
    g = load_golden("env2_events.npz")
    d = make_params("env2", 21, plasticity_drift_freq=10 ** 6, transient_state_len=117.5,
                    total_episode_len=9., spatial_var_freq=4)
    assert np.array_equal(d["w0"], g["in_w0"])
    env = ko.OracleEnv(d)
    for r in range(len(g["elec"])):
        if r > 0:
            env.reset()
        assert np.array_equal(np.array(env.elec_coords)[0], g["elec"][r]), r
        assert np.array_equal(np.array(env.rec_coords)[0], g["rec"][r]), r
        assert env.encapsulation_coeff == g["encaps"][r]
        assert np.array_equal(env.kuramoto.w0, g["w0"][r])
        assert np.array_equal(env.init_state, g["init_state"][r])
        np.testing.assert_allclose(env.theta_state[0][:8], g["window_head"][r], rtol=0, atol=1e-10)
        assert env.elec_drift_episode == g["elec_drift_episode"][r]
        assert env.elec_encaps_episode == g["encaps_episode"][r]


def test_env2_as_shipped_fails_like_the_reference():
    d = make_params("env2", 3)
    with pytest.raises(AssertionError):
        ko.OracleEnv(d)                                             # env.py:368 (plasticity_drift_freq: 1)


def test_dopri5_against_scipy_dop853():
    """Independent check of the solver restatement: same ODE with scipy at tight tolerance."""
    from scipy.integrate import solve_ivp
    rng = np.random.default_rng(0)
    n = 24
    w = rng.uniform(0.5, 2.0, n)
    A = np.cos(rng.uniform(0, 1, (n, n)))
    A = 0.5 * (A + A.T)
    y0 = rng.normal(np.pi, 0.6, n)
    f = lambda t, y, a=None: ko.kuramoto_rhs(y, w, 0.52 / n, A, np.zeros(n))   # noqa: E731
    ts = np.arange(0., 20., 0.05)
    sol = diffeqsolve(ODETerm(f), Dopri5(), t0=ts[0], t1=ts[-1], dt0=0.05, y0=y0, saveat=SaveAt(ts=ts),
                      stepsize_controller=PIDController(rtol=1e-5, atol=1e-5))
    ref = solve_ivp(lambda t, y: f(t, y), (ts[0], ts[-1]), y0, method="DOP853", t_eval=ts, rtol=1e-12, atol=1e-12)
    assert np.max(np.abs(sol.ys - ref.y.T)) < 5e-3        # rtol=atol=1e-5: global error of that order
    tight = diffeqsolve(ODETerm(f), Dopri5(), t0=ts[0], t1=ts[-1], dt0=0.05, y0=y0, saveat=SaveAt(ts=ts),
                        stepsize_controller=PIDController(rtol=1e-10, atol=1e-10))
    assert np.max(np.abs(tight.ys - ref.y.T)) < 1e-6      # converges to the true solution (5th order + interpolant)
    # as-written and mat-vec right-hand sides agree
    y = rng.normal(3, 2, n)
    a = ko.kuramoto_rhs(y, w, 0.02, A, w * 0.1, "as_written")
    b = ko.kuramoto_rhs(y, w, 0.02, A, w * 0.1, "matvec")
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-14)


@pytest.mark.skipif(not os.path.isdir("/root/reference/environment"), reason="reference checkout not present")
def test_oracle_matches_live_reference():
    """Where the reference checkout exists: run it verbatim under the shims beside the oracle."""
    from oracle.run_reference import load_reference
    ref_env, ref_utils, ref_cfgs = load_reference()
    d = make_params("env1", 5, transient_state_len=118.0)
    e_ref = ref_env.SpatialKuramoto(copy.deepcopy(d))
    e_or = ko.OracleEnv(copy.deepcopy(d))
    np.testing.assert_allclose(e_or.sol_state[-1], e_ref.sol_state[-1], rtol=0, atol=1e-9)
    for a in (0.7, -0.2):
        o1 = e_ref.step(np.array([a], dtype=np.float32))
        o2 = e_or.step(np.array([a], dtype=np.float32))
        np.testing.assert_allclose(e_or.sol_state_, e_ref.sol_state_, rtol=0, atol=1e-9)
        assert o2[1] == pytest.approx(o1[1], rel=1e-9)
        assert np.array_equal(o1[0], o2[0])
