"""Parity of the kernel and the configuration the BENCHMARK runs (BASELINE.json configs[2]: env1, 512 oscillators,
4096 environments, float32) -- pinned DIRECTLY to the reference-generated goldens and to the CPU oracle, not through
another GPU kernel.

"Oracle" here is always ``oracle/kuramoto_oracle.py`` over ``oracle/diffrax_restated.py``: the reference's own
``environment/env.py`` run verbatim agrees with it (tests/test_oracle_golden.py), but its integrator, diffrax 0.7.0, is a
restatement of the published algorithm (the package is not installable here) -- **parity unpinned** in that sense.

Stated tolerances (BASELINE.json north_star; every number below is asserted):
* per step, teacher-forced (each step restarted from the reference state), float32: phases within 1e-5 rad,
  LFP within 2e-6, RK sub-step / RHS counters exact, sample counts exact;
* free-running float32 against the float64 oracle from the same state, step k = 1..40: phases within
  2e-5 * (1 + k) rad (float32 rounding accumulates roughly linearly at these parameters: the dynamics are only weakly
  chaotic on this horizon, tests/test_gpu_episode_stats.py), LFP within 2e-6 * (1 + k);
* a full 2048-step float32 episode (one PPO rollout, README.md:66) against the float64 oracle: order parameter
  (theta_mean) trajectory within 5e-3 absolute over the whole episode and within 1e-4 over its first 200 steps,
  beta-band power of the true LFP (evaluate_HF_DBS.py:122-135) within 1 % relative, per-step rewards: mean within
  1 % and the two distributions within a Kolmogorov-Smirnov distance of 0.02.
"""
import copy

import numpy as np
import pytest

from conftest import load_golden, make_params

pytestmark = pytest.mark.gpu

MW_KERNELS = [("single", {"mw": False}, 3), ("multi_worker", {"mw": True}, 4)]


def _core(dicts, precision="f32", engine_options=None, transfer="full"):
    from dbsgym_b200.batched import BatchedKuramoto
    return BatchedKuramoto(copy.deepcopy(dicts), precision=precision, transfer=transfer, engine_options=engine_options)


@pytest.mark.parametrize("golden,cfg,seed,kw,n_steps", [
    ("step_env0.npz", "env0", 10, {}, 70),
    ("step_env1.npz", "env1", 11, {}, 12),
    ("step_env1_directed.npz", "env1", 12, dict(reward="temp_const_action", directed_stimulation=True,
                                                elec_coords=[[5, 2, 3]], rec_coords=[[3, 5, 1]]), 6)])
@pytest.mark.parametrize("name,options,variant", MW_KERNELS)
def test_f32_kernels_teacher_forced_against_reference_goldens(golden, cfg, seed, kw, n_steps, name, options, variant):
    """Both float32 step kernels for the 8 x 8 x 8 grid -- one CTA per environment (variant 3) and the multi-worker kernel
    the benchmark runs (variant 4) -- against the fixtures the reference's own env.py produced."""
    from test_gpu_parity import _teacher_forced
    g = load_golden(golden)
    d = make_params(cfg, seed, **kw)
    B = 3                                    # (3 of the 8 workers of a multi-worker CTA busy, 5 idle)
    core = _core([d] * B, engine_options=options)
    assert core.engine.step_variant() == variant
    core.engine.counters(reset=True)
    c = _teacher_forced(core, g, "f32", n_steps, 5e-3 if kw else 2e-4)
    K = min(n_steps, len(g["actions"]))
    assert (c["accepted"], c["rejected"], c["rhs_evals"]) == (B * K * 5, 0, B * K * 32)
    core.close()


def _oracle_for(core, e, d):
    """A CPU oracle environment carrying environment e's model (w0 after the host's non-positive fix, electrode)."""
    from oracle import kuramoto_oracle as ko
    st = np.random.get_state()
    orc = ko.OracleEnv(copy.deepcopy(d))
    np.random.set_state(st)
    orc.kuramoto.w0 = np.array(core.w0_model[e], dtype=np.float64)
    host = core.hosts[e]
    assert orc.elec_coords == host.elec_coords and orc.rec_coords == host.rec_coords
    return orc


def _sync_oracle(orc, core, e, y, win):
    orc.sol_state = y[None, :].copy()
    orc.theta_state = win[None, :].copy()
    orc.current_step = int(core.current_step[e])
    orc.current_time = core.current_time(e)


def test_bench_config_4096_env1_f32_against_oracle():
    """BASELINE configs[2] as bench.py builds it (env1, 4096 environments, float32, multi-worker kernel): a sample of
    environments is checked against the CPU oracle per step -- teacher-forced at 1e-5 rad, then free-running with the
    stated per-step tolerance -- plus exact counters for the whole batch."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import build_params
    B = 4096
    dicts = build_params(B, seed0=10)
    core = _core(dicts)
    eng = core.engine
    assert eng.step_variant() == 4
    sample = [0, 7, 1183, 1184, 2500, 4095]          # first / last worker slots, both sides of a wave boundary
    orcs = {e: _oracle_for(core, e, dicts[e]) for e in sample}
    rng = np.random.default_rng(5)
    eng.counters(reset=True)
    n_tf, n_free = 6, 40
    # ---- teacher-forced: the oracle restarts every step from the GPU's state before the step ----
    for k in range(n_tf):
        y_before = eng.state(sample)
        w_before = eng.window_values(sample)
        for j, e in enumerate(sample):
            _sync_oracle(orcs[e], core, e, y_before[j], w_before[j])
        acts = rng.uniform(-1, 1, B).astype(np.float32)
        obs, rew, done = core.step(acts)
        y_after = eng.state(sample)
        t, r, n = eng.lfp()
        rr, uu = eng.rewards()
        for j, e in enumerate(sample):
            orc = orcs[e]
            o_ref, r_ref, _, _, _ = orc.step(np.array([acts[e]], dtype=np.float32))
            s = len(orc.theta_mean)
            assert n[e] == s
            err = np.max(np.abs(y_after[j] - orc.sol_state[-1]))
            assert err < 1e-5, (k, e, err)
            assert np.max(np.abs(t[e, :s] - orc.theta_mean)) < 2e-6
            assert np.max(np.abs(r[e, :s] - orc.theta_records)) < 2e-6
            assert uu[e] == orc.u[0]
            assert rr[e] == pytest.approx(r_ref, rel=2e-4, abs=1e-7)
            np.testing.assert_allclose(obs[e], o_ref[0], rtol=0, atol=3e-6)
    c = eng.counters()
    assert c["status"] == 0
    assert (c["accepted"], c["rejected"], c["rhs_evals"]) == (B * n_tf * 5, 0, B * n_tf * 32)
    # ---- free-running: both sides continue from the same state without re-synchronisation ----
    y0 = eng.state(sample)
    w0 = eng.window_values(sample)
    for j, e in enumerate(sample):
        _sync_oracle(orcs[e], core, e, y0[j], w0[j])
    worst = 0.0
    for k in range(1, n_free + 1):
        acts = rng.uniform(-1, 1, B).astype(np.float32)
        core.step(acts)
        y = eng.state(sample)
        t, r, n = eng.lfp()
        for j, e in enumerate(sample):
            orc = orcs[e]
            orc.step(np.array([acts[e]], dtype=np.float32))
            err = np.max(np.abs(y[j] - orc.sol_state[-1]))
            worst = max(worst, err / (1 + k))
            assert err < 2e-5 * (1 + k), (k, e, err)
            s = len(orc.theta_mean)
            assert np.max(np.abs(r[e, :s] - orc.theta_records)) < 2e-6 * (1 + k)
    print(f"free-running f32 vs f64 oracle: worst phase error / (1 + k) = {worst:.2e} rad")
    assert eng.counters()["status"] == 0
    core.close()


def test_env2_reset_parity_f32():
    """env2 (temporal drift: electrode movement, encapsulation, plasticity walk; compat_env2 fixes SURVEY.md F7) in float32:
    after every reset the host-side event bookkeeping must equal the oracle's exactly and the device transient must
    reproduce the oracle's reset state (window 5e-4, phases 5e-3 rad over the 118-unit transient with rejections)."""
    from oracle import kuramoto_oracle as ko
    from dbsgym_b200.batched import BatchedKuramoto
    d = make_params("env2", 31, transient_state_len=118.0, total_episode_len=2.7)
    np.random.seed(123)
    orc = ko.OracleEnv(copy.deepcopy(d), compat_env2=True)
    np.random.seed(123)
    core = BatchedKuramoto([copy.deepcopy(d)], precision="f32", compat_env2=True)
    host = core.hosts[0]
    for ep in range(9):
        assert host.elec_coords == orc.elec_coords and host.rec_coords == orc.rec_coords
        assert host.encapsulation_coeff == orc.encapsulation_coeff
        assert host.plasticity_process_count == orc.plasticity_process_count
        assert np.array_equal(np.asarray(host.init_state), orc.init_state)
        assert np.array_equal(np.asarray(core.w0_model[0]), orc.kuramoto.w0)
        y = core.engine.state()[0]
        w = core.engine.window_values()[0]
        assert np.max(np.abs(y - orc.sol_state[-1])) < 5e-3, ep
        assert np.max(np.abs(w - orc.theta_state[0])) < 5e-4, ep
        for a in (0.7, -0.2):
            o_ref, r_ref, *_ = orc.step(np.array([a], dtype=np.float32))
            obs, rew, done = core.step(np.array([a], dtype=np.float32))
            assert np.max(np.abs(obs[0] - o_ref[0])) < 5e-4
        st = np.random.get_state()                    # both draw from the global stream: replay it for the second one
        orc.reset()
        st_after = np.random.get_state()
        np.random.set_state(st)
        core.reset_envs([0])
        now = np.random.get_state()
        assert np.array_equal(now[1], st_after[1]) and now[2:] == st_after[2:]     # same number of draws consumed
    assert core.engine.counters()["status"] == 0
    core.close()


def _ks_distance(a, b):
    a, b = np.sort(a), np.sort(b)
    grid = np.concatenate([a, b])
    return float(np.max(np.abs(np.searchsorted(a, grid, side="right") / a.size -
                               np.searchsorted(b, grid, side="right") / b.size)))


def test_full_2048_step_episode_f32_against_oracle():
    """One whole PPO rollout (2048 steps, README.md:66) of an env0 environment, float32 GPU against the float64 oracle:
    order-parameter trajectory, beta-band power and the reward distribution within the tolerances stated at the top."""
    from oracle import kuramoto_oracle as ko
    from test_gpu_episode_stats import eval_bbpow
    n = 2048
    d = make_params("env0", 10)
    acts = np.random.default_rng(0).uniform(-1, 1, n).astype(np.float32)
    orc = ko.OracleEnv(copy.deepcopy(d))
    tm_ref, rew_ref = [], np.zeros(n)
    for k, a in enumerate(acts):
        _, rew_ref[k], *_ = orc.step(np.array([a], dtype=np.float32))
        tm_ref.append(orc.theta_mean.copy())
    core = _core([d] * 2, engine_options={"mw": True})
    assert core.engine.step_variant() == 4
    tm, rew = [], np.zeros(n)
    for k, a in enumerate(acts):
        core.step(np.full(2, a, dtype=np.float32))
        t, _, ns = core.engine.lfp()
        tm.append(t[0, :ns[0]].copy())
        rew[k] = core.engine.rewards()[0][0]
    st = core.engine.counters()
    core.close()
    assert st["status"] == 0 and st["rejected"] == 0 and st["accepted"] == 2 * n * 5
    assert [len(x) for x in tm] == [len(x) for x in tm_ref]                 # sample schedule: exact
    x, x_ref = np.concatenate(tm), np.concatenate(tm_ref)
    err = np.abs(x - x_ref)
    n200 = sum(len(v) for v in tm[:200])
    print(f"2048-step f32 episode: order-parameter error first 200 steps {err[:n200].max():.2e}, whole episode {err.max():.2e}")
    assert err[:n200].max() < 1e-4
    assert err.max() < 5e-3
    bb, bb_ref = eval_bbpow(x), eval_bbpow(x_ref)
    assert abs(bb - bb_ref) / bb_ref < 1e-2
    assert abs(rew.mean() - rew_ref.mean()) < 1e-2 * abs(rew_ref.mean())
    ks = _ks_distance(rew, rew_ref)
    print(f"beta power gpu {bb:.6e} oracle {bb_ref:.6e}; reward mean gpu {rew.mean():.5f} oracle {rew_ref.mean():.5f}; KS {ks:.4f}")
    assert ks < 0.02
