"""The observation tail fused into the step kernel: ring append + incrementally updated rfft bins (beta-power
rewards R1 / R3, reference env.py:447-454, :638-650, :669-688, utils.py:21-27) against a from-scratch numpy rfft
of the device window, and against the separate observation kernel (engine option no_fused_obs)."""
import numpy as np
import pytest

from conftest import load_golden, make_params

pytestmark = pytest.mark.gpu

W = 2340


def _engine(B, precision, reward="bbpow_action", options=None):
    from dbsgym_b200.batched import BatchedKuramoto
    from dbsgym_b200.engine import KuramotoEngine
    from dbsgym_b200.geometry import coupling_table
    d = make_params("env1", 10)
    core = BatchedKuramoto([d], precision=precision)
    table = coupling_table(d["neur_coords"], d["neur_grid"], d["grid_size"], "cos")
    eng = KuramotoEngine(B, 512, [8, 8, 8], W, d["K"], precision=precision, coupling_table=table, options=options)
    eng.set_schedule(core.schedule)
    eng.set_recording(True)
    eng.set_reward(reward, 0.05)
    sched = core.schedule
    core.close()
    g = load_golden("step_env0.npz")
    rng = np.random.default_rng(3)
    w0 = np.abs(rng.normal(1.0, 0.4, (B, 512))) + 0.05
    stim = np.tile(g["stim_cond"][0], (B, 1))
    eng.set_env_params(None, w0=w0, stim=stim, rec=stim, y0=rng.normal(np.pi, 0.6, (B, 512)))
    eng.set_window(rng.uniform(-0.3, 0.3, (B, W)))
    eng.set_episode(None, step_idx=0, episode_len=2 ** 30)
    return eng, sched, rng


def _reward_from_window(win, u, kind="bbpow_action"):
    """env.py:638-650 / 669-688 on the chronological window (float64)."""
    from dbsgym_b200.utils import beta_bins, units2sec
    lo, hi = beta_bins(W, units2sec(0.05), 12.5, 21)
    X = np.fft.rfft(win, axis=-1) / W
    pw = (2 * np.abs(X[:, lo:hi + 1]) ** 2).sum(axis=1)
    if kind == "bbpow_action":
        return -1e4 * pw - 1e-2 * np.abs(u)
    return -5.0 * (1e4 * pw > 20.0) - np.abs(u)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_incremental_bins_track_the_window_over_many_steps(precision):
    """400 steps replace the whole 2340-sample window three times over; the running bins must still give the
    reward a from-scratch rfft of the device window gives (no drift), also after a partial reset transient."""
    B = 6
    eng, sched, rng = _engine(B, precision)
    for k in range(400):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        eng.step_host(a)
        if k % 57 == 0 or k == 399:
            r, u = eng.rewards()
            np.testing.assert_allclose(r, _reward_from_window(eng.window_values(), u), rtol=1e-9, atol=1e-12)
    # partial reset: environments 1 and 4 run the transient (ring rewritten, bins re-initialised from it)
    ts = np.arange(0.0, 118.0, 0.05)
    eng.transient(ts, env_ids=[1, 4])
    eng.set_episode([1, 4], step_idx=0, episode_len=2 ** 30)
    for k in range(3):
        eng.step_host(rng.uniform(-1, 1, B).astype(np.float32))
    r, u = eng.rewards()
    np.testing.assert_allclose(r, _reward_from_window(eng.window_values(), u), rtol=1e-9, atol=1e-12)
    assert eng.counters()["status"] == 0
    eng.close()


@pytest.mark.parametrize("reward", ["bbpow_action", "bbpow_threth_action"])
def test_fused_tail_equals_separate_observation_kernel(reward):
    """Same inputs through both paths: identical phases, windows, observations, done flags and step
    counters; rewards agree to the float32 DFT accuracy of the separate kernel."""
    B, n = 5, 12
    outs = {}
    for mode in ("fused", "separate"):
        eng, sched, rng = _engine(B, "f32", reward, options={"no_fused_obs": mode == "separate"})
        eng.set_episode(None, step_idx=0, episode_len=n - 2)
        rec = []
        for k in range(n):
            obs, rew, done = eng.step_host(rng.uniform(-1, 1, B).astype(np.float32))
            rec.append((obs.copy(), rew.copy(), done.copy()))
        outs[mode] = (rec, eng.state(), eng.window_values(), eng.episode()[0], eng.lfp()[2])
        eng.close()
    (ra, ya, wa, sa, na), (rb, yb, wb, sb, nb) = outs["fused"], outs["separate"]
    assert np.array_equal(ya, yb) and np.array_equal(wa, wb) and np.array_equal(sa, sb) and np.array_equal(na, nb)
    for (oa, rwa, da), (ob, rwb, db) in zip(ra, rb):
        assert np.array_equal(oa, ob) and np.array_equal(da, db)
        np.testing.assert_allclose(rwa, rwb, rtol=2e-4, atol=1e-6)
    assert ra[-1][2].all() and not ra[0][2].any()


def test_host_mirror_and_sample_transfer_with_fused_tail():
    """The pinned host mirror written by the step kernel's tail holds the chronological windows."""
    B = 4
    eng, sched, rng = _engine(B, "f32")
    mirror = eng.host_mirror()
    rew = np.empty(B, dtype=np.float32); done = np.empty(B, dtype=np.uint8)
    for k in range(5):
        a = rng.uniform(-1, 1, B).astype(np.float32)
        pos, n_new = eng.step_host_mirror(a, rew, done)
        assert n_new == sched.n_I[k] + sched.n_II[k] - 1
        np.testing.assert_array_equal(mirror[:, pos:pos + W], eng.window_values().astype(np.float32))
    full = eng.obs_host()
    np.testing.assert_array_equal(full, mirror[:, pos:pos + W])
    eng.close()


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_device_evaluation_metric_matches_scipy_on_the_recorded_trace(precision):
    """The evaluation trace recorded by the step kernel's tail is the concatenation of the per-step TRUE-LFP
    samples, and the device kernels (odd-extension filtfilt + weighted band power) reproduce
    calc_psd_for_simple_eval (aDBS_RL/evaluate_HF_DBS.py:122-135, scipy on the host) to 1e-9."""
    from dbsgym_b200.evaluation import calc_psd_for_simple_eval, device_bbpow
    B, n_steps = 7, 120
    eng, sched, rng = _engine(B, precision)
    eng.trace_begin(n_steps * eng.max_step_samples)
    host = [[] for _ in range(B)]
    for k in range(n_steps):
        eng.step_host(rng.uniform(-1, 1, B).astype(np.float32))
        t, _, n = eng.lfp()
        for b in range(B):
            host[b].append(t[b, :n[b]].copy())
    eng.trace_end()
    eng.step_host(rng.uniform(-1, 1, B).astype(np.float32))          # not recorded any more
    tr, ln = eng.trace()
    sig = np.stack([np.concatenate(h) for h in host])
    assert np.all(ln == sig.shape[1])
    assert np.array_equal(tr[:, :ln[0]], sig)
    ref = calc_psd_for_simple_eval(sig, 0.0005)
    np.testing.assert_allclose(device_bbpow(eng), ref, rtol=1e-9)
    eng.close()


def test_fsal_reuse_across_segments_and_launches():
    """fp32 mode takes the first stage of every segment from the last stage of the previous accepted sub-step (same
    state, only the pulse term changes) instead of evaluating the RHS again -- also across kernel launches and after
    a reset transient.  The counters keep the reference's (logical) count of 32 evaluations per step; 2 per step
    are reused (1 in the very first step after the vectors were uploaded); phases, windows and rewards agree with the
    kernel that evaluates every stage (engine option no_fsal_reuse) to float32 rounding."""
    B, n = 6, 8
    outs = {}
    for mode in ("reuse", "evaluate"):
        eng, sched, rng = _engine(B, "f32", options={"no_fsal_reuse": mode == "evaluate"})
        eng.counters(reset=True)
        for k in range(n):
            eng.step_host(rng.uniform(-1, 1, B).astype(np.float32))
        c1, r1 = eng.counters(), eng.rhs_reused()
        eng.transient(np.arange(0.0, 118.0, 0.05), env_ids=[2, 3])
        eng.set_episode([2, 3], step_idx=0, episode_len=2 ** 30)
        eng.counters(reset=True)
        obs, rew, done = eng.step_host(rng.uniform(-1, 1, B).astype(np.float32))
        c2, r2 = eng.counters(), eng.rhs_reused()
        outs[mode] = (c1, r1, c2, r2, eng.state(), eng.window_values(), rew.copy())
        eng.close()
    (c1, r1, c2, r2, y, w, rew), (d1, s1, d2, s2, y0, w0, rew0) = outs["reuse"], outs["evaluate"]
    assert c1 == d1 and c2 == d2 and c1["rhs_evals"] == B * n * 32 and c1["status"] == 0
    assert (s1, s2) == (0, 0)
    assert r1 == B * (2 * n - 1)            # the first segment after set_env_params has nothing to reuse
    assert r2 == B * 2                      # the transient of environments 2 and 3 left its last stage behind as well
    # nine free-running steps: the one-ulp differences of the carried stage grow like any fp32 rounding difference
    # (tests/test_gpu_episode_stats.py measures that horizon); per-step accuracy against the fp64 oracle is what the
    # teacher-forced parity tests in tests/test_gpu_parity.py check, with the reuse active
    np.testing.assert_allclose(y, y0, rtol=0, atol=1e-3)
    np.testing.assert_allclose(w, w0, rtol=0, atol=2e-5)
    np.testing.assert_allclose(rew, rew0, rtol=1e-3, atol=1e-5)
