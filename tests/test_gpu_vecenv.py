"""Boundary tests on the GPU: the gymnasium env, the batched VecEnv and size-independent properties
at BASELINE.json's full batch size (4096 environments)."""
import copy

import numpy as np
import pytest

from conftest import load_golden, make_params

pytestmark = pytest.mark.gpu


def test_spatial_kuramoto_gym_api_matches_reference_golden():
    from environment.env import SpatialKuramoto
    g = load_golden("step_env0.npz")
    env = SpatialKuramoto(make_params("env0", 10, precision="f64"))
    assert env.observation_space.shape == (1, 2340) and env.action_space.shape == (1,)
    assert env.total_episode_counts == 5555 and env.observe_wind_idxs == 2340
    assert env.current_time == g["t_after_transient"]
    np.testing.assert_allclose(env.theta_state[0], g["window0"], rtol=0, atol=1e-9)
    for k in range(5):
        obs, r, done, trunc, info = env.step(np.array([g["actions"][k]], dtype=np.float32))
        assert obs.shape == (1, 2340) and obs.dtype == np.float32
        assert isinstance(r, float) and done is False and trunc is False and info == {}
        s = g["nI"][k] + g["nII"][k] - 1
        assert env.theta_mean.shape == (s,) and env.current_time == g["t_cur"][k]
        assert env.u == [g["u"][k]]
        np.testing.assert_allclose(env.theta_mean, g["lfp_true"][k, :s], rtol=0, atol=1e-8)
        np.testing.assert_allclose(env.sol_state[-1], g["y_end"][k], rtol=0, atol=1e-7)
        assert r == pytest.approx(g["reward"][k], rel=1e-6)
    assert env.kuramoto.dbs.conductances[0].shape == (512,) and env.kuramoto.neur_grid.shape == (512, 3)
    obs, info = env.reset()
    assert obs.shape == (1, 2340) and info == {} and env.current_step == 0 and env.reset_count == 1
    env.close()


def test_constructor_errors_match_reference():
    from environment.env import SpatialKuramoto
    with pytest.raises(ValueError, match="Wrong reward function"):
        SpatialKuramoto(make_params("env0", 1, reward="x"))
    with pytest.raises(ValueError, match="Wrong recording kernel"):
        SpatialKuramoto(make_params("env0", 1, recording_kernel="x"))
    with pytest.raises(ValueError, match="Transient state"):
        SpatialKuramoto(make_params("env0", 1, transient_state_len=50.))
    with pytest.raises(AssertionError):
        SpatialKuramoto(make_params("env0", 1, electrode_amps=[0., 0.]))


def test_vecenv_semantics_autoreset_monitor_and_rng_order():
    """Short episodes: done flags, terminal_observation, Monitor-style episode info, and the
    sequential-DummyVecEnv RNG order of the initial phases across environments and resets."""
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
    from dbsgym_b200.host_env import HostEnvState
    B = 3
    dicts = [make_params("env1", 20 + e, transient_state_len=118.0, total_episode_len=3.6, rand_seed=50 + e,
                         precision="f64") for e in range(B)]
    # expected draws: replay the host logic of a sequential DummyVecEnv([...]) with the same dicts
    ref_hosts, ref_y0 = [], []
    for d in copy.deepcopy(dicts):
        h = HostEnvState(d); ref_hosts.append(h); ref_y0.append(h.begin_episode().y0)
    venv = BatchedKuramotoVecEnv(dicts)
    assert venv.num_envs == B and venv.observation_space.shape == (1, 2340)
    for i in range(B):
        assert np.array_equal(venv.core.hosts[i].init_state, ref_y0[i])
    # both share the global np.random stream: replay the expected draws, rewind, then let the VecEnv draw
    st = np.random.get_state()
    ref_y0 = [h.begin_episode().y0 for h in ref_hosts]
    np.random.set_state(st)
    obs = venv.reset()                                   # SB3 calls reset() after construction
    for i in range(B):
        assert np.array_equal(venv.core.hosts[i].init_state, ref_y0[i])
    assert obs.shape == (B, 1, 2340) and obs.dtype == np.float32
    total = venv.get_attr("total_episode_counts")[0]
    assert total == 4
    rets = np.zeros(B)
    for k in range(total):
        if k == total - 1:                               # the auto-reset inside this step_wait draws, in index order
            st = np.random.get_state()
            ref_y0 = [h.begin_episode().y0 for h in ref_hosts]
            np.random.set_state(st)
        venv.step_async(np.full((B, 1), 0.5, dtype=np.float32))
        obs2, rew, done, infos = venv.step_wait()
        rets += rew
        assert rew.shape == (B,) and done.dtype == np.bool_ and len(infos) == B
        tm = venv.get_attr("theta_mean")
        if k < total - 1:
            assert not done.any() and "terminal_observation" not in infos[0]
            assert len(tm[0]) in (17, 18, 19) and venv.get_attr("u")[1] == [2.5]
    assert done.all()
    for i in range(B):
        assert infos[i]["TimeLimit.truncated"] is False
        assert infos[i]["terminal_observation"].shape == (1, 2340)
        assert infos[i]["episode"]["l"] == total
        assert infos[i]["episode"]["r"] == pytest.approx(rets[i], rel=1e-5)
        assert np.array_equal(venv.core.hosts[i].init_state, ref_y0[i])
        assert not np.array_equal(infos[i]["terminal_observation"], obs2[i])      # obs is the fresh episode's
    assert venv.get_attr("current_step", [0, 2])[1] == 0
    assert venv.get_attr("params_dict")[1] is dicts[1]
    assert venv.env_is_wrapped(type("Monitor", (), {})) == [True] * B
    venv.close()


def test_new_returns_vecenv_when_num_envs_gt_1():
    from environment.env import SpatialKuramoto
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
    d = make_params("env0", 3, transient_state_len=118.0, num_envs=4)
    v = SpatialKuramoto(d)
    assert isinstance(v, BatchedKuramotoVecEnv) and v.num_envs == 4
    v.close()


@pytest.mark.parametrize("precision", ["f32"])
def test_full_size_batch_properties(precision):
    """B = 4096 (BASELINE config 2/3 size): size-independent properties.
    * environments with identical inputs give bit-identical outputs wherever they sit in the batch
    * every environment agrees with the same environment run alone (B = 1)
    * the observation window slides by exactly the step's sample count
    * counters: 5 accepted sub-steps and 32 RHS evaluations per env-step."""
    from dbsgym_b200.batched import BatchedKuramoto
    B = 4096
    g = load_golden("step_env0.npz")
    d = make_params("env0", 10)
    core = BatchedKuramoto([d] * 1, precision=precision)
    # build the 4096 batch directly on the engine (host reset of 4096 envs is not what is tested here)
    from dbsgym_b200.engine import KuramotoEngine
    from dbsgym_b200.geometry import coupling_table
    table = coupling_table(d["neur_coords"], d["neur_grid"], d["grid_size"], "cos")
    eng = KuramotoEngine(B, 512, [8, 8, 8], 2340, d["K"], precision=precision, coupling_table=table)
    eng.set_schedule(core.schedule)
    eng.set_reward("bbpow_action", 0.05)
    rng = np.random.default_rng(0)
    kinds = 8                                             # 8 distinct environments, tiled
    w0 = np.abs(rng.normal(1.0, 0.4, (kinds, 512))) + 0.05
    y0 = rng.normal(np.pi, 0.6, (kinds, 512)) + 40.0
    stim = np.tile(g["stim_cond"][0], (kinds, 1))
    win = rng.uniform(-0.3, 0.3, (kinds, 2340))
    acts = rng.uniform(-1, 1, kinds).astype(np.float32)
    rep = B // kinds
    tile = lambda a: np.tile(a, (rep,) + (1,) * (a.ndim - 1))     # noqa: E731
    eng.set_env_params(None, w0=tile(w0), stim=tile(stim), rec=tile(stim), y0=tile(y0))
    eng.set_window(tile(win))
    eng.set_episode(None, step_idx=0, episode_len=3)
    eng.counters(reset=True)
    obs, rew, done = eng.step_host(tile(acts))
    y = eng.state()
    for k in range(kinds):
        assert np.all(y[k::kinds] == y[k]) and np.all(obs[k::kinds] == obs[k]) and np.all(rew[k::kinds] == rew[k])
    c = eng.counters()
    assert (c["accepted"], c["rejected"], c["rhs_evals"], c["status"]) == (B * 5, 0, B * 32, 0)
    s = int(core.schedule.n_I[0] + core.schedule.n_II[0] - 1)
    assert np.all(eng.lfp()[2] == s)
    assert np.allclose(obs[:kinds, :2340 - s], win[:, s:].astype(np.float32), atol=1e-7)
    # same environments alone
    solo = KuramotoEngine(kinds, 512, [8, 8, 8], 2340, d["K"], precision=precision, coupling_table=table)
    solo.set_schedule(core.schedule); solo.set_reward("bbpow_action", 0.05)
    solo.set_env_params(None, w0=w0, stim=stim, rec=stim, y0=y0); solo.set_window(win)
    solo.set_episode(None, step_idx=0, episode_len=3)
    o1, r1, d1 = solo.step_host(acts)
    if precision == "f64":
        assert np.array_equal(solo.state(), y[:kinds]) and np.array_equal(o1, obs[:kinds]) and np.array_equal(r1, rew[:kinds])
    else:
        # fp32: the 4096-batch runs the multi-worker kernel, the 8-batch one CTA per environment; the lane order of
        # the warp reductions (LFP samples, error norm) differs between the two, everything else is the same sum
        assert eng.step_variant() == 4 and solo.step_variant() == 3
        np.testing.assert_allclose(solo.state(), y[:kinds], rtol=0, atol=2e-6)
        np.testing.assert_allclose(o1, obs[:kinds], rtol=0, atol=3e-7)
        np.testing.assert_allclose(r1, rew[:kinds], rtol=2e-6)
    # done after episode_len steps
    assert not done.any()
    eng.step_host(tile(acts)); _, _, done = eng.step_host(tile(acts))
    assert done.all() and np.all(eng.episode()[0] == 3)
    eng.close(); solo.close(); core.close()


def test_delta_transfer_equals_full_observation_readback():
    """The host-mirrored window (only the new samples cross PCIe) must give bit-identical observations,
    rewards and done flags to the full [B, W] read-back, across a mirror compaction (> 130 steps) and an
    episode boundary with auto-reset."""
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
    B, steps = 3, 150
    mk = lambda: [make_params("env1", 30 + e, transient_state_len=118.0, total_episode_len=126.0,   # noqa: E731
                              rand_seed=70 + e) for e in range(B)]
    outs = {}
    for mode in ("full", "delta"):
        venv = BatchedKuramotoVecEnv(mk(), transfer=mode)
        obs = venv.reset()
        rng = np.random.default_rng(9)
        rec = [obs.copy()]
        dones = 0
        for k in range(steps):
            o, r, d, infos = venv.step(rng.uniform(-1, 1, (B, 1)).astype(np.float32))
            rec.append(o.copy()); rec.append(r.copy()); rec.append(d.copy())
            if d.any():
                dones += 1
                rec.append(np.stack([infos[i]["terminal_observation"] for i in range(B)]))
        assert dones == 1 and venv.get_attr("total_episode_counts")[0] == 140
        outs[mode] = rec
        venv.close()
    assert len(outs["full"]) == len(outs["delta"])
    for a, b in zip(outs["full"], outs["delta"]):
        assert np.array_equal(a, b)


def test_step_tensor_matches_host_step():
    import torch
    from dbsgym_b200.batched import BatchedKuramoto
    d = [make_params("env0", 40 + e, transient_state_len=118.0) for e in range(2)]
    a = BatchedKuramoto(copy.deepcopy(d)); b = BatchedKuramoto(copy.deepcopy(d))
    acts = np.array([0.3, -0.9], dtype=np.float32)
    for _ in range(3):
        o1, r1, d1 = a.step(acts)
        o2, r2, d2 = b.step_tensor(torch.from_numpy(acts).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(o1, o2.cpu().numpy()) and np.array_equal(r1, r2.cpu().numpy())
        assert np.array_equal(d1, d2.cpu().numpy().astype(bool))
    a.close(); b.close()


def test_batched_evaluation_reproduces_paper_hf_dbs_row():
    """evaluate_batched + HFDBS on env0 eval configs: paper table HF-DBS beta power 2.34e-3 (sd 0.2e-3),
    energy 5555 (data/kur-table-metrics.xlsx row 5; the reference multiplies |action|=1 by ... 5555 = 1111 x 5)."""
    from dbsgym_b200.controllers import HFDBS
    from dbsgym_b200.evaluation import evaluate_batched
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
    dicts = [make_params("env0", 11 + 9 * e, total_episode_len=1000, rand_seed=11 + e) for e in range(4)]
    venv = BatchedKuramotoVecEnv(dicts)
    res = evaluate_batched(HFDBS(1.0), venv)
    assert res["bbpow"].shape == (4,) and len(res["true_lfp"][0]) > 1111 * 17
    assert abs(res["summary"]["bbpow_mean"] - 2.34e-3) < 3 * 0.2e-3
    assert np.allclose(res["energy"], 1111.0)            # sum |a| with a = 1; x5 after rescaling = the paper's 5555
    # the same episode with the trace recorded and the metric evaluated on the device (same seeds -> same episode)
    venv.close()
    venv = BatchedKuramotoVecEnv([make_params("env0", 11 + 9 * e, total_episode_len=1000, rand_seed=11 + e) for e in range(4)])
    dev = evaluate_batched(HFDBS(1.0), venv, on_device=True)
    np.testing.assert_allclose(dev["bbpow"], res["bbpow"], rtol=1e-7)
    assert np.array_equal(np.asarray(dev["true_lfp"]), np.asarray(res["true_lfp"]))
    venv.close()


def test_torch_policy_rollout_example_runs():
    import subprocess, sys, os
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "rollout_torch_policy.py"), "--envs", "64",
                          "--steps", "4", "--cfg", "env0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "env-steps/s" in out.stdout


def test_env2_drift_resets_match_reference_golden():
    """env2 (temporal drift: electrode movement, encapsulation, plasticity-walk regeneration, spatial re-draw) through
    the GPU-backed SpatialKuramoto: 14 resets against the fixture the REFERENCE's own env.py produced
    (tests/golden/make_golden.py:record_env2_events; env.py:483-557 + the transient of env.py:605-612)."""
    from environment.env import SpatialKuramoto
    g = load_golden("env2_events.npz")
    d = make_params("env2", 21, plasticity_drift_freq=10 ** 6, transient_state_len=117.5,
                    total_episode_len=9., spatial_var_freq=4, precision="f64")
    assert np.array_equal(d["w0"], g["in_w0"])
    env = SpatialKuramoto(d)
    for r in range(len(g["elec"])):
        if r > 0:
            obs, info = env.reset()
            assert obs.shape == (1, 2340) and info == {}
        assert np.array_equal(np.array(env.elec_coords)[0], g["elec"][r]), r
        assert np.array_equal(np.array(env.rec_coords)[0], g["rec"][r]), r
        assert env.encapsulation_coeff == g["encaps"][r]
        assert np.array_equal(env.kuramoto.w0, g["w0"][r])
        assert np.array_equal(env.init_state, g["init_state"][r])
        np.testing.assert_allclose(env.theta_state[0][:8], g["window_head"][r], rtol=0, atol=1e-9)
        assert env.elec_drift_episode == g["elec_drift_episode"][r]
        assert env.elec_encaps_episode == g["encaps_episode"][r]
    env.close()


def _short_dicts(B, seed=30, episode_len=126.0):
    return [make_params("env1", seed + e, transient_state_len=118.0, total_episode_len=episode_len,
                        rand_seed=70 + seed + e) for e in range(B)]


def test_step_wait_observation_stays_intact_across_following_steps_and_resets():
    """ADVICE r1 (high) / VERDICT weak #2: the array step_wait() returns must not be modified by later steps.  The
    default (zero-copy) VecEnv hands out a view of the pinned host log; it is held here WITHOUT copying across the
    next 13 steps -- among them the auto-reset of every environment -- and compared with a copy taken at once."""
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
    B = 4
    venv = BatchedKuramotoVecEnv(_short_dicts(B, episode_len=18.0))          # 20-step episodes
    held = [(venv.reset(),)]
    held[0] = (held[0][0], held[0][0].copy(), 0)
    rng = np.random.default_rng(4)
    n_done = 0
    for k in range(1, 60):
        obs, rew, done, infos = venv.step(rng.uniform(-1, 1, (B, 1)).astype(np.float32))
        n_done += int(done.all())
        held.append((obs, obs.copy(), k))
        for view, snapshot, born in held:
            assert np.array_equal(view, snapshot), f"observation of step {born} changed during step {k}"
        held = held[-13:]                                   # guaranteed lifetime: 13 further steps
    assert n_done == 2
    venv.close()


class _RolloutBufferStandIn:
    """The part of stable_baselines3.common.buffers.RolloutBuffer.add that matters here: the observation is copied into
    the buffer WHEN add() IS CALLED -- and OnPolicyAlgorithm.collect_rollouts calls add(self._last_obs, ...) only after
    env.step() of the following step has returned (same for _store_transition of the off-policy algorithms)."""

    def __init__(self, n_steps, B, W):
        self.observations = np.zeros((n_steps, B, 1, W), dtype=np.float32)
        self.rewards = np.zeros((n_steps, B), dtype=np.float32)
        self.episode_starts = np.zeros((n_steps, B), dtype=np.float32)
        self.pos = 0

    def add(self, obs, reward, episode_start):
        self.observations[self.pos] = np.array(obs)
        self.rewards[self.pos] = np.array(reward)
        self.episode_starts[self.pos] = np.array(episode_start)
        self.pos += 1


def _collect_rollouts(venv, n_steps, seed):
    """stable_baselines3 OnPolicyAlgorithm.collect_rollouts, call for call, with a random policy."""
    B = venv.num_envs
    buf = _RolloutBufferStandIn(n_steps, B, venv.observation_space.shape[1])
    rng = np.random.default_rng(seed)
    last_obs = venv.reset()
    last_episode_starts = np.ones(B, dtype=bool)
    terminal = []
    for _ in range(n_steps):
        actions = rng.uniform(-1, 1, (B, 1)).astype(np.float32)       # policy(obs_as_tensor(self._last_obs))
        new_obs, rewards, dones, infos = venv.step(actions)
        for idx, done in enumerate(dones):
            if done and infos[idx].get("terminal_observation") is not None:
                terminal.append(np.array(infos[idx]["terminal_observation"]))
        buf.add(last_obs, rewards, last_episode_starts)               # <- after the step, exactly like SB3
        last_obs = new_obs
        last_episode_starts = dones
    return buf, terminal


def test_sb3_collect_rollouts_call_order_gives_the_same_buffer_as_owned_copies():
    """The zero-copy VecEnv driven in SB3's call order (rollout_buffer.add(self._last_obs) AFTER the next env.step) must
    fill the rollout buffer with exactly what a VecEnv that returns owned copies (DummyVecEnv semantics) gives."""
    from dbsgym_b200.vec_env import BatchedKuramotoVecEnv
    B, n_steps = 3, 45
    bufs = {}
    for mode, kw in (("views", {}), ("copies", {"copy_obs": True}), ("full", {"transfer": "full", "copy_obs": True})):
        venv = BatchedKuramotoVecEnv(_short_dicts(B, seed=80, episode_len=18.0), **kw)
        bufs[mode] = _collect_rollouts(venv, n_steps, seed=2)
        venv.close()
    ref, ref_term = bufs["full"]
    assert ref.episode_starts.sum() == B * 3 and len(ref_term) == 2 * B      # two episode ends inside the rollout
    for mode in ("views", "copies"):
        buf, term = bufs[mode]
        assert np.array_equal(buf.observations, ref.observations), mode
        assert np.array_equal(buf.rewards, ref.rewards) and np.array_equal(buf.episode_starts, ref.episode_starts)
        assert all(np.array_equal(a, b) for a, b in zip(term, ref_term))
    # consecutive stored observations are genuinely different windows (the check above is not vacuous)
    assert not np.array_equal(ref.observations[3], ref.observations[4])


def test_out_of_lockstep_batch_falls_back_to_owned_readback_and_recovers():
    """After a reset of a SUBSET the environments sit at different step indices (their steps append 17 / 18 / 19 samples
    at different times), so the windows are no longer one common slice of the host log: step() must notice, return a
    read-back (two alternating buffers: intact for one further step) that equals the device windows, and go back to the
    zero-copy path after the next reset of all environments."""
    from dbsgym_b200.batched import BatchedKuramoto
    B = 3
    core = BatchedKuramoto(_short_dicts(B, seed=60, episode_len=5400.0))
    rng = np.random.default_rng(1)
    for _ in range(70):                                   # past step 65, where the schedule changes from 18 to 19 / 17 samples
        core.step(rng.uniform(-1, 1, B).astype(np.float32))
    core.reset_envs([1])
    prev = None
    fell_back = False
    for k in range(70):
        obs, rew, done = core.step(rng.uniform(-1, 1, B).astype(np.float32))
        np.testing.assert_array_equal(obs, core.engine.window_values().astype(np.float32))
        fell_back = fell_back or not np.shares_memory(obs, core._mirror)
        if prev is not None:
            assert np.array_equal(prev[0], prev[1])       # last step's array survived this step
        prev = (obs, obs.copy())
    assert fell_back
    core.reset_envs(range(B))
    obs, rew, done = core.step(rng.uniform(-1, 1, B).astype(np.float32))
    assert np.shares_memory(obs, core._mirror)
    np.testing.assert_array_equal(obs, core.engine.window_values().astype(np.float32))
    core.close()


def test_snapshot_restore_round_trip_is_bit_exact():
    """dbsgym_get_state / dbsgym_set_state: run 5 steps, snapshot, run 6 more; restore into the same handle AND into a
    freshly built one, repeat the 6 steps: identical observations, rewards, phases, counters."""
    from dbsgym_b200.batched import BatchedKuramoto
    B = 3
    dicts = _short_dicts(B, seed=90, episode_len=5400.0)
    a = BatchedKuramoto(copy.deepcopy(dicts), transfer="full")
    acts = np.random.default_rng(3).uniform(-1, 1, (11, B)).astype(np.float32)
    for k in range(5):
        a.step(acts[k])
    blob = a.engine.get_state()

    def tail(core):
        out = []
        for k in range(5, 11):
            o, r, d = core.step(acts[k])
            out.append((o.copy(), r.copy(), d.copy(), core.engine.state().copy(), core.engine.rewards()[0].copy()))
        return out, core.engine.counters()
    ref, c_ref = tail(a)
    a.engine.set_state(blob)
    again, c_again = tail(a)
    b = BatchedKuramoto(copy.deepcopy(dicts), transfer="full")         # another handle of the same shape
    b.engine.set_state(blob)
    other, c_other = tail(b)
    assert c_ref == c_again == c_other
    for x, y, z in zip(ref, again, other):
        for u, v, w in zip(x, y, z):
            assert np.array_equal(u, v) and np.array_equal(u, w)
    with pytest.raises(Exception, match="shape|bytes"):
        c = BatchedKuramoto(copy.deepcopy(dicts[:2]), transfer="full")
        try:
            c.engine.set_state(blob)
        finally:
            c.close()
    a.close(); b.close()


def test_trace_capacity_can_shrink_between_evaluations():
    """a second, SHORTER trace on the same handle must use its own capacity as the row stride."""
    from dbsgym_b200.batched import BatchedKuramoto
    B = 2
    core = BatchedKuramoto(_short_dicts(B, seed=95, episode_len=5400.0))
    eng = core.engine
    acts = np.full(B, 0.3, dtype=np.float32)
    eng.trace_begin(19 * 6)
    for _ in range(6):
        core.step(acts)
    t1, n1 = eng.trace()
    assert t1.shape == (B, 19 * 6) and np.all(n1 > 17 * 5)
    eng.trace_begin(19 * 2)
    lf = []
    for _ in range(2):
        core.step(acts)
        t, _, n = eng.lfp()
        lf.append([t[i, :n[i]].copy() for i in range(B)])
    t2, n2 = eng.trace()
    assert t2.shape == (B, 19 * 2)
    for i in range(B):
        want = np.concatenate([lf[0][i], lf[1][i]])
        assert n2[i] == len(want) and np.array_equal(t2[i, :n2[i]], want)
    core.close()


def test_step_tensor_is_ordered_against_own_stream_work_without_explicit_sync():
    """step_tensor launches on torch's current stream, reset_envs / observations / step on the
    handle's private stream.  No torch.cuda.synchronize() between them here: the library orders the two streams itself."""
    import torch
    from dbsgym_b200.batched import BatchedKuramoto
    B = 64
    dicts = _short_dicts(2, seed=97, episode_len=5400.0)
    dicts = [copy.deepcopy(dicts[e % 2]) for e in range(B)]
    res = []
    for sync in (True, False):
        core = BatchedKuramoto(copy.deepcopy(dicts))
        acts = torch.full((B,), 0.4, device="cuda")
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(4):
                o, r, d = core.step_tensor(acts)
                if sync:
                    torch.cuda.synchronize()
            y_mid = core.engine.state()                  # (device-synchronising getter)
            o, r, d = core.step_tensor(acts)
            if sync:
                torch.cuda.synchronize()
            w = core.observations().copy()               # own stream right behind a user-stream step
            o2, r2, d2 = core.step(np.full(B, -0.2, dtype=np.float32))
            o3, r3, d3 = core.step_tensor(acts)          # user stream right behind an own-stream step
        torch.cuda.synchronize()
        res.append((y_mid, w, np.array(o2), r2.copy(), o3.cpu().numpy(), r3.cpu().numpy(), core.engine.state()))
        core.close()
    for u, v in zip(*res):
        assert np.array_equal(u, v)
