"""Product HOST code (dbsgym_b200/*.py) against the reference-generated fixtures.  CPU only.
Integer / index results must be bit-exact; float vectors agree to rounding."""
import re

import numpy as np
import pytest

from conftest import load_golden, make_params
from dbsgym_b200 import geometry, utils
from dbsgym_b200.host_env import HostEnvState, generate_perturbations
from dbsgym_b200.schedule import StepSchedule, transient_grid


def test_grid_order_and_contact_indices_bit_exact():
    g = load_golden("geometry.npz")
    coords, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    assert np.array_equal(grid, g["neur_grid"])
    assert geometry.contact_index([4, 3, 4], [8, 8, 8]) == 284 and grid[284].tolist() == [3, 4, 4]
    for n, gs in ((256, (8, 8, 8)), (100, (5, 5, 5))):
        assert np.array_equal(utils.generate_neuron_grid_3D(*gs, n, 0.1)[1], g[f"grid_{n}_{gs[0]}"])
    with pytest.raises(ValueError):
        utils.generate_neuron_grid_3D(2, 2, 2, 9)


def test_electrode_model_matches_reference():
    g = load_golden("geometry.npz")
    _, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    for i in range(5):
        el = geometry.ElectrodeModel([8, 8, 8], grid, float(g[f"case{i}_cm"]), g[f"case{i}_elec"].tolist(),
                                     g[f"case{i}_rec"].tolist(), [0.], bool(g[f"case{i}_directed"]), "dummy")
        assert el.elec_idxs == g[f"case{i}_elec_idx"].tolist()
        assert el.rec_idxs == g[f"case{i}_rec_idx"].tolist()
        # which neurons are stimulated / recorded is an index set: bit-exact
        assert np.array_equal(el.conductances[0] > 0, g[f"case{i}_cond"] > 0)
        assert np.array_equal(el.rec_conductances[0] > 0, g[f"case{i}_rec_cond"] > 0)
        np.testing.assert_allclose(el.conductances[0], g[f"case{i}_cond"], rtol=0, atol=5e-16)
        np.testing.assert_allclose(el.rec_conductances[0], g[f"case{i}_rec_cond"], rtol=0, atol=5e-16)
        if g[f"case{i}_directed"]:
            assert np.array_equal(np.array(el.directional_masks_list[0]), g[f"case{i}_masks"])
    # notebook known answer (explore_kuramoto_dynamics.ipynb cell 3): 512 neurons, min 0.307, max 1.0
    el = geometry.ElectrodeModel([8, 8, 8], grid, 0.1, [[4, 3, 4]], [[1, 1, 1]], [0.])
    c = el.conductances[0]
    assert (np.count_nonzero(c > 0), round(c.min(), 3), round(c.max(), 3)) == (512, 0.307, 1.0)
    with pytest.raises(AssertionError):
        geometry.ElectrodeModel([8, 8, 8], grid, 0.1, [[4, 3, 4]], [[1, 1, 1]], [0., 1.])
    with pytest.raises(ValueError):
        geometry.ElectrodeModel([8, 8, 8], grid, 0.1, [[4, 3, 4]], [[1, 1, 1]], [0.], prc_type="bogus")


def test_locus_mask_and_w0_generation_bit_exact():
    g = load_golden("geometry.npz")
    _, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    assert np.array_equal(geometry.locus_mask(grid, [8, 8, 8], [4, 4, 4], 0.55), g["locus_mask_444_055"])
    s = load_golden("step_env0.npz")
    d = make_params("env0", 10)
    for k in ("w0", "w0_without_locus", "locus_without_w0", "locus_mask"):
        assert np.array_equal(d[k], s[k]), k


def test_coupling_table_expands_to_dense_alpha():
    g = load_golden("geometry.npz")
    coords, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    for kern, amp, st in (("cos", 1.0, 0.6), ("wavelet", 1.3, 0.6)):
        t = geometry.coupling_table(coords, grid, [8, 8, 8], kern, amp, st)
        assert t is not None and t.shape == (512,)
        alpha = geometry.coupling_rows(coords, np.arange(512), kern, amp, st)
        d = np.abs(grid[:, None, :] - grid[None, :, :])
        expanded = t.reshape(8, 8, 8)[d[..., 2], d[..., 0], d[..., 1]]
        assert np.max(np.abs(expanded - alpha)) < 1e-13
    np.testing.assert_allclose(geometry.coupling_rows(coords, [0, 284], "cos"),
                               np.stack([g["alpha_row0"], g["alpha_row284"]]), rtol=0, atol=3e-16)
    # half grid (first 256 rows) is still whole z-planes; shuffled or non-8 lines are not
    c2, g2 = geometry.neuron_grid(8, 8, 8, 256, 0.1)
    assert geometry.coupling_table(c2, g2, [8, 8, 8], "cos") is not None
    perm = np.random.default_rng(0).permutation(512)
    assert geometry.coupling_table(coords[perm], grid[perm], [8, 8, 8], "cos") is None
    c3, g3 = geometry.neuron_grid(5, 5, 5, 100, 0.1)
    assert geometry.coupling_table(c3, g3, [5, 5, 5], "cos") is None
    with pytest.raises(ValueError):
        geometry.kernel_values(np.zeros(3), "gauss")


def test_schedule_bit_exact():
    g = load_golden("schedule.npz")
    tt = transient_grid(200., 0.05)
    assert len(tt) == 4000 and tt[-1] == 199.95000000000002
    for tag, n in (("train", 5555), ("eval", 1111)):
        s = StepSchedule(n, tt[-1], 0.15, 0.75, 0.05)
        assert np.array_equal(s.n_I, g[f"nI_{tag}"]) and np.array_equal(s.n_II, g[f"nII_{tag}"])
        assert s.t_after[-1] == g[f"t_final_{tag}"]
        assert np.array_equal(s.t_after[::100], g[f"t_cur_every100_{tag}"])
        assert s.max_samples == 19 and (s.max_I, s.max_II) == (4, 16)
    st = load_golden("step_env0.npz")
    s = StepSchedule(70, tt[-1], 0.15, 0.75, 0.05)
    assert np.array_equal(s.n_I, st["nI"]) and np.array_equal(s.n_II, st["nII"])
    for k in range(70):
        assert np.array_equal(s.offs_I[k, :s.n_I[k]], st["offs_I"][k, :s.n_I[k]])
        assert np.array_equal(s.offs_II[k, :s.n_II[k]], st["offs_II"][k, :s.n_II[k]])
    assert np.array_equal(s.t_after, st["t_cur"])


def test_host_rewards_and_linear_functional():
    g = load_golden("rewards.npz")
    lo, hi = utils.beta_bins(2340, 0.0005, 12.5, 21)
    assert (lo, hi) == (15, 24)
    gvec = utils.temp_const_functional(2340, 2000.0, order=2)
    for w, u, r, bb in zip(g["windows"], g["u"], g["rewards"], g["bbpow"]):
        assert utils.calc_beta_band_power(w, 0.0005, 12.5, 21) == pytest.approx(bb, rel=1e-12)
        # R2 as a dot product with the precomputed functional == filtfilt pipeline of the reference
        r2 = -1e3 * float(gvec @ w) ** 2 - 1e-2 * abs(u)
        assert r2 == pytest.approx(r[1], rel=1e-7, abs=1e-9)
        # circular-shift invariance used by the observation kernel
        assert utils.calc_beta_band_power(np.roll(w, 777), 0.0005, 12.5, 21) == pytest.approx(bb, rel=1e-11)


def test_host_env_state_replays_env2_events_and_rng_order():
    g = load_golden("env2_events.npz")
    d = make_params("env2", 21, plasticity_drift_freq=10 ** 6, transient_state_len=117.5,
                    total_episode_len=9., spatial_var_freq=4)
    h = HostEnvState(d)
    assert h.total_episode_counts == 10 and h.observe_wind_idxs == 2340
    for r in range(len(g["elec"])):
        s = h.begin_episode()
        assert h.reset_count == r
        assert np.array_equal(np.array(h.elec_coords)[0], g["elec"][r]), r
        assert np.array_equal(np.array(h.rec_coords)[0], g["rec"][r]), r
        assert h.encapsulation_coeff == g["encaps"][r]
        assert np.array_equal(s.w0, g["w0"][r])
        assert np.array_equal(s.y0, g["init_state"][r])
        assert h.elec_drift_episode == g["elec_drift_episode"][r]
        assert h.elec_encaps_episode == g["encaps_episode"][r]


def test_host_env_state_errors_like_reference():
    with pytest.raises(AssertionError):
        HostEnvState(make_params("env2", 3))                        # env.py:368
    h = HostEnvState(make_params("env2", 3), compat_env2=True)      # explicit compatibility switch
    h.begin_episode()
    for bad in (dict(reward_func="nope"), dict(recording_kernel="nope"), dict(transient_state_len=100.)):
        with pytest.raises(ValueError):
            HostEnvState(make_params("env0", 3, **bad))


def test_generate_perturbations_matches_definition():
    v = np.random.default_rng(1).uniform(0.1, 2, 64)
    np.random.seed(4)
    w = generate_perturbations(v, M=5, step_scale=0.02)
    np.random.seed(4)
    exp = [v.copy()]
    for _ in range(5):
        exp.append(exp[-1] + 0.02 * np.std(v, ddof=1) * np.random.randn(64))
    assert np.array_equal(w, np.array(exp)) and w.shape == (6, 64)


def test_configs_expose_reference_names():
    import environment.env_configs.env0 as e0
    import environment.env_configs.env1 as e1
    import environment.env_configs.env2 as e2
    import data.configs.env2 as d2
    for m in (e0, e1, e2, d2):
        assert len(m.eval_envs_list) == 5 and m.n_neurons == 512 and m.grid_size == [8, 8, 8]
        assert m.params_dict_train["observe_wind_counts"] == 130
    assert len(e1.stim_rec_locus_coordinates) == 15 and len(e2.stim_rec_locus_coordinates) == 40
    assert e0.params_dict_train["recording_kernel"] == "naive" and e1.params_dict_train["recording_kernel"] == "gaussian"
    assert e2.params_dict_train["temporal_drift"] and e2.eval0["electrode_drift_freq"] == 2
    assert e0.eval0["total_episode_len"] == 1000 and e0.eval2["rand_seed"] == 20


def test_batched_fast_reset_draws_like_the_sequential_loop():
    """BatchedKuramoto's one-call initial-phase draw must leave every host (and the global RNG) exactly
    where the per-environment begin_episode() loop would."""
    import copy
    from dbsgym_b200.batched import BatchedKuramoto

    class _Stub(BatchedKuramoto):                 # host logic only: no engine, no GPU
        def __init__(self, hosts, n):
            self.hosts, self.n_osc = hosts, n

    dicts = [make_params("env1", 3 + e, rand_seed=90 + e) for e in range(5)]
    assert any(np.any(utils.apply_locus_mask(d["w0_without_locus"], d["locus_without_w0"], d["locus_mask"]) <= 0)
               for d in dicts)                     # exercise the remove_negative_w0 draws too
    slow = [HostEnvState(copy.deepcopy(d)) for d in dicts]
    fast = [HostEnvState(copy.deepcopy(d)) for d in dicts]
    np.random.seed(123)
    for h in slow:
        h.begin_episode()
    np.random.seed(123)
    for h in fast:
        h.begin_episode()                          # first reset is always the ordinary path (fills the caches)
    for rnd in range(12):                          # env1: spatial re-draw fires at reset_count 10 -> slow path there
        np.random.seed(1000 + rnd)
        ref = [h.begin_episode() for h in slow]
        state_ref = np.random.get_state()[1].copy()
        np.random.seed(1000 + rnd)
        stub = _Stub(fast, 512)
        got = stub._begin_episodes_fast(range(5))
        used_fast = got is not None
        if got is None:
            got = [h.begin_episode() for h in fast]
        assert used_fast == (fast[0].reset_count != 10)        # whether or not some w0 has non-positive entries
        assert np.array_equal(np.random.get_state()[1], state_ref)
        for a, b, hs, hf in zip(ref, got, slow, fast):
            assert np.array_equal(a.y0, b.y0) and np.array_equal(a.w0, b.w0)
            assert np.array_equal(a.stim, b.stim) and np.array_equal(a.rec, b.rec)
            assert hs.reset_count == hf.reset_count and hs.elec_coords == hf.elec_coords
            assert hs.spatial_var_episode == hf.spatial_var_episode


def test_batched_eval_metric_equals_row_by_row_reference_formula():
    """calc_psd_for_simple_eval (aDBS_RL/evaluate_HF_DBS.py:122-135): the batched 2-D path (one filtfilt / rfft /
    filtfilt along the last axis) gives the same beta-band power as the reference's per-signal loop."""
    from scipy.signal import butter, filtfilt
    from dbsgym_b200.evaluation import calc_psd_for_simple_eval
    rng = np.random.default_rng(5)
    t = np.arange(19998) * 0.0005
    x = 0.05 * rng.normal(size=(6, t.size)).cumsum(axis=1) / 30 + 0.1 * np.sin(2 * np.pi * 17.0 * t)[None, :] * rng.uniform(0.5, 1.5, (6, 1))
    ref = []
    for sig in x:                                    # the reference's loop, written out
        b, a = butter(2, [12 / 1000.0, 30 / 1000.0], btype="band")
        f = filtfilt(b, a, sig)
        ft = np.abs(np.fft.rfft(f) / f.shape[0]) ** 2 * 2
        freq = np.fft.rfftfreq(f.shape[0], 0.0005)
        ft = filtfilt([1] * 12, 5, ft)
        ref.append(np.sum(ft[np.where((freq > 12.5) & (freq < 21))]))
    np.testing.assert_allclose(calc_psd_for_simple_eval(x, 0.0005), ref, rtol=1e-12)
    np.testing.assert_allclose(calc_psd_for_simple_eval([r for r in x], 0.0005), ref, rtol=1e-12)


def test_coupling_table_for_cubic_grids_with_lines_of_16_and_32():
    """The structured kernels also take lines of 16 / 32 oscillators (cubic 16^3 / 32^3 grids):
    the table still expands to the dense alpha of utils.py:457-466 + env.py:219-223; shapes the kernels cannot serve
    (odd extents, too few fundamental lines for warp-uniform chunks, lines of 24) fall back to DENSE (None)."""
    for G, gz in ((16, 4), (32, 4)):
        n = G * G * gz
        coords, grid = geometry.neuron_grid(G, G, G, n, 0.1)
        t = geometry.coupling_table(coords, grid, [G, G, G], "cos")
        assert t is not None and t.shape == (n,)
        rows = np.array([0, 17, n // 2 + 5, n - 1])
        alpha = geometry.coupling_rows(coords, rows, "cos")
        d = np.abs(grid[rows][:, None, :] - grid[None, :, :])
        expanded = t.reshape(gz, G, G)[d[..., 2], d[..., 0], d[..., 1]]
        assert np.max(np.abs(expanded - alpha)) < 1e-13
    c, g = geometry.neuron_grid(16, 16, 16, 16 * 16 * 3, 0.1)          # odd number of populated z-planes
    assert geometry.coupling_table(c, g, [16, 16, 16], "cos") is None
    c, g = geometry.neuron_grid(4, 16, 16, 4 * 16 * 4, 0.1)            # 2 x 2 fundamental lines: chunks not warp-uniform
    assert geometry.coupling_table(c, g, [4, 16, 16], "cos") is None
    c, g = geometry.neuron_grid(8, 24, 8, 8 * 24 * 8, 0.1)             # lines of 24
    assert geometry.coupling_table(c, g, [8, 24, 8], "cos") is None


def test_device_metric_weights_fold_smoothing_and_band_sum():
    """evaluation.bbpow_spec: sum(filtfilt([1]*12, 5, ft)[band]) == w @ ft[k_lo : k_lo + len(w)] for arbitrary spectra --
    the identity the device evaluation kernel relies on (evaluate_HF_DBS.py:130-134)."""
    from scipy.signal import filtfilt
    from dbsgym_b200.evaluation import bbpow_spec
    for n in (19039, 19998, 4001):
        sp = bbpow_spec(n, 0.0005)
        freq = np.fft.rfftfreq(n, 0.0005)
        band = (freq > 12.5) & (freq < 21)
        ft = np.random.default_rng(n).uniform(0, 1, freq.size) ** 3
        ref = filtfilt([1] * 12, 5, ft)[band].sum()
        w = sp["weights"]
        assert abs(w @ ft[sp["k_lo"]:sp["k_lo"] + w.size] - ref) < 1e-12 * abs(ref)
        assert sp["padlen"] == 15 and len(sp["b"]) == 5 and sp["a"][0] == 1.0


def test_vecenv_infos_are_per_environment_dicts():
    """infos[i] must be a dict of its own (wrappers mutate them), created lazily."""
    from dbsgym_b200.vec_env import _InfoList
    infos = _InfoList(5)
    infos[1]["x"] = 1
    assert "x" not in infos[0] and infos[1] == {"TimeLimit.truncated": False, "x": 1}
    assert len(infos) == 5 and all(isinstance(d, dict) for d in infos) and infos[-1] is infos[4]
    assert len({id(d) for d in infos}) == 5
    infos[2] = {"terminal_observation": 3}
    assert infos[2] == {"terminal_observation": 3} and [d for d in infos[1:3]] == [infos[1], infos[2]]


def test_batch_rejects_environments_with_different_geometry():
    """the coupling operator is built from the first dict only, so neur_coords / neur_grid are shared keys."""
    import copy
    from dbsgym_b200.batched import _SHARED_KEYS, _same
    assert "neur_coords" in _SHARED_KEYS and "neur_grid" in _SHARED_KEYS
    d = make_params("env0", 3)
    e = copy.deepcopy(d)
    e["neur_coords"] = e["neur_coords"] * 1.5
    assert _same(d["neur_grid"], e["neur_grid"]) and not _same(d["neur_coords"], e["neur_coords"])
    from dbsgym_b200.batched import BatchedKuramoto
    with pytest.raises(ValueError, match="neur_coords"):
        BatchedKuramoto([d, e])                        # refused before any device is touched


def test_spectral_factors_reproduce_the_coupling_operator():
    """geometry.spectral_factors (the generalised mean-field form the float32 kernel evaluates): the truncated sector
    eigen-decomposition must reproduce alpha = cos(distance) (env.py:219-223) to the reported residual, the leading mode
    must be the classical mean field, and the ranks must fit the compiled kernel (9 even / 4 odd at 1e-10)."""
    coords, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    table = geometry.coupling_table(coords, grid, [8, 8, 8], "cos")
    alpha = geometry.coupling_rows(coords, np.arange(512), "cos")
    vecs, vals, ranks, residual = geometry.spectral_factors(table, 8, 8, 8, tol=1e-10)
    assert ranks == [9, 4, 4, 4, 4, 4, 4, 1] and residual < 3e-8
    lam = np.linalg.eigvalsh(alpha)
    assert abs(vals[0, 0] - lam[-1]) < 1e-9 * lam[-1] and np.all(vecs[0, :, 0] * np.sign(vecs[0, 0, 0]) > 0)
    rng = np.random.default_rng(0)
    th = rng.uniform(0, 2 * np.pi, 512)
    for x in (np.sin(th), np.cos(th), np.eye(512)[137]):
        err = np.max(np.abs(geometry.spectral_apply(vecs, vals, x, 8, 8, 8) - alpha @ x))
        assert err <= residual * np.linalg.norm(x) * 1.0001 + 1e-13
    # RHS error of the identity, in rad per time unit, at the shipped K / N: far below float32 rounding (~1e-7)
    s, c = np.sin(th), np.cos(th)
    exact = 0.52 / 512 * (c * (alpha @ s) - s * (alpha @ c))
    approx = 0.52 / 512 * (c * geometry.spectral_apply(vecs, vals, s, 8, 8, 8) - s * geometry.spectral_apply(vecs, vals, c, 8, 8, 8))
    assert np.max(np.abs(exact - approx)) < 1e-10
    # all modes kept: exact to rounding
    v2, w2, r2, res2 = geometry.spectral_factors(table, 8, 8, 8, tol=0.0)
    assert sum(r2) == 512 and res2 == 0.0
    assert np.max(np.abs(geometry.spectral_apply(v2, w2, s, 8, 8, 8) - alpha @ s)) < 1e-11


def test_lowrank_factors_reproduce_a_shuffled_operator():
    """geometry.lowrank_factors (the form the float32 DENSE kernel evaluates, step-kernel variant 11): a shuffled 8 x 8 x 8
    grid has the 32 modes of the regular one; the measured residual is the true spectral norm of what was dropped; a
    rough kernel is reported as incompressible; the randomised path (N > 1024) agrees with the full diagonalisation."""
    from dbsgym_b200 import geometry
    coords, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    perm = np.random.default_rng(3).permutation(512)
    alpha = geometry.coupling_rows(coords[perm], np.arange(512), "cos")
    vecs, vals, residual = geometry.lowrank_factors(alpha, tol=1e-9)
    assert vecs.shape == (32, 512) and np.all(np.abs(vals[:-1]) >= np.abs(vals[1:]))
    rec = vecs.T @ (vals[:, None] * vecs)
    assert abs(np.linalg.norm(alpha - rec, 2) - residual) < 1e-3 * residual and residual < 1e-9 * np.abs(vals[0])
    rng = np.random.default_rng(0)
    th = rng.uniform(0, 2 * np.pi, 512)
    s, c = np.sin(th), np.cos(th)
    exact = 0.52 / 512 * (c * (alpha @ s) - s * (alpha @ c))
    approx = 0.52 / 512 * (c * (rec @ s) - s * (rec @ c))
    assert np.max(np.abs(exact - approx)) < 1e-10                       # rad per time unit
    assert geometry.lowrank_factors(np.exp(-40.0 * geometry.distances_from(coords, np.arange(512)) ** 2), tol=1e-9, max_rank=64) is None
    coords2, _ = geometry.neuron_grid(16, 16, 16, 2048, 0.1)
    alpha2 = geometry.coupling_rows(coords2, np.arange(2048), "cos")
    v2, w2, res2 = geometry.lowrank_factors(alpha2, tol=1e-9)
    wf = np.linalg.eigvalsh(alpha2)
    wf = wf[np.argsort(-np.abs(wf))]
    assert len(w2) == np.count_nonzero(np.abs(wf) > 1e-9 * np.abs(wf[0])) and np.allclose(w2, wf[:len(w2)], rtol=1e-6, atol=1e-9 * abs(wf[0]))
    assert abs(res2 - np.abs(wf[len(w2)])) < 1e-2 * res2


def test_grid_sector_factors_and_grid_lowrank_factors_agree_with_the_full_matrix():
    """The two factorisations of a regular grid's operator that never form the N x N matrix (sector-wise eigenpairs over the
    fundamental octant for dbsgym_set_coupling_lowrank_sectors; the same carried back to the grid for
    dbsgym_set_coupling_lowrank) against the dense alpha of a 16 x 16 x 8 grid: eigenvalues, reconstruction, padding."""
    from dbsgym_b200 import geometry
    gx, gy, gz, N = 16, 16, 8, 2048
    coords, grid = geometry.neuron_grid(gx, gy, gz, N, 0.1)
    table = geometry.coupling_table(coords, grid, [gx, gy, gz], "cos")
    alpha = geometry.coupling_rows(coords, np.arange(N), "cos")
    soff, z, w, res = geometry.grid_sector_factors(table, gx, gy, gz, tol=1e-9)
    assert soff.shape == (9,) and soff[0] == 0 and np.all(np.diff(soff) % 4 == 0) and z.shape == (soff[8], N // 8)
    v, lam, res2 = geometry.grid_lowrank_factors(table, gx, gy, gz, tol=1e-9)
    kept = np.sort(np.abs(w[w != 0]))[::-1]
    assert len(kept) == len(lam) and np.allclose(kept, np.abs(lam), rtol=1e-9)
    rec = v.T @ (lam[:, None] * v)
    assert abs(np.linalg.norm(alpha - rec, 2) - res2) < 2e-2 * res2 and res2 < 1e-9 * np.abs(lam[0]) and abs(res - res2) < 1e-2 * res2
    # sector form applied by hand: x -> sector coordinates over the octant -> modes -> back, against alpha @ x
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N).reshape(gz, gx, gy)
    hz, hx, hy = gz // 2, gx // 2, gy // 2
    out = np.zeros_like(x)
    for s in range(8):
        py, pz, px = (-1.0 if s & 4 else 1.0), (-1.0 if s & 2 else 1.0), (-1.0 if s & 1 else 1.0)
        X = np.zeros((hz, hx, hy))
        imgs = []
        for mz, sz in ((0, 1.0), (1, pz)):
            for mx, sx in ((0, 1.0), (1, px)):
                for my, sy in ((0, 1.0), (1, py)):
                    sl = (slice(None, None, -1) if mz else slice(None), slice(None, None, -1) if mx else slice(None),
                          slice(None, None, -1) if my else slice(None))
                    X += sz * sx * sy * x[sl][:hz, :hx, :hy]
                    imgs.append((sl, sz * sx * sy))
        zs, ws = z[soff[s]:soff[s + 1]], w[soff[s]:soff[s + 1]]
        y = (zs.T @ (ws * (zs @ X.ravel()))).reshape(hz, hx, hy) / 8.0
        for sl, sign in imgs:
            full = np.zeros_like(x)
            full[:hz, :hx, :hy] = sign * y
            out += full[sl]
    assert np.max(np.abs(out.ravel() - alpha @ x.ravel())) < 1e-8 * np.abs(alpha @ x.ravel()).max()


def test_grid_permutation_recognises_a_shuffled_regular_grid():
    """geometry.grid_permutation: the order map a shuffled grid is handed to the library with (dbsgym_set_oscillator_order)."""
    from dbsgym_b200 import geometry
    coords, grid = geometry.neuron_grid(8, 8, 8, 512, 0.1)
    assert geometry.grid_permutation(grid, [8, 8, 8]) is None                    # already in grid order
    perm = np.random.default_rng(2).permutation(512)
    order = geometry.grid_permutation(grid[perm], [8, 8, 8])
    assert order is not None and np.array_equal(grid[perm][order], grid) and np.array_equal(perm[order], np.arange(512))
    table = geometry.coupling_table(coords[perm][order], grid[perm][order], [8, 8, 8], "cos")
    assert table is not None and geometry.coupling_table(coords[perm], grid[perm], [8, 8, 8], "cos") is None
    assert geometry.grid_permutation(grid[perm][:300], [8, 8, 8]) is None       # a random subset of the grid is not a grid
    holes = grid[perm].copy(); holes[0] = holes[1]
    assert geometry.grid_permutation(holes, [8, 8, 8]) is None                   # not a permutation
    assert geometry.grid_permutation(grid[perm] * 0.1, [8, 8, 8]) is None        # coordinates, not grid indices


def test_compiled_rank_lists_cover_the_shipped_operator():
    """The register-resident spectral kernels are compiled for fixed lists of modes per parity sector
    (csrc/step_f32_warp.cu, step_f32_warp1.cu, step_f32_oct.cu).  The lists must cover what geometry.spectral_factors finds for
    the shipped cos(distance) kernel at the default truncation (1e-9) on every grid they are meant for -- otherwise the library
    silently keeps a slower kernel -- and batched.HALF_GRID_RANK_SETS must restate the half-grid lists."""
    import os
    import re
    from dbsgym_b200 import batched, geometry
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dbsgym_b200", "csrc")

    def lists(fname, pattern):
        text = open(os.path.join(csrc, fname)).read()
        return [tuple(int(v) for v in m.split(",")) for m in re.findall(pattern, text)]

    def ranks(gx, gy, gz):
        coords, grid = geometry.neuron_grid(gx, gy, gz, gx * gy * gz, 0.1)
        table = geometry.coupling_table(coords, grid, [gx, gy, gz], "cos")
        return geometry.spectral_factors(table, gx, gy, gz, tol=1e-9)[2]

    def covered(r, sets):
        return any(all(a <= b for a, b in zip(r, s)) for s in sets)

    full = lists("step_f32_warp.cu", r"RankSet<([0-9, ]+)>")
    half = lists("step_f32_warp1.cu", r"RankSet<([0-9, ]+)>")
    assert covered(ranks(8, 8, 8), full) and covered(ranks(8, 8, 4), half)
    assert sorted(set(half)) == sorted(batched.HALF_GRID_RANK_SETS)
    octs = re.findall(r"\{(\d+), \{([0-9, ]+)\}\}", open(os.path.join(csrc, "step_f32_oct.cu")).read())
    octs = [(int(n), tuple(int(v) for v in r.split(","))) for n, r in octs]
    assert len(octs) == 7 and all(sum(r) % 2 == 0 for _, r in octs)            # (the expansion takes the modes in pairs)
    launched = lists("step_f32_oct.cu", r"launch_o<RankSet<([0-9, ]+)>")
    assert launched == [r for _, r in octs]                                    # table and dispatch switch in the same order
    for g in ((16, 16, 4), (16, 16, 8), (16, 16, 16), (8, 8, 16), (8, 8, 32), (8, 8, 64), (32, 32, 8)):
        n = g[0] * g[1] * g[2]
        assert covered(ranks(*g), [r for m, r in octs if m == n]), g
