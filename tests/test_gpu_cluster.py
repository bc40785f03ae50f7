"""Cluster mode of the step kernel (one environment = a thread-block cluster, N > 4096)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(N, gz, B, force_cluster=None, precision="f32", gx=8, gy=8, mw=None, sectors=False, no_warp_kernel=False):
    from dbsgym_b200.engine import KuramotoEngine
    from dbsgym_b200.geometry import coupling_table, distances_from, neuron_grid
    from dbsgym_b200.schedule import StepSchedule, transient_grid
    options = {"force_cluster": force_cluster or 0, "mw": mw, "no_warp_kernel": no_warp_kernel}
    coords, grid = neuron_grid(gx, gy, gz, N, 0.1)
    table = coupling_table(coords, grid, [gx, gy, gz], "cos")
    assert table is not None
    eng = KuramotoEngine(B, N, [gx, gy, gz], 2340, 0.52, precision=precision, coupling_table=table, options=options)
    if sectors:                                  # sector form of the low-rank operator: before any vector is uploaded
        from dbsgym_b200.geometry import grid_sector_factors
        f = grid_sector_factors(table, gx, gy, gz, tol=1e-9)
        assert f is not None and f[3] < 2e-9 * np.abs(f[2]).max()
        eng.set_coupling_lowrank_sectors(*f)
    tt = transient_grid(200.0, 0.05)
    sched = StepSchedule(80, tt[-1], 0.15, 0.75, 0.05)
    eng.set_schedule(sched)
    eng.set_reward("bbpow_action", 0.05)
    eng.set_recording(True)
    rng = np.random.default_rng(N + B)
    centre = int(np.argmin(np.abs(grid - np.array([gx // 2, gy // 2 - 1, gz // 2])).sum(axis=1)))
    stim = np.tile(np.maximum(0.0, 1.0 - distances_from(grid * 0.1, [centre])[0]), (B, 1))
    rec = np.tile(np.maximum(0.0, 1.0 - distances_from(grid * 0.1, [centre // 2])[0]), (B, 1))
    w0 = np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02
    y0 = rng.normal(np.pi, 0.6, (B, N)) + 25.0
    eng.set_env_params(None, w0=w0, stim=stim, rec=rec, y0=y0)
    eng.set_window(rng.uniform(-0.2, 0.2, (B, 2340)))
    eng.set_episode(None, step_idx=0, episode_len=1000)
    return eng, dict(table=table, grid=grid, w0=w0, stim=stim, rec=rec, y0=y0, sched=sched, tt=tt)


@pytest.mark.parametrize("N,gz,C,gx,gy", [(1024, 16, 2, 8, 8), (1024, 16, 4, 8, 8), (4096, 64, 8, 8, 8), (2048, 32, 2, 8, 8),
                                         (4096, 16, 2, 16, 16), (4096, 4, 4, 32, 32)])
def test_cluster_mode_equals_single_cta_mode(N, gz, C, gx, gy):
    """Same inputs through the single-CTA kernel and through the C-CTA cluster kernel (forced at a size both can run):
    steps, a transient with rejections, LFP, rewards and counters must agree (only reduction orders differ)."""
    acts = np.random.default_rng(1).uniform(-1, 1, (3, 3)).astype(np.float32)
    res = {}
    for mode in (None, C):
        eng, d = _engine(N, gz, 3, mode, gx=gx, gy=gy)
        eng.counters(reset=True)
        out = []
        for a in acts:
            obs, rew, done = eng.step_host(a)
            out.append((eng.state().copy(), obs.copy(), rew.copy(), eng.lfp()[0].copy(), eng.lfp()[1].copy()))
        eng.transient(np.arange(0.0, 130.0, 0.05))
        out.append((eng.state().copy(), eng.obs_host().copy()))
        res[mode] = (out, eng.counters())
        eng.close()
    (a, ca), (b, cb) = res[None], res[C]
    assert ca == cb and ca["status"] == 0
    assert ca["rejected"] > 0 or gy == 32            # (the flat 32 x 32 x 4 slab integrates its transient without a rejection)
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            np.testing.assert_allclose(u, v, rtol=0, atol=2e-5 if u.ndim == 2 and u.shape[1] >= 1024 else 5e-6)


@pytest.mark.parametrize("gx,gy,gz", [(8, 8, 128), (32, 32, 8)])
def test_n8192_cluster_step_matches_oracle(gx, gy, gz):
    """N = 8192 (8 x 8 x 128 grid, or the first 8 z-planes of the cubic 32 x 32 x 32 grid with four threads per
    32-oscillator line; 2 CTAs per environment): one step() against the fp64 oracle integrator with a chunked evaluation
    of the same coupling operator."""
    from oracle.diffrax_restated import Dopri5, ODETerm, PIDController, SaveAt, diffeqsolve
    N = 8192
    eng, d = _engine(N, gz, 2, gx=gx, gy=gy)
    assert eng.step_variant() == 5
    a = np.array([0.7, -0.4], dtype=np.float32)
    obs, rew, done = eng.step_host(a)
    y_gpu = eng.state()
    lfp_t, lfp_r, ns = eng.lfp()
    c = eng.counters()
    assert c["status"] == 0 and c["rhs_evals"] == 2 * 32
    table, grid = d["table"].reshape(gz, gx, gy), d["grid"]

    def coupled(v):                                    # alpha @ v, alpha_ij = table[|dz|,|dx|,|dy|], in row chunks
        out = np.empty((N, v.shape[1]))
        for lo in range(0, N, 512):
            dd = np.abs(grid[lo:lo + 512, None, :] - grid[None, :, :])
            out[lo:lo + 512] = table[dd[..., 2], dd[..., 0], dd[..., 1]] @ v
        return out

    sched = d["sched"]
    for e in range(1):                                 # one environment on the CPU (32 chunked RHS evaluations)
        u = -5 + (10 * (float(a[e]) + 1)) / 2
        y = d["y0"][e].copy()
        segs = [(sched.offs_I[0, :sched.n_I[0]], u), (sched.offs_II[0, :sched.n_II[0]], 0.0)]
        rows = []
        for ts, amp in segs:
            pulse = amp * d["stim"][e]

            def rhs(t, yy, args, pulse=pulse):
                th = np.fmod(yy, 2 * np.pi)
                sc = coupled(np.stack([np.sin(th), np.cos(th)], axis=1))
                return d["w0"][e] + (0.52 / N) * (np.cos(th) * sc[:, 0] - np.sin(th) * sc[:, 1]) + pulse
            sol = diffeqsolve(ODETerm(rhs), Dopri5(), t0=ts[0], t1=ts[-1], dt0=0.05, y0=y, saveat=SaveAt(ts=ts),
                              stepsize_controller=PIDController(rtol=1e-5, atol=1e-5))
            y = sol.ys[-1]
            rows.append(sol.ys)
        allrows = np.concatenate(rows)[:-1]
        assert np.max(np.abs(y_gpu[e] - y)) < 1e-5
        n = ns[e]
        np.testing.assert_allclose(lfp_t[e, :n], np.mean(np.cos(allrows), axis=1), rtol=0, atol=2e-6)
        np.testing.assert_allclose(lfp_r[e, :n], np.mean(np.cos(allrows) * d["rec"][e], axis=1), rtol=0, atol=2e-6)
    eng.close()


@pytest.mark.parametrize("N,gz,C,gx,gy", [(1024, 16, 2, 8, 8), (4096, 16, 4, 16, 16), (8192, 128, None, 8, 8), (8192, 8, None, 32, 32)])
def test_low_rank_kernel_on_grid_handles_single_cta_and_cluster(N, gz, C, gx, gy):
    """The coupling operator of a regular grid in its truncated eigenbasis (geometry.grid_lowrank_factors: factorised sector
    by sector, never forming the N x N matrix), step-kernel variant 11, against the exact parity-sector kernels on the same
    GRID handle: one CTA per environment, forced clusters of C CTAs (the mode sums of the CTAs meet in global memory), and
    N = 8192 where the cluster is the only way to run -- in the plain form and in the sector form (oscillators stored in
    octant order inside the library, eigenvectors over the octant only).  Steps at float32 rounding, counters exact."""
    from dbsgym_b200.geometry import grid_lowrank_factors
    acts = np.random.default_rng(1).uniform(-1, 1, (3, 3)).astype(np.float32)
    res = {}
    modes = ([("exact", None, False), ("lowrank", None, True), ("sectors", None, "sectors"), ("sectors_block", None, "sectors")] +
             ([("lowrank_cluster", C, True), ("sectors_cluster", C, "sectors")] if C else []))
    for name, force, lr in modes:
        eng, d = _engine(N, gz, 3, force, gx=gx, gy=gy, sectors=(lr == "sectors"), no_warp_kernel=(name == "sectors_block"))
        if lr == "sectors":
            # one CTA per environment and 1024 ... 4096 oscillators: the register-resident kernel (oct_kernel.cuh, variant 13)
            # when a compiled rank list covers the operator; the block kernel (variant 11) with no_warp_kernel, in clusters, above
            # (32 x 32 x 8: the same kernel as a cluster of 2 CTAs; 8 x 8 x 128 has 89 modes, more than any compiled list)
            oct = name == "sectors" and (N <= 4096 or gx == 32)
            assert eng.step_variant() == (13 if oct else 11) and eng.lowrank["sectors"]
        elif lr:
            f = grid_lowrank_factors(d["table"], gx, gy, gz, tol=1e-9)
            assert f is not None and f[0].shape[0] <= 128 and f[2] < 2e-9 * np.abs(f[1][0])
            eng.set_coupling_lowrank(*f)
            assert eng.step_variant() == 11
        eng.counters(reset=True)
        out = []
        for a in acts:
            obs, rew, done = eng.step_host(a)
            out.append((eng.state().copy(), obs.copy(), rew.copy(), eng.lfp()[0].copy(), eng.lfp()[1].copy()))
        res[name] = (out, eng.counters())
        if (N <= 4096 or gx == 32) and name in ("exact", "sectors", "sectors_block"):       # a reset transient (with rejections) as well
            eng.transient(np.arange(0.0, 125.0, 0.05))
            res[name + "_transient"] = (eng.state().copy(), eng.obs_host().copy(), eng.counters())
        eng.close()
    if N <= 4096 or gx == 32:
        y_ref, o_ref, c_ref = res["exact_transient"]
        for name in ("sectors", "sectors_block"):
            y, o, c = res[name + "_transient"]
            assert c["status"] == 0 and abs(c["accepted"] - c_ref["accepted"]) <= 6
            assert c["rejected"] > 0 or gx == 32         # (the flat 32 x 32 x 8 slab integrates its transient without a rejection)
            # (free-running 125 units: float32 rounding differences are amplified by the dynamics, cf. the lines-of-16 test)
            assert np.max(np.abs(y - y_ref)) < 0.25 and np.max(np.abs(o - o_ref)) < 5e-3
    ref, cref = res["exact"]
    assert cref["status"] == 0
    for name in [m[0] for m in modes[1:]]:
        out, c = res[name]
        assert c == cref, (name, c, cref)
        for k, (x, y) in enumerate(zip(ref, out)):
            for u, v in zip(x, y):
                np.testing.assert_allclose(u, v, rtol=2e-5, atol=1e-5 * (k + 1) if u.ndim == 2 and u.shape[1] == N else 5e-6)


def test_ragged_cloud_above_8192_oscillators_runs_in_low_rank_cluster_mode_and_matches_oracle():
    """The first 9000 rows of a 21 x 21 x 21 grid (odd extents, a partial last plane: the shape of BASELINE configs[4]'s
    "first 65536 rows of 41^3" point, at a size the CPU oracle can integrate): a DENSE handle above 8192 oscillators keeps no
    matrix -- the operator comes from the coordinates as eigenpairs (geometry.lowrank_factors_points) and a cluster of 4 CTAs
    (16384 slots, 7384 of them inert padding) integrates it.  One step() against the fp64 oracle integrator with a chunked
    evaluation of alpha = cos(distance)."""
    from dbsgym_b200.engine import KuramotoEngine
    from dbsgym_b200.geometry import coupling_rows, distances_from, lowrank_factors_points, neuron_grid
    from dbsgym_b200.schedule import StepSchedule, transient_grid
    from oracle.diffrax_restated import Dopri5, ODETerm, PIDController, SaveAt, diffeqsolve
    N, B = 9000, 2
    coords, grid = neuron_grid(21, 21, 21, N, 0.1)
    f = lowrank_factors_points(coords, "cos", tol=1e-9)
    assert f is not None and f[0].shape[0] < 100 and f[2] < 2e-9 * abs(f[1][0])
    eng = KuramotoEngine(B, N, [21, 21, 21], 2340, 0.52, precision="f32", lowrank=f)
    assert eng.coupling == "dense" and eng.step_variant() == 11
    with pytest.raises(Exception):
        eng._ck(eng.lib.dbsgym_set_coupling_dense(eng._h, None))          # (no full matrix above 8192 oscillators)
    tt = transient_grid(200.0, 0.05)
    sched = StepSchedule(80, tt[-1], 0.15, 0.75, 0.05)
    eng.set_schedule(sched); eng.set_reward("bbpow_action", 0.05); eng.set_recording(True)
    rng = np.random.default_rng(3)
    stim = np.tile(np.maximum(0.0, 1.0 - distances_from(coords, [N // 2])[0]), (B, 1))
    rec = np.tile(np.maximum(0.0, 1.0 - distances_from(coords, [N // 3])[0]), (B, 1))
    w0 = np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02
    y0 = rng.normal(np.pi, 0.6, (B, N)) + 25.0
    eng.set_env_params(None, w0=w0, stim=stim, rec=rec, y0=y0)
    eng.set_window(rng.uniform(-0.2, 0.2, (B, 2340)))
    eng.set_episode(None, step_idx=0, episode_len=1000)
    assert np.max(np.abs(eng.state() - y0)) < 4e-6                        # (wrapped float32 phase + winding count)
    a = np.array([0.7, -0.4], dtype=np.float32)
    obs, rew, done = eng.step_host(a)
    y_gpu = eng.state()
    lfp_t, lfp_r, ns = eng.lfp()
    c = eng.counters()
    assert c["status"] == 0 and c["rhs_evals"] == B * 32 and c["rejected"] == 0

    def coupled(v):                                    # alpha @ v in row chunks
        out = np.empty((N, v.shape[1]))
        for lo in range(0, N, 1000):
            out[lo:lo + 1000] = coupling_rows(coords, np.arange(lo, min(lo + 1000, N)), "cos") @ v
        return out

    e = 0
    u = -5 + (10 * (float(a[e]) + 1)) / 2
    y = y0[e].copy()
    rows = []
    for ts, amp in [(sched.offs_I[0, :sched.n_I[0]], u), (sched.offs_II[0, :sched.n_II[0]], 0.0)]:
        pulse = amp * stim[e]

        def rhs(t, yy, args, pulse=pulse):
            th = np.fmod(yy, 2 * np.pi)
            sc = coupled(np.stack([np.sin(th), np.cos(th)], axis=1))
            return w0[e] + (0.52 / N) * (np.cos(th) * sc[:, 0] - np.sin(th) * sc[:, 1]) + pulse
        sol = diffeqsolve(ODETerm(rhs), Dopri5(), t0=ts[0], t1=ts[-1], dt0=0.05, y0=y, saveat=SaveAt(ts=ts),
                          stepsize_controller=PIDController(rtol=1e-5, atol=1e-5))
        y = sol.ys[-1]
        rows.append(sol.ys)
    allrows = np.concatenate(rows)[:-1]
    assert np.max(np.abs(y_gpu[e] - y)) < 1e-5
    n = ns[e]
    np.testing.assert_allclose(lfp_t[e, :n], np.mean(np.cos(allrows), axis=1), rtol=0, atol=2e-6)
    np.testing.assert_allclose(lfp_r[e, :n], np.mean(np.cos(allrows) * rec[e], axis=1), rtol=0, atol=2e-6)
    eng.close()


@pytest.mark.parametrize("B", [5, 21])
def test_multi_worker_kernel_equals_the_single_environment_kernel(B):
    """options mw=True forces the multi-worker step kernel (8 environments per CTA sharing the precomputed sector
    coefficients, persistent loop, named barriers) at a batch size where the default would use one CTA per
    environment.  The contraction is the same sum in the same order; only the lane order of the warp reductions
    (LFP samples, error norm) differs, so results agree to an ulp of float32 and the counters exactly -- for full
    launches (partially filled CTAs, several environments per worker) and for a partial reset transient."""
    acts = np.random.default_rng(2).uniform(-1, 1, (4, B)).astype(np.float32)
    res = {}
    for mode in ("0", "1"):
        eng, d = _engine(512, 8, B, mw=(mode == "1"))
        assert eng.step_variant() == (4 if mode == "1" else 3)
        eng.counters(reset=True)
        out = []
        for a in acts[:3]:
            obs, rew, done = eng.step_host(a)
            out.append((eng.state().copy(), obs.copy(), rew.copy(), eng.lfp()[0].copy(), eng.lfp()[1].copy(), eng.lfp()[2].copy()))
        ids = [1, B - 2, 3]
        eng.transient(np.arange(0.0, 130.0, 0.05), env_ids=ids)
        out.append((eng.state().copy(), eng.obs_host().copy()))
        obs, rew, done = eng.step_host(acts[3])
        out.append((eng.state().copy(), obs.copy(), rew.copy()))
        res[mode] = (out, eng.counters())
        eng.close()
    (a, ca), (b, cb) = res["0"], res["1"]
    assert ca == cb and ca["status"] == 0 and ca["rejected"] > 0
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            if u.dtype.kind == "i":
                assert np.array_equal(u, v)
            else:
                np.testing.assert_allclose(u, v, rtol=2e-6, atol=2e-6 if u.ndim == 2 and u.shape[1] == 512 else 3e-7)


@pytest.mark.parametrize("N,G,variant", [(1024, 16, 7), (4096, 16, 7), (4096, 32, 8)])
def test_lines_of_16_and_32_cubic_grids_match_dense_path(N, G, variant):
    """First N rows of the G x G x G grid (the cubic grids of SURVEY.md 8d config 5): the structured GRID_SYM kernel with
    two / four threads per 16- / 32-oscillator line against the generic DENSE kernel on the same alpha = cos(distance),
    in fp32 and (where it fits) against fp64 DENSE; steps and a transient with rejections."""
    from dbsgym_b200.engine import KuramotoEngine
    from dbsgym_b200.geometry import coupling_rows, coupling_table, distances_from, neuron_grid
    from dbsgym_b200.schedule import StepSchedule, transient_grid
    B = 3
    coords, grid = neuron_grid(G, G, G, N, 0.1)
    table = coupling_table(coords, grid, [G, G, G], "cos")
    assert table is not None and table.size == N
    alpha = coupling_rows(coords, np.arange(N), "cos")
    rng = np.random.default_rng(N)
    centre = int(np.argmin(np.abs(grid - np.array([G // 2, G // 2 - 1, grid[:, 2].max() // 2])).sum(axis=1)))
    stim = np.tile(np.maximum(0.0, 1.0 - distances_from(grid * 0.1, [centre])[0]), (B, 1))
    w0 = np.abs(rng.normal(0.6, 0.4, (B, N))) + 0.02
    y0 = rng.normal(np.pi, 0.6, (B, N)) + 12.0
    win = rng.uniform(-0.2, 0.2, (B, 2340))
    acts = rng.uniform(-1, 1, (2, B)).astype(np.float32)
    tt = transient_grid(200.0, 0.05)
    res = {}
    from dbsgym_b200.geometry import lowrank_factors
    lr = lowrank_factors(alpha, tol=1e-9)
    assert lr is not None and lr[0].shape[0] <= 64 and lr[2] < 1e-8 * np.abs(lr[1][0])
    engines = [("grid", dict(coupling_table=table), "f32"), ("dense", dict(alpha=alpha), "f32"), ("lowrank", dict(lowrank=lr), "f32")]
    if N <= 2048:                                     # (the fp64 DENSE kernel keeps 11 N doubles in shared memory)
        engines.append(("dense64", dict(alpha=alpha), "f64"))
    for name, kw, prec in engines:
        eng = KuramotoEngine(B, N, [G, G, G], 2340, 0.52, precision=prec, **kw)
        eng.set_schedule(StepSchedule(80, tt[-1], 0.15, 0.75, 0.05)); eng.set_reward("bbpow_action", 0.05)
        eng.set_recording(True)
        eng.set_env_params(None, w0=w0, stim=stim, rec=stim, y0=y0)
        eng.set_window(win); eng.set_episode(None, step_idx=0, episode_len=1000)
        if name == "grid":
            assert eng.step_variant() == variant
        if name == "lowrank":
            assert eng.step_variant() == 11 and eng.coupling == "dense"
        eng.counters(reset=True)
        out = []
        for a in acts:
            obs, rew, done = eng.step_host(a)
            out.append((eng.state().copy(), eng.lfp()[1].copy(), rew.copy()))
        if N == 1024:
            eng.transient(np.arange(0.0, 125.0, 0.05))
            out.append((eng.state().copy(), eng.obs_host().copy()))
        res[name] = (out, eng.counters())
        eng.close()
    (g, cg), (d, cd) = res["grid"], res["dense"]
    d64, c64 = res.get("dense64", res["dense"])
    assert cg == cd == c64 and cg["status"] == 0
    lo, clo = res["lowrank"]                          # the DENSE operator in its truncated eigenbasis (variant 11)
    assert clo == c64
    for k, (x, z) in enumerate(zip(lo, d64)):
        for u, w in zip(x, z):
            if k < len(acts):
                np.testing.assert_allclose(u, w, rtol=2e-5, atol=5e-5 if u.ndim == 2 and u.shape[1] == N else 5e-6)
            else:             # 125 free-running time units: float32 rounding differences amplified (measured 0.07 rad on 5 of 3072 phases)
                np.testing.assert_allclose(u, w, rtol=2e-5, atol=0.25 if u.shape[1] == N else 5e-3)
    for k, (x, y, z) in enumerate(zip(g, d, d64)):
        for u, v, w in zip(x, y, z):
            if k < len(acts):                         # single steps: fp32 rounding only
                tol = 5e-5 if u.ndim == 2 and u.shape[1] == N else 5e-6
            else:                                     # 125 time units of free-running transient: fp32 rounding differences
                tol = 2e-2 if u.shape[1] == N else 2e-3   # have grown (tests/test_gpu_episode_stats.py measures the horizon)
            np.testing.assert_allclose(u, w, rtol=2e-5, atol=tol)          # structured fp32 vs fp64 dense
            if k < len(acts):                         # (over the long transient the sequential fp32 sums of the DENSE
                np.testing.assert_allclose(u, v, rtol=2e-5, atol=tol)      #  kernel drift further from fp64 than the structured kernel does)
