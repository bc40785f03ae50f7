"""Host-side geometry of the oscillator grid and the DBS electrode (numpy, vectorised).

Mirrors what the reference computes with Python loops (paths relative to the reference):
  utils.py:478-497  generate_neuron_grid_3D      -> :func:`neuron_grid`
  utils.py:457-466  create_distance_matrix       -> :func:`distances_from` (only the rows used)
  env.py:219-229    alpha = cos(D) | wavelet(D)  -> :func:`coupling_rows`, :func:`coupling_table`
  env.py:61-171     SimpleDBS                    -> :class:`ElectrodeModel`
  utils.py:30-57    create_directed_stim_masks   -> :func:`sector_masks`
  utils.py:885-891  create_oscillation_locus     -> :func:`locus_mask`
Index conventions are the reference's, quirks included.
"""
from __future__ import annotations

import numpy as np

LINE = 8   # oscillators per grid line handled by one GPU thread (csrc/step_kernel.cuh kRows)


def neuron_grid(gx, gy, gz, n_neurons, coord_modif=0.1):
    """Integer grid points, row ``i = z*gx*gy + x*gy + y`` holding ``[x, y, z]``; first n rows."""
    if n_neurons > gx * gy * gz:
        raise ValueError("Number of neurons should be less than grid size.")
    z, x, y = np.unravel_index(np.arange(n_neurons), (gz, gx, gy))
    grid = np.column_stack([x, y, z]).astype(np.int64)
    return grid * coord_modif, grid


def distances_from(points, rows):
    """Euclidean distance from ``points[rows]`` to every point: [len(rows), N]."""
    pts = np.asarray(points, dtype=np.float64)
    diff = pts[np.atleast_1d(rows)][:, None, :] - pts[None, :, :]
    return np.sqrt(np.einsum("rnk,rnk->rn", diff, diff))


def kernel_values(dist, spatial_kernel, wavelet_amp=1.0, wavelet_steepness=1.0):
    if spatial_kernel == "cos":
        return np.cos(dist)
    if spatial_kernel == "wavelet":
        s = wavelet_steepness
        return wavelet_amp * (-s) * (12 * s ** 4 * dist ** 2 - 8 * s ** 2) * np.exp(-s * dist ** 2) / (2 * np.pi)
    raise ValueError(f"Wrong distance matrix type: {spatial_kernel}")


def coupling_rows(neur_coords, rows, spatial_kernel, wavelet_amp=1.0, wavelet_steepness=1.0):
    return kernel_values(distances_from(neur_coords, rows), spatial_kernel, wavelet_amp, wavelet_steepness)


def grid_structure(neur_grid, grid_size):
    """Return (gx, gy, gz_used) when ``neur_grid`` is the un-shuffled regular grid in whole
    z-planes with gy == LINE, 2 * LINE or 4 * LINE (the layouts the GRID kernels handle), else None."""
    g = np.asarray(neur_grid)
    gx, gy, gz = (int(v) for v in grid_size)
    n = g.shape[0]
    if gy not in (LINE, 2 * LINE, 4 * LINE) or n % (gx * gy) != 0 or n > gx * gy * gz or (n // LINE) % 32 != 0:
        return None            # the GRID kernels want whole warps of 8-oscillator (partial) lines (n % 256 == 0)
    if gy != LINE:             # lines of 16 / 32 (cubic 16^3 / 32^3 grids): fp32 mirror-symmetric kernel
        gzu = n // (gx * gy)
        if gx % 2 or gzu % 2 or ((gx // 2) * (gzu // 2)) % 8 or n > 65536:
            return None
    _, ref = neuron_grid(gx, gy, gz, n, 1.0)
    if g.shape != ref.shape or not np.array_equal(g, ref):
        return None
    return gx, gy, n // (gx * gy)


def grid_permutation(neur_grid, grid_size):
    """``order`` with ``neur_grid[order]`` = the un-shuffled regular grid (row d = z*gx*gy + x*gy + y), when ``neur_grid`` is a
    permutation of the first N rows of that grid (utils.py:483-497 with shuffle=True and n_neurons = whole z-planes); None when it
    is not a permutation of them -- or already in grid order."""
    g = np.asarray(neur_grid)
    gx, gy, gz = (int(v) for v in grid_size)
    n = g.shape[0]
    if g.ndim != 2 or g.shape[1] != 3 or n % (gx * gy) != 0 or n > gx * gy * gz or not np.issubdtype(g.dtype, np.integer):
        return None
    if g.min() < 0 or g[:, 0].max() >= gx or g[:, 1].max() >= gy or g[:, 2].max() >= gz:
        return None
    nat = (g[:, 2] * gx + g[:, 0]) * gy + g[:, 1]            # grid row of every caller row
    order = np.argsort(nat, kind="stable")
    if not np.array_equal(nat[order], np.arange(n)) or np.array_equal(order, np.arange(n)):
        return None
    return order.astype(np.int32)


def coupling_table(neur_coords, neur_grid, grid_size, spatial_kernel, wavelet_amp=1.0,
                   wavelet_steepness=1.0, check_rows=6, tol=1e-12):
    """Block-Toeplitz table ``T[(dz*gx+dx)*gy + dy]`` of the coupling operator, or None when the
    coordinates are not an affine image of the regular grid (then the DENSE path is used).
    Verified against directly computed rows of alpha."""
    st = grid_structure(neur_grid, grid_size)
    if st is None:
        return None
    gx, gy, gzu = st
    n = gx * gy * gzu
    row0 = coupling_rows(neur_coords, [0], spatial_kernel, wavelet_amp, wavelet_steepness)[0]
    table = row0.reshape(gzu, gx, gy).copy()            # offsets from neuron 0 == (dz, dx, dy)
    rng = np.random.default_rng(12345)
    rows = np.unique(np.concatenate([[0, n - 1], rng.integers(0, n, check_rows)]))
    direct = coupling_rows(neur_coords, rows, spatial_kernel, wavelet_amp, wavelet_steepness)
    g = np.asarray(neur_grid)
    for r, a in zip(rows, direct):
        d = np.abs(g - g[r])
        if np.max(np.abs(table[d[:, 2], d[:, 0], d[:, 1]] - a)) > tol:
            return None
    return np.ascontiguousarray(table.reshape(-1))


def contact_index(coord, grid_size):
    # env.py:94,97 -- note grid_size[2]**2 and the axis order (x<->z permuted w.r.t. the grid rows)
    return int(coord[0] * grid_size[2] ** 2 + coord[1] * grid_size[1] + coord[2])


def sector_masks(neur_grid, center, forced_idx):
    g = np.asarray(neur_grid, dtype=np.float64)
    az = np.arctan2(g[:, 1] - center[1], g[:, 0] - center[0])
    third = np.pi / 3
    masks = [(az >= -third) & (az < third), (az >= third) & (az <= np.pi), (az >= -np.pi) & (az < -third)]
    for m in masks:
        m[forced_idx] = True
    return masks


def locus_mask(neur_grid, grid_size, locus_coord, locus_size):
    d = distances_from(np.asarray(neur_grid) * locus_size, [contact_index(locus_coord, grid_size)])[0]
    return np.where(1 - d < 0.0, 0.0, 1.0)


class ElectrodeModel:
    """Contact -> neuron index, stimulation and recording conductances (env.py:61-171)."""

    def __init__(self, grid_size, neur_grid, conduct_modifier, elec_coords, rec_coords, amplitudes,
                 directed_stimulation=False, prc_type="I", naive=False, verbose=False):
        if len(amplitudes) != len(elec_coords):
            raise AssertionError("Number of amplitudes is not equal to number of electrode coordinates!")
        if prc_type not in ("I", "II", "Gaussian", "dummy"):
            raise ValueError("Wrong type of PRC function!")
        self.neur_grid = neur_grid
        self.elec_idxs = [contact_index(c, grid_size) for c in elec_coords]
        self.rec_idxs = [contact_index(c, grid_size) for c in rec_coords]
        scaled = np.asarray(neur_grid) * conduct_modifier
        n = scaled.shape[0]

        def conductance_rows(idxs):
            if naive:
                return [np.ones(n) for _ in idxs]
            d = distances_from(scaled, idxs) if len(idxs) else np.zeros((0, n))
            return [np.where(1 - row < 0.0, 0, 1 - row) for row in d]

        self.conductances = conductance_rows(self.elec_idxs)
        if verbose and not directed_stimulation:
            for c in self.conductances:
                print(f"DBS affects {np.count_nonzero(c > 0.0)} neurons, min={round(c.min(), 3)} & max={round(c.max(), 3)}")
        self.directional_masks_list = []
        if directed_stimulation:
            stale = self.elec_idxs[-1]            # env.py:128-131 passes the loop variable left over
            self.directional_masks_list = [sector_masks(neur_grid, np.asarray(c), stale) for c in elec_coords]
            self.directional_mask = [m[0] for m in self.directional_masks_list]
            self.conductances = [c * m for c, m in zip(self.conductances, self.directional_mask)]
        self.rec_conductances = conductance_rows(self.rec_idxs)
        self._stim_vec = self._rec_vec = None     # cached device-side vectors (the lists above stay authoritative)

    def stim_vector(self):
        """What step() applies: only the first contact is driven by the 1-d action (env.py:419-423)."""
        if self._stim_vec is None:
            self._stim_vec = np.asarray(self.conductances[0], dtype=np.float64)
        return self._stim_vec

    def rec_vector(self):
        """Sum over recording contacts (env.py:409-411 sums the per-contact means)."""
        if self._rec_vec is None:
            out = np.zeros(len(self.neur_grid))
            for c in self.rec_conductances:
                out = out + c
            self._rec_vec = out
        return self._rec_vec


# ---- spectral form of the coupling operator (generalised mean-field identity) ----------------------------------------
# BASELINE.json's north star asks for the coupling term "through the mean-field identity (a warp-shuffle and block
# reduction of sum e^{i theta})".  That identity is exact only for alpha == 1; the reference's alpha = cos(distance)
# (env.py:219-223) is not uniform, but it is a smooth kernel of a small grid and therefore numerically of very low rank:
#     alpha = sum_m lambda_m v_m v_m^T,   sum_j alpha_ij sin(th_j - th_i) = sum_m lambda_m v_m[i] (cos th_i S_m - sin th_i C_m),
#     S_m = sum_j v_m[j] sin th_j,  C_m = sum_j v_m[j] cos th_j          ("order parameters" weighted by the eigenvectors)
# with 52 of the 512 eigenvalues above 1e-12 |lambda_max| on the shipped 8 x 8 x 8 grid (mode 0 IS the classical mean field).
# alpha commutes with the three reflections of the grid, so every eigenvector lives in one of the 8 parity sectors: the
# 64 x 64 sector blocks are diagonalised separately (at most 9 modes each).
def sector_block(table, gx, gy, gz, s):
    """Parity-sector block s = 4 * [odd in y] + 2 * [odd in z] + [odd in x] of the block-Toeplitz operator
    alpha_ij = table[|dz|, |dx|, |dy|] on a gx*gy*gz grid with even extents: block[a, b] = sum_g chi_s(g) alpha(a, g b) over
    the 8 reflections g, with a, b in the fundamental octant, index a = ((zq * gx/2 + xq) * gy/2 + yq).  In the
    (unnormalised) sector coordinates X_s[b] = sum_g chi_s(g) x[g b] the operator acts as y_s = block[s] @ X_s and
    (alpha x)[g a] = 1/8 sum_s chi_s(g) y_s[a]."""
    T = np.asarray(table, dtype=np.float64).reshape(gz, gx, gy)
    hz, hx, hy = gz // 2, gx // 2, gy // 2
    zq, xq, yq = np.meshgrid(np.arange(hz), np.arange(hx), np.arange(hy), indexing="ij")
    zq, xq, yq = zq.ravel(), xq.ravel(), yq.ravel()
    py, pz, px = (-1.0 if s & 4 else 1.0), (-1.0 if s & 2 else 1.0), (-1.0 if s & 1 else 1.0)
    blk = np.zeros((zq.size, zq.size))
    for mz, sz in ((0, 1.0), (1, pz)):
        zb = gz - 1 - zq if mz else zq
        for mx, sx in ((0, 1.0), (1, px)):
            xb = gx - 1 - xq if mx else xq
            for my, sy in ((0, 1.0), (1, py)):
                yb = gy - 1 - yq if my else yq
                blk += sz * sx * sy * T[np.abs(zq[:, None] - zb[None, :]), np.abs(xq[:, None] - xb[None, :]),
                                        np.abs(yq[:, None] - yb[None, :])]
    return blk


def sector_blocks(table, gx, gy, gz):
    """The 8 parity-sector blocks (see sector_block)."""
    return [sector_block(table, gx, gy, gz, s) for s in range(8)]


def spectral_factors(table, gx, gy, gz, tol=1e-12):
    """Eigen-decomposition of the 8 sector blocks, truncated at ``|lambda| > tol * |lambda|_max``.
    Returns (vecs [8][n_fund][r_max], vals [8][r_max], ranks [8], residual): eigenvectors of unit length (zero padded),
    eigenvalues sorted by magnitude, and ``residual`` = the largest dropped |eigenvalue| = the spectral norm of
    (alpha - truncated alpha)."""
    blocks = sector_blocks(table, gx, gy, gz)
    eig = [np.linalg.eigh(0.5 * (b + b.T)) for b in blocks]
    lam_max = max(np.abs(w).max() for w, _ in eig)
    keep = [np.argsort(-np.abs(w)) for w, _ in eig]
    ranks = [int(np.count_nonzero(np.abs(w) > tol * lam_max)) for w, _ in eig]
    r_max = max(max(ranks), 1)
    n_f = blocks[0].shape[0]
    vecs, vals = np.zeros((8, n_f, r_max)), np.zeros((8, r_max))
    residual = 0.0
    for s, ((w, v), order, r) in enumerate(zip(eig, keep, ranks)):
        vecs[s, :, :r] = v[:, order[:r]]
        vals[s, :r] = w[order[:r]]
        if r < w.size:
            residual = max(residual, float(np.abs(w[order[r]])))
    return vecs, vals, ranks, residual


def spectral_apply(vecs, vals, x, gx, gy, gz):
    """alpha @ x through the truncated sector eigen-decomposition (float64 host restatement of what the CUDA kernel does;
    used by the tests to measure the truncation error against the dense operator)."""
    x = np.asarray(x, dtype=np.float64).reshape(gz, gx, gy)
    hz, hx, hy = gz // 2, gx // 2, gy // 2
    out = np.zeros_like(x)
    for s in range(8):
        py, pz, px = (-1.0 if s & 4 else 1.0), (-1.0 if s & 2 else 1.0), (-1.0 if s & 1 else 1.0)
        X = np.zeros((hz, hx, hy))
        for mz, sz in ((0, 1.0), (1, pz)):
            for mx, sx in ((0, 1.0), (1, px)):
                for my, sy in ((0, 1.0), (1, py)):
                    sub = x[::-1][:hz] if mz else x[:hz]
                    sub = sub[:, ::-1][:, :hx] if mx else sub[:, :hx]
                    sub = sub[:, :, ::-1][:, :, :hy] if my else sub[:, :, :hy]
                    X += sz * sx * sy * sub
        y = (vecs[s] * vals[s]) @ (vecs[s].T @ X.ravel())
        y = y.reshape(hz, hx, hy) / 8.0
        for mz, sz in ((0, 1.0), (1, pz)):
            for mx, sx in ((0, 1.0), (1, px)):
                for my, sy in ((0, 1.0), (1, py)):
                    blk = sz * sx * sy * y
                    blk = blk[::-1] if mz else blk
                    blk = blk[:, ::-1] if mx else blk
                    blk = blk[:, :, ::-1] if my else blk
                    out[(slice(hz, None) if mz else slice(0, hz)), (slice(hx, None) if mx else slice(0, hx)),
                        (slice(hy, None) if my else slice(0, hy))] += blk
    return out.ravel()


def lowrank_factors(alpha, tol=1e-9, max_rank=256):
    """Truncated eigen-decomposition of a symmetric coupling matrix (any neuron ordering): the eigenpairs with
    ``|lambda| > tol * |lambda|_max``.  Returns (vecs [r][N], vals [r], residual) -- orthonormal eigenvectors as rows,
    sorted by |eigenvalue|, and ``residual`` = the spectral norm of (alpha - truncated alpha), measured -- or None when more
    than ``max_rank`` modes would be needed (the operator is not compressible: keep the full matrix).
    A smooth kernel of a compact neuron cloud (cos(distance) with coord_modif = 0.1: env.py:219-223, utils.py:469-475)
    has a few dozen such modes whatever the ordering of the neurons (utils.py:490 shuffle=True only permutes the rows).
    Small matrices are diagonalised in full; larger ones by randomised subspace iteration (the spectrum decays
    geometrically, so three power iterations resolve every kept mode), O(N^2 r) instead of O(N^3)."""
    a = np.asarray(alpha, dtype=np.float64)
    if a.ndim != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("alpha must be a square matrix")
    a = 0.5 * (a + a.T)
    n = a.shape[0]
    if n <= 1024:
        w, v = np.linalg.eigh(a)
    else:
        rng = np.random.default_rng(0)
        k = 96
        while True:
            k = min(k, n)
            q = np.linalg.qr(a @ rng.standard_normal((n, k)))[0]
            for _ in range(3):
                q = np.linalg.qr(a @ q)[0]
            w, u = np.linalg.eigh(q.T @ a @ q)
            v = q @ u
            # the subspace holds every mode above the threshold once a margin of captured modes lies below it
            if k == n or np.count_nonzero(np.abs(w) <= tol * np.abs(w).max()) >= 16:
                break
            if k > 2 * max_rank + 32:
                return None
            k *= 2
    order = np.argsort(-np.abs(w))
    w, v = w[order], v[:, order]
    r = int(np.count_nonzero(np.abs(w) > tol * np.abs(w[0])))
    if r > max_rank:
        return None
    vecs, vals = np.ascontiguousarray(v[:, :r].T), np.ascontiguousarray(w[:r])
    # spectral norm of what was dropped: power iteration on the residual operator
    x = np.random.default_rng(1).standard_normal(n)
    residual = 0.0
    for _ in range(30):
        x /= np.linalg.norm(x)
        y = a @ x - vecs.T @ (vals * (vecs @ x))
        residual = float(np.linalg.norm(y))
        if residual == 0.0:
            break
        x = y
    return vecs, vals, residual


def grid_lowrank_factors(table, gx, gy, gz, tol=1e-9, max_rank=1024):
    """lowrank_factors for the block-Toeplitz operator of a regular grid with even extents WITHOUT forming the N x N matrix:
    every eigenvector of alpha lives in one parity sector (sector_block), so the 8 blocks of size N/8 are factorised one
    after the other and their eigenvectors z are carried back to the grid, v[g a] = chi_s(g) z[a] / sqrt(8) (same
    eigenvalue).  Returns (vecs [r][N] in the natural neuron order i = (z * gx + x) * gy + y, vals [r], residual) or None.
    N = 65536 (32 x 32 x 64) takes 8 blocks of 8192 x 8192 instead of a 34 GB matrix."""
    hz, hx, hy = gz // 2, gx // 2, gy // 2
    zq, xq, yq = np.meshgrid(np.arange(hz), np.arange(hx), np.arange(hy), indexing="ij")
    zq, xq, yq = zq.ravel(), xq.ravel(), yq.ravel()
    per = []
    for s in range(8):
        f = lowrank_factors(sector_block(table, gx, gy, gz, s), tol=tol * 1e-3, max_rank=max_rank)
        if f is None:
            return None
        per.append(f)
    lam_max = max(np.abs(f[1][0]) for f in per if len(f[1]))
    vecs, vals, residual = [], [], 0.0
    for s, (zv, w, res) in enumerate(per):
        py, pz, px = (-1.0 if s & 4 else 1.0), (-1.0 if s & 2 else 1.0), (-1.0 if s & 1 else 1.0)
        keep = np.abs(w) > tol * lam_max
        residual = max(residual, res, float(np.abs(w[~keep]).max()) if np.any(~keep) else 0.0)
        for zvec, lam in zip(zv[keep], w[keep]):
            full = np.zeros((gz, gx, gy))
            z3 = zvec.reshape(hz, hx, hy) / np.sqrt(8.0)
            for mz, sz in ((0, 1.0), (1, pz)):
                for mx, sx in ((0, 1.0), (1, px)):
                    for my, sy in ((0, 1.0), (1, py)):
                        blk = sz * sx * sy * z3
                        blk = blk[::-1] if mz else blk
                        blk = blk[:, ::-1] if mx else blk
                        blk = blk[:, :, ::-1] if my else blk
                        full[(slice(hz, None) if mz else slice(0, hz)), (slice(hx, None) if mx else slice(0, hx)),
                             (slice(hy, None) if my else slice(0, hy))] = blk
            vecs.append(full.ravel())
            vals.append(lam)
    order = np.argsort(-np.abs(np.array(vals)))
    if len(order) > max_rank:
        return None
    return np.ascontiguousarray(np.array(vecs)[order]), np.ascontiguousarray(np.array(vals)[order]), residual


def grid_sector_factors(table, gx, gy, gz, tol=1e-9, max_modes=1024):
    """The operator of a regular grid with even extents as sector-wise eigenpairs over the fundamental octant (the form
    dbsgym_set_coupling_lowrank_sectors takes): returns (soff [9], zvecs [modes][N / 8], vals [modes], residual) with the modes
    sorted by sector s = 4 [odd y] + 2 [odd z] + [odd x], every sector padded to a multiple of 4 modes (zero rows, zero
    eigenvalues), eigenvalues of the sector blocks (sector_block) kept where |lambda| > tol * |lambda|_max; ``residual`` = the
    spectral norm of what was dropped.  None when more than ``max_modes`` would be needed."""
    per = []
    for s in range(8):
        f = lowrank_factors(sector_block(table, gx, gy, gz, s), tol=tol * 1e-3, max_rank=max_modes)
        if f is None:
            return None
        per.append(f)
    lam_max = max(np.abs(f[1][0]) for f in per if len(f[1]))
    soff, rows, vals, residual = [0], [], [], 0.0
    n_f = (gx // 2) * (gy // 2) * (gz // 2)
    for zv, w, res in per:
        keep = np.abs(w) > tol * lam_max
        residual = max(residual, res, float(np.abs(w[~keep]).max()) if np.any(~keep) else 0.0)
        r = int(keep.sum())
        pad = (-r) % 4
        rows.append(zv[keep]); vals.append(w[keep])
        if pad:
            rows.append(np.zeros((pad, n_f))); vals.append(np.zeros(pad))
        soff.append(soff[-1] + r + pad)
    if soff[-1] > max_modes:
        return None
    return (np.array(soff, dtype=np.int32), np.ascontiguousarray(np.concatenate(rows)), np.ascontiguousarray(np.concatenate(vals)),
            residual)


def lowrank_factors_points(neur_coords, spatial_kernel, wavelet_amp=1.0, wavelet_steepness=1.0, tol=1e-9, max_rank=1024,
                           block=4096, use_cuda=None):
    """lowrank_factors for alpha_ij = kernel(|r_i - r_j|) given by the neuron coordinates alone -- for clouds too large to form
    the N x N matrix on the host (the ragged "first 65536 rows of a 41^3 grid" of BASELINE configs[4]: 34 GB) and without the
    reflection symmetry grid_sector_factors needs.  Randomised subspace iteration in float64 with the matrix regenerated block
    by block for every product, on the GPU through torch when one is there (set-up work, seconds), else with numpy.
    Returns (vecs [r][N], vals [r], residual) like lowrank_factors; ``residual`` is the largest captured eigenvalue below the
    threshold (the subspace keeps a margin of them), or None when more than ``max_rank`` modes would be needed."""
    pts = np.ascontiguousarray(np.asarray(neur_coords, dtype=np.float64))
    n = pts.shape[0]
    if n <= 8192:
        return lowrank_factors(coupling_rows(pts, np.arange(n), spatial_kernel, wavelet_amp, wavelet_steepness), tol=tol,
                               max_rank=min(max_rank, 1024))
    torch = None
    if use_cuda is not False:
        try:
            import torch as _t
            if _t.cuda.is_available():
                torch = _t
        except Exception:  # noqa: BLE001
            torch = None
    if torch is not None:
        x = torch.from_numpy(pts).cuda()

        def kern(d):
            if spatial_kernel == "cos":
                return torch.cos(d)
            if spatial_kernel == "wavelet":
                s_ = wavelet_steepness
                return wavelet_amp * (-s_) * (12 * s_ ** 4 * d ** 2 - 8 * s_ ** 2) * torch.exp(-s_ * d ** 2) / (2 * np.pi)
            raise ValueError(f"Wrong distance matrix type: {spatial_kernel}")

        def apply(q):                                 # alpha @ q without ever holding alpha
            out = torch.empty_like(q)
            for lo in range(0, n, block):
                d = torch.cdist(x[lo:lo + block], x, compute_mode="donot_use_mm_for_euclid_dist")
                out[lo:lo + block] = kern(d) @ q
            return out
        randn = lambda k: torch.from_numpy(np.random.default_rng(0).standard_normal((n, k))).cuda()     # noqa: E731
        qr = lambda a: torch.linalg.qr(a)[0]                                                             # noqa: E731
        eigh = lambda b: tuple(t.cpu().numpy() for t in torch.linalg.eigh(0.5 * (b + b.T)))              # noqa: E731
        to_np = lambda a: a.cpu().numpy()                                                                # noqa: E731
        from_np = lambda a: torch.from_numpy(a).cuda()                                                   # noqa: E731
    else:
        def apply(q):
            out = np.empty_like(q)
            for lo in range(0, n, block):
                out[lo:lo + block] = coupling_rows(pts, np.arange(lo, min(lo + block, n)), spatial_kernel, wavelet_amp,
                                                   wavelet_steepness) @ q
            return out
        randn = lambda k: np.random.default_rng(0).standard_normal((n, k))       # noqa: E731
        qr = lambda a: np.linalg.qr(a)[0]                                         # noqa: E731
        eigh = lambda b: np.linalg.eigh(0.5 * (b + b.T))                          # noqa: E731
        to_np = from_np = lambda a: a                                             # noqa: E731
    k = 128
    while True:
        q = qr(apply(randn(k)))
        for _ in range(3):
            q = qr(apply(q))
        w, u = eigh(q.T @ apply(q))
        if np.count_nonzero(np.abs(w) <= tol * np.abs(w).max()) >= 16:
            break
        if k > 2 * max_rank:
            return None
        k *= 2
    order = np.argsort(-np.abs(w))
    w, u = w[order], u[:, order]
    r = int(np.count_nonzero(np.abs(w) > tol * np.abs(w[0])))
    if r > max_rank:
        return None
    vecs = to_np((q @ from_np(np.ascontiguousarray(u[:, :r]))).T)
    return np.ascontiguousarray(vecs), np.ascontiguousarray(w[:r]), float(np.abs(w[r]))
