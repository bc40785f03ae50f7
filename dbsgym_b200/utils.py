"""Host-side helpers with the names callers of the reference's ``environment/utils.py`` import.

Only the non-plotting half of that module is provided (plotting / gif /
logger helpers are out of scope).  Reference lines are cited per function.
"""
from __future__ import annotations

import numpy as np

from .geometry import (contact_index, distances_from, kernel_values, locus_mask, neuron_grid,
                       sector_masks)


def units2sec(x):
    """utils.py:830-832 -- 100 model units per second."""
    return x / 100


def sec2units(x):
    """utils.py:826-828."""
    return x * 100


def calc_beta_band_power(sig, dt, beta_a, beta_b):
    """utils.py:21-27 -- one-sided power summed over rfft bins strictly inside (beta_a, beta_b)."""
    n = sig.shape[0]
    spec = np.fft.rfft(sig) / n
    power = (spec.real ** 2 + spec.imag ** 2) * 2
    f = np.fft.rfftfreq(n, dt)
    return np.sum(power[(f > beta_a) & (f < beta_b)])


def beta_bins(n, dt, beta_a, beta_b):
    """Inclusive (lo, hi) range of the rfft bins :func:`calc_beta_band_power` sums."""
    f = np.fft.rfftfreq(n, dt)
    idx = np.nonzero((f > beta_a) & (f < beta_b))[0]
    if idx.size == 0 or not np.array_equal(idx, np.arange(idx[0], idx[-1] + 1)):
        raise ValueError("beta band selects no contiguous bin range")
    return int(idx[0]), int(idx[-1])


def band_pass_envelope(signal, fs, lowcut=12, highcut=30, order=5):
    """utils.py:794-816 -- zero-phase Butterworth band-pass and its Hilbert envelope."""
    from scipy.signal import butter, filtfilt, hilbert
    nyq = 0.5 * fs
    b, a = butter(order, [lowcut / nyq, highcut / nyq], btype="band")
    filtered = filtfilt(b, a, signal)
    return filtered, np.abs(hilbert(filtered))


def temp_const_functional(window_len, fs, order=2):
    """Vector g with ``x_f[-1] - mean(x_f) == g @ x`` for ``x_f = filtfilt(butter(order, [12,30] Hz))``
    (the bracket of reward R2, env.py:663-664).  filtfilt with its default odd padding and
    ``lfilter_zi`` start-up is linear in x, so g is obtained by filtering the identity."""
    from scipy.signal import butter, filtfilt
    nyq = 0.5 * fs
    b, a = butter(order, [12 / nyq, 30 / nyq], btype="band")
    resp = filtfilt(b, a, np.eye(window_len), axis=0)       # column j = response to e_j
    return resp[-1, :] - resp.mean(axis=0)


def remove_negative_w0(w0):
    """utils.py:819-823 -- in place; draws from the global ``np.random`` stream."""
    bad = np.flatnonzero(w0 <= 0.)
    if bad.size == 0:              # randn(0) draws nothing and the assignment is empty: skip the mean as well
        return w0
    noise = np.random.randn(bad.size) * 0.05
    w0[bad] = np.abs(noise) + np.mean(w0)
    return w0


def apply_locus_mask(w0, w_locus, lmask):
    """utils.py:902-906."""
    return w0 * (lmask * -1 + 1) + w_locus * lmask


def generate_neuron_grid_3D(greed_size_x, greed_size_y, greed_size_z, n_neurons, coord_modif=0.1,
                            shuffle=False):
    """utils.py:478-497."""
    coords, grid = neuron_grid(greed_size_x, greed_size_y, greed_size_z,
                               greed_size_x * greed_size_y * greed_size_z, 1.0)
    if n_neurons > grid.shape[0]:
        raise ValueError("Number of neurons should be less than grid size.")
    if shuffle:
        np.random.shuffle(grid)
    grid = grid[:n_neurons]
    return grid * coord_modif, grid


def create_distance_matrix(neur_coords):
    """utils.py:457-466 (vectorised)."""
    pts = np.asarray(neur_coords)
    return distances_from(pts, np.arange(pts.shape[0]))


def wavelet_kernel_matrix(distances, amplitude, steepness):
    """utils.py:469-475."""
    return kernel_values(distances, "wavelet", amplitude, steepness)


def create_oscillation_locus(neur_grid, grid_size, locus_coord, locus_size):
    """utils.py:885-891."""
    return locus_mask(neur_grid, grid_size, locus_coord, locus_size)


def create_directed_stim_masks(grid_points, center, center_idx):
    """utils.py:41-57."""
    return tuple(sector_masks(grid_points, center, center_idx))


_PDF_X = [0, 1.8, 2.5, 3.3, 4.5, 5.5, 8, 12.5, 18, 20, 22, 25, 30, 35, 40, 45, 50, 55, 60]


def generate_w0_samples(N, lf_peak=6, beta_peak=10, show=False):
    """utils.py:847-882 -- natural frequencies (Hz) by inverse-CDF sampling of a degree-10
    polynomial fitted to a hand-drawn spectrum; consumes ``np.random.rand(N)``."""
    from scipy.integrate import quad
    from scipy.interpolate import interp1d
    pdf_y = [6, 7.7, lf_peak, 7.7, 4, 3.5, 4, 5, 5.7, beta_peak, 5.7, 4.9, 2.3, 1.2, 0.8, 0.75, 0.7, 0.7, 0.68]
    poly = np.poly1d(np.polyfit(_PDF_X, pdf_y, 10))
    density = lambda v: np.maximum(poly(v), 0)     # noqa: E731
    total, _ = quad(density, min(_PDF_X), max(_PDF_X))
    support = np.linspace(min(_PDF_X), 30, 1000)
    cdf = np.cumsum(density(support) / total)
    cdf /= cdf[-1]
    inverse = interp1d(cdf, support, bounds_error=False, fill_value=(support[0], support[-1]))
    return inverse(np.random.rand(N))


def generate_w0_with_locus(n_neurons, grid_size, coord_modif, locus_center, locus_size, wmuL, wsdL,
                           show=True, vertical_layer=4):
    """utils.py:909-942 -- returns (w0, neur_coords, neur_grid, w0_without_locus, w_locus,
    locus_mask); frequencies converted to rad/unit by x0.065."""
    hz = generate_w0_samples(n_neurons, show=False)
    coords, grid = generate_neuron_grid_3D(*grid_size, n_neurons, coord_modif=coord_modif)
    mask = create_oscillation_locus(grid, grid_size, locus_coord=locus_center, locus_size=locus_size)
    locus_hz = np.random.uniform(low=wmuL - wsdL, high=wmuL + wsdL, size=(n_neurons))
    with_locus = apply_locus_mask(hz, locus_hz, mask)
    return with_locus * 0.065, coords, grid, hz * 0.065, locus_hz * 0.065, mask


def contact_to_index(coord, grid_size):
    return contact_index(coord, grid_size)
