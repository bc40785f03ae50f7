"""The float64-rounding dependent time grid of ``SpatialKuramoto.step``.

``np.arange(t, t + 0.15, 0.05)`` has 3 *or* 4 points and ``np.arange(t, t + 0.75, 0.05)`` 15 *or*
16 depending on how ``t`` rounds (reference environment/env.py:426-441), and ``current_time``
chains through the last element.  The grid is a pure function of the step index, identical for
every environment, so it is replayed once on the host WITH NUMPY ITSELF and shipped to the GPU
as a table of per-step sample offsets.
"""
from __future__ import annotations

import numpy as np


def transient_grid(transient_state_len, verbose_dt):
    """env.py:606-609 -- sample times of the reset transient (starts at current_time = 0)."""
    return np.arange(0.0, transient_state_len, verbose_dt)


class StepSchedule:
    def __init__(self, n_steps, t_start, electrode_width, electrode_pause, verbose_dt):
        grids = []
        ct = t_start
        for _ in range(n_steps):
            seg_on = np.arange(ct, ct + electrode_width, verbose_dt)
            ct = seg_on[-1]
            seg_off = np.arange(ct, ct + electrode_pause, verbose_dt)
            ct = seg_off[-1]
            grids.append((seg_on, seg_off))
        self.n_steps = n_steps
        self.n_I = np.array([len(a) for a, _ in grids], dtype=np.int32)
        self.n_II = np.array([len(b) for _, b in grids], dtype=np.int32)
        self.max_I, self.max_II = int(self.n_I.max()), int(self.n_II.max())
        self.offs_I = np.zeros((n_steps, self.max_I))
        self.offs_II = np.zeros((n_steps, self.max_II))
        self.t_after = np.empty(n_steps)
        for k, (a, b) in enumerate(grids):
            self.offs_I[k, :len(a)] = a - a[0]
            self.offs_II[k, :len(b)] = b - b[0]
            self.t_after[k] = b[-1]
        self.max_samples = int((self.n_I + self.n_II - 1).max())

    def samples_in_step(self, k):
        return int(self.n_I[k] + self.n_II[k] - 1)
