"""CPU restatement of the diffrax 0.7.0 pieces the reference calls.

TEST INFRASTRUCTURE ONLY -- this file is part of ``oracle/``.  Nothing in the
product package may import it; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

PARITY UNPINNED: diffrax is a third-party dependency of the reference
(``requirements_pip.txt:12`` pins ``diffrax==0.7.0``) whose source is *not*
under ``/root/reference`` and which is not installable in this image.  This is a
restatement of its published algorithm (Dormand-Prince 5(4) with FSAL, the
``PIDController`` step-size controller with its default I-only coefficients,
``SaveAt(ts=...)`` dense output through the 4th-order Dopri5 interpolant),
anchored on the reference's call site ``environment/env.py:247-249`` (solver
objects) and ``environment/env.py:260-271`` (the ``diffeqsolve`` call).  No
reference test or golden vector pins per-step results, so the
only external anchors are statistical (paper table beta-band power).

The names exported here (``diffeqsolve, Dopri5, ODETerm, SaveAt,
PIDController``) are exactly the ones ``environment/env.py:10`` imports, so
``oracle/shims/diffrax.py`` can re-export them and the reference module runs
verbatim.

Arithmetic follows the dtype of ``y0`` (float64 under the numpy shim; pass a
float32 ``y0`` and ``time_dtype=np.float32`` to emulate the reference's
JAX-default float32 run).
"""
from __future__ import annotations

import numpy as np

# --- Dormand-Prince 5(4) tableau (Dormand & Prince 1980; diffrax Dopri5) -----
_A = [
    [],
    [1 / 5],
    [3 / 40, 9 / 40],
    [44 / 45, -56 / 15, 32 / 9],
    [19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729],
    [9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656],
    [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84],
]
_B_SOL = [35 / 384, 0.0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0.0]
_B_HAT = [1951 / 21600, 0.0, 22642 / 50085, 451 / 720, -12231 / 42400,
          649 / 6300, 1 / 60]
_B_ERR = [bs - bh for bs, bh in zip(_B_SOL, _B_HAT)]
# mid-point coefficients of the Dopri5 dense output (Shampine 1986), halved
_C_MID = [0.5 * v for v in (
    6025192743 / 30085553152,
    0.0,
    51252292925 / 65400821598,
    -2691868925 / 45128329728,
    187940372067 / 1594534317056,
    -1776094331 / 19743644256,
    11237099 / 235043384,
)]

TABLEAU = {"a": _A, "b_sol": _B_SOL, "b_err": _B_ERR, "c_mid": _C_MID}

# process-wide counters (the reference discards ``solution.stats``; the golden
# generator reads these to record accepted / rejected / RHS counts)
GLOBAL_STATS = {"num_accepted_steps": 0, "num_rejected_steps": 0, "num_rhs_evals": 0}


class Dopri5:
    order = 5
    n_stages = 7


class ODETerm:
    def __init__(self, vector_field):
        self.vector_field = vector_field


class SaveAt:
    def __init__(self, ts=None, t0=False, t1=False):
        self.ts = ts
        self.t0 = t0
        self.t1 = t1


class PIDController:
    def __init__(self, rtol, atol, pcoeff=0.0, icoeff=1.0, dcoeff=0.0,
                 safety=0.9, factormin=0.2, factormax=10.0):
        if pcoeff != 0.0 or dcoeff != 0.0:
            raise NotImplementedError("only the I-controller the reference uses")
        self.rtol, self.atol = rtol, atol
        self.icoeff = icoeff
        self.safety, self.factormin, self.factormax = safety, factormin, factormax


class Solution:
    def __init__(self, ts, ys, stats):
        self.ts, self.ys, self.stats = ts, ys, stats


def _clip_to_end(tprev, tnext, t1, keep, tol):
    if tnext > t1 - tol:
        return t1 if keep else tprev + 0.5 * (t1 - tprev)
    return tnext


def diffeqsolve(terms, solver, t0, t1, dt0, y0, args=None, saveat=None,
                stepsize_controller=None, max_steps=4096, time_dtype=None,
                trace=None):
    """Adaptive Dopri5 solve with dense output at ``saveat.ts``.

    ``trace`` (optional list) receives one tuple per attempted step:
    ``(t_start, t_end, accepted, scaled_error)``.
    """
    f = terms.vector_field
    ctl = stepsize_controller
    y = np.array(y0, copy=True)
    ydt = y.dtype if y.dtype.kind == "f" else np.dtype(np.float64)
    y = y.astype(ydt, copy=False)
    tdt = np.dtype(time_dtype) if time_dtype is not None else np.dtype(np.float64)
    tol = 1e-10 if tdt == np.float64 else 1e-6
    T = tdt.type
    R = ydt.type

    ts = np.asarray(saveat.ts, dtype=tdt) if saveat is not None and saveat.ts is not None else None
    n_save = 0 if ts is None else len(ts)
    ys = np.empty((n_save,) + y.shape, dtype=ydt)
    save_idx = 0

    t_end = T(t1)
    tprev = T(t0)
    tnext = min(T(tprev + T(dt0)), t_end)

    a = [[R(v) for v in row] for row in _A]
    b_sol = [R(v) for v in _B_SOL]
    b_err = [R(v) for v in _B_ERR]
    c_mid = [R(v) for v in _C_MID]

    f0 = np.asarray(f(tprev, y, args), dtype=ydt)       # FSAL start (env.py:260: fresh per forward())
    n_rhs, n_acc, n_rej = 1, 0, 0
    steps = 0
    while tprev < t_end:
        if steps >= max_steps:
            raise RuntimeError("max_steps reached")
        steps += 1
        dt = R(tnext - tprev)
        fs = [f0]
        k = [dt * f0]
        for s in range(1, 7):
            ys_ = y.copy()
            for j in range(s):
                if a[s][j] != 0:
                    ys_ = ys_ + a[s][j] * k[j]
            c_s = sum(_A[s])
            fs.append(np.asarray(f(tprev + T(c_s) * T(dt), ys_, args), dtype=ydt))
            k.append(dt * fs[s])
            n_rhs += 1
        y1 = ys_                                          # a[6] == b_sol: last stage input is y1
        y_err = b_err[0] * k[0]
        for j in range(1, 7):
            if b_err[j] != 0:
                y_err = y_err + b_err[j] * k[j]
        scale = R(ctl.atol) + np.maximum(np.abs(y), np.abs(y1)) * R(ctl.rtol)
        r = y_err / scale
        err = np.sqrt(np.mean(r * r))
        keep = bool(err < 1)
        if err == 0:
            inv = np.inf
        else:
            inv = 1.0 / float(err)
        factormin = 1.0 if keep else ctl.factormin
        factor = min(max(ctl.safety * inv ** (ctl.icoeff / 5.0), factormin), ctl.factormax)
        dt_next = T(dt) * T(factor)
        if trace is not None:
            trace.append((float(tprev), float(tnext), keep, float(err)))
        if keep:
            n_acc += 1
            # dense output on [tprev, tnext] for every requested ts <= tnext
            if save_idx < n_save and ts[save_idx] <= tnext:
                ymid = y.copy()
                for j in range(7):
                    if c_mid[j] != 0:
                        ymid = ymid + c_mid[j] * k[j]
                f0i, f1i = k[0], k[6]
                ca = 2 * (f1i - f0i) - 8 * (y1 + y) + 16 * ymid
                cb = 5 * f0i - 3 * f1i + 18 * y + 14 * y1 - 32 * ymid
                cc = f1i - 4 * f0i - 11 * y - 5 * y1 + 16 * ymid
                while save_idx < n_save and ts[save_idx] <= tnext:
                    if tnext == tprev:
                        tau = R(0)
                    else:
                        tau = R((ts[save_idx] - tprev) / (tnext - tprev))
                    ys[save_idx] = (((ca * tau + cb) * tau + cc) * tau + f0i) * tau + y
                    save_idx += 1
            y = y1
            f0 = fs[6]                                   # FSAL: f(y1) reused as the next k1
            new_t0 = tnext
        else:
            n_rej += 1
            new_t0 = tprev
        new_t1 = T(new_t0 + dt_next)
        tprev = min(new_t0, t_end)
        tnext = T(_clip_to_end(tprev, new_t1, t_end, keep, tol))
    stats = {"num_accepted_steps": n_acc, "num_rejected_steps": n_rej,
             "num_rhs_evals": n_rhs, "num_steps": steps}
    for key in GLOBAL_STATS:
        GLOBAL_STATS[key] += stats[key]
    return Solution(ts, ys, stats)
