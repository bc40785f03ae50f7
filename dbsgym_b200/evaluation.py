"""Batched evaluation with the reference's reported metrics.

``evaluate_hf_dbs`` / ``evaluate_policy_`` / ``calc_psd_for_simple_eval`` of reference
aDBS_RL/evaluate_HF_DBS.py:33-174 run one environment at a time; here every environment of a
:class:`BatchedKuramotoVecEnv` is evaluated in the same pass.  Metrics (per environment, then mean / sd
over environments as the paper table does): beta-band (12.5-21 Hz) power of the concatenated TRUE LFP
after a zero-phase order-2 band-pass and 12-tap spectrum smoothing (evaluate_HF_DBS.py:122-135), the
"energy" sum |action| (evaluate_HF_DBS.py:163), and the episode return.
"""
from __future__ import annotations

import functools

import numpy as np

from .utils import band_pass_envelope


def calc_psd_for_simple_eval(sig_envs, psd_dt, beta_a=12.5, beta_b=21):
    """evaluate_HF_DBS.py:122-135 for every row of ``sig_envs``.  Rows of equal length (a 2-D array: the batched
    case, environments in lockstep) go through ONE filtfilt / rfft / filtfilt along the last axis -- scipy's
    filters act on each row exactly as on a 1-D signal; ragged inputs fall back to the row loop."""
    from scipy.signal import butter, filtfilt
    if isinstance(sig_envs, np.ndarray) and sig_envs.ndim == 2:
        nyq = 0.5 / psd_dt
        b, a = butter(2, [12 / nyq, 30 / nyq], btype="band")       # band_pass_envelope(..., order=2), utils.py:794-816
        filt = filtfilt(b, a, sig_envs, axis=-1)
        ft = np.abs(np.fft.rfft(filt, axis=-1) / filt.shape[-1]) ** 2 * 2
        freq = np.fft.rfftfreq(filt.shape[-1], psd_dt)
        ft = filtfilt([1] * 12, 5, ft, axis=-1)
        return np.sum(ft[:, (freq > beta_a) & (freq < beta_b)], axis=1)
    out = []
    for sig in sig_envs:
        sig = np.asarray(sig)
        filt, _ = band_pass_envelope(sig, 1 / psd_dt, order=2)
        ft = np.abs(np.fft.rfft(filt) / filt.shape[0]) ** 2 * 2
        freq = np.fft.rfftfreq(filt.shape[0], psd_dt)
        ft = filtfilt([1] * 12, 5, ft)
        out.append(np.sum(ft[(freq > beta_a) & (freq < beta_b)]))
    return np.asarray(out)


@functools.lru_cache(maxsize=8)
def bbpow_spec(n, psd_dt, beta_a=12.5, beta_b=21):
    """Everything the device metric (csrc/eval_kernel.cuh) needs for traces of n samples: the band-pass of
    band_pass_envelope(order=2) with scipy's filtfilt start-up values, and the weights w with
    sum(filtfilt([1]*12, 5, ft)[band]) == w @ ft[k_lo:k_lo+len(w)] -- the smoothing + band sum of
    evaluate_HF_DBS.py:130-134, obtained by running scipy's own filtfilt on unit vectors."""
    from scipy.signal import butter, filtfilt, lfilter_zi
    nyq = 0.5 / psd_dt
    b, a = butter(2, [12 / nyq, 30 / nyq], btype="band")
    zi = lfilter_zi(b, a)
    padlen = 3 * max(len(a), len(b))
    freq = np.fft.rfftfreq(n, psd_dt)
    band = np.flatnonzero((freq > beta_a) & (freq < beta_b))
    margin = 24                                    # the two 12-tap passes reach 11 bins to either side
    k_lo, k_hi = max(0, band[0] - margin), min(freq.size - 1, band[-1] + margin)
    unit = np.zeros((freq.size, k_hi - k_lo + 1))
    unit[np.arange(k_lo, k_hi + 1), np.arange(k_hi - k_lo + 1)] = 1.0
    resp = filtfilt([1] * 12, 5, unit, axis=0)     # column j = smoothed spectrum of the unit vector e_{k_lo+j}
    w = resp[band].sum(axis=0)
    if k_lo > 0 and k_hi < freq.size - 1:
        assert w[0] == 0.0 and w[-1] == 0.0, "weight support exceeds the margin"
    return {"b": b, "a": a, "zi": zi, "padlen": padlen, "k_lo": int(k_lo), "weights": w}


def device_bbpow(engine, psd_dt=0.0005):
    """calc_psd_for_simple_eval of every environment's recorded device trace, computed on the GPU."""
    n = engine.trace_lengths()
    if not np.all(n == n[0]):
        raise ValueError("device evaluation needs traces of equal length (environments in lockstep)")
    sp = bbpow_spec(int(n[0]), psd_dt)
    return engine.eval_bbpow(sp["b"], sp["a"], sp["zi"], sp["padlen"], sp["k_lo"], sp["weights"])


def evaluate_batched(model, venv, n_steps=None, deterministic=True, on_device=False):
    """Run ``model.predict`` on all environments of ``venv`` for one episode (or ``n_steps``) and return
    a dict of per-environment arrays plus the paper-style summary (mean, sd with ddof=1).
    ``on_device=True``: the TRUE-LFP trace is recorded by the step kernel and the beta-band power is evaluated by
    the device kernels (no per-step LFP read-back); needs a beta-power reward and environments in lockstep."""
    core = venv.core
    B = venv.num_envs
    if n_steps is None:
        n_steps = int(min(h.total_episode_counts for h in core.hosts))
    obs = venv.reset()
    smax = core.engine.max_step_samples
    if on_device:
        core.engine.trace_begin(n_steps * smax)
        energy, ret = np.zeros(B), np.zeros(B)
        state, starts = None, np.ones(B, dtype=bool)
        for _ in range(n_steps):
            act, state = model.predict(obs, state=state, episode_start=starts, deterministic=deterministic)
            act = np.asarray(act, dtype=np.float32)
            act = np.broadcast_to(act.reshape(-1, 1) if act.size == B else act.reshape(1, 1), (B, 1))
            obs, rew, done, infos = venv.step(act)
            energy += np.abs(act[:, 0])
            ret += rew
            starts = done
        core.engine.trace_end()
        bb = device_bbpow(core.engine)
        tr, ln = core.engine.trace()
        return _eval_result(bb, energy, ret, tr[:, :ln[0]])
    trace = np.empty((B, n_steps * smax))           # TRUE LFP of every environment, concatenated over the steps
    fill = np.zeros(B, dtype=np.int64)
    rows = np.arange(B)[:, None]
    energy = np.zeros(B)
    ret = np.zeros(B)
    state, starts = None, np.ones(B, dtype=bool)
    for _ in range(n_steps):
        act, state = model.predict(obs, state=state, episode_start=starts, deterministic=deterministic)
        act = np.asarray(act, dtype=np.float32)
        act = np.broadcast_to(act.reshape(-1, 1) if act.size == B else act.reshape(1, 1), (B, 1))
        obs, rew, done, infos = venv.step(act)
        t, _, n = core.engine.lfp()
        if np.all(n == n[0]):                        # lockstep: one block copy
            k = int(n[0])
            if np.all(fill == fill[0]):
                trace[:, fill[0]:fill[0] + k] = t[:, :k]
            else:
                trace[rows, fill[:, None] + np.arange(k)[None, :]] = t[:, :k]
            fill += k
        else:
            for i in range(B):
                trace[i, fill[i]:fill[i] + n[i]] = t[i, :n[i]]
            fill += n
        energy += np.abs(act[:, 0])
        ret += rew
        starts = done
    if np.all(fill == fill[0]):
        sig = trace[:, :fill[0]]
        bb = calc_psd_for_simple_eval(sig, psd_dt=0.0005)
    else:
        sig = [trace[i, :fill[i]] for i in range(B)]
        bb = calc_psd_for_simple_eval(sig, psd_dt=0.0005)
    return _eval_result(bb, energy, ret, sig)


def _eval_result(bb, energy, ret, sig):
    sd = (lambda v: float(np.std(v, ddof=1)) if len(v) > 1 else 0.0)
    return {"bbpow": bb, "energy": energy, "episode_return": ret, "true_lfp": sig,
            "summary": {"bbpow_mean": float(bb.mean()), "bbpow_sd": sd(bb), "energy_mean": float(energy.mean()),
                        "energy_sd": sd(energy), "return_mean": float(ret.mean()), "return_sd": sd(ret)}}
