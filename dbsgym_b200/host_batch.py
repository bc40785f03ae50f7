"""Batched HOST logic of reset(): what ``host_env.HostEnvState.begin_episode`` does for one environment
(reference environment/env.py:483-598, :21-57; utils.py:819-823, :902-906), for thousands at once.

* the state of all environments lives in arrays (event counters, electrode / recording contacts, encapsulation
  coefficient, plasticity walks, natural frequencies);
* every random number of one batched reset is drawn by ONE native call that consumes numpy's legacy global stream in the
  reference's order, environment after environment (csrc/host_rng.cu, np_stream.py), so the batch gets exactly the numbers
  a sequential ``DummyVecEnv`` of reference environments would get and ``np.random`` is left in the same state;
* the arithmetic on those numbers uses the reference's own expressions on whole arrays (element-wise operations and
  row-wise ``np.mean`` / ``np.std`` are bit-identical to their per-environment forms; tests/test_host_batch.py).

The per-environment ``HostEnvState`` objects stay the single source of constants and are refreshed from the arrays
whenever somebody looks at one (``HostList``), so ``hosts[i].elec_coords`` etc. keep working.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .host_env import cached_electrode, stim_rec_table
from .np_stream import NumpyGlobalStream

_COMMON = ("num_oscillators", "grid_size", "random_freq_update", "reset_plasticity_episode", "init_state_mean",
           "init_state_sd", "electrode_amps", "directed_stimulation", "electrode_prc_type", "naive_dbs")


class HostList(list):
    """``hosts[i]`` refreshes HostEnvState i from the batch arrays before handing it out."""

    def __init__(self, hosts, batch):
        super().__init__(hosts)
        self._batch = batch

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self[i] for i in range(*k.indices(len(self)))]
        h = list.__getitem__(self, k)
        self._batch.sync_to_host(k if k >= 0 else k + len(self), h)
        return h

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _Prepared:
    """One batched reset up to (and including) its random draws, not yet applied to the batch."""


def _copy_state(st):
    out = _capi.DbsGymNpState()
    C.memmove(C.byref(out), C.byref(st), C.sizeof(_capi.DbsGymNpState))
    return out


def _same_state(a, b):
    return (a.pos == b.pos and a.has_gauss == b.has_gauss and (not a.has_gauss or a.gauss == b.gauss) and
            bytes(a.key) == bytes(b.key))


class HostBatch:
    @staticmethod
    def supported(hosts):
        """The batched path covers the shipped configurations: one stimulation and one recording contact, no event
        logging to disk, quiet, freshly drawn initial phases, and common values for the keys the native draw routine
        takes as scalars.  Anything else keeps the per-environment path."""
        if not hosts:
            return False
        p0 = hosts[0].params_dict
        for h in hosts:
            p = h.params_dict
            if h.save_init or h.verbose or (p["save_events"] and p["log_path"] is not None):
                return False
            if len(h.elec_coords) != 1 or len(h.rec_coords) != 1:
                return False
            if p["temporal_drift"] != p0["temporal_drift"]:
                return False
            if p["neur_grid"] is not p0["neur_grid"] and not np.array_equal(p["neur_grid"], p0["neur_grid"]):
                return False
            for k in _COMMON:
                if k in ("reset_plasticity_episode", "random_freq_update") and not p0["temporal_drift"]:
                    continue
                a, b = p[k], p0[k]
                if not (a == b if not isinstance(a, (list, np.ndarray)) else np.array_equal(a, b)):
                    return False
        return True

    def __init__(self, hosts, speculate=None):
        self.hosts = hosts
        # prepare the next full reset on a worker thread (see begin_episodes); on by default for batches large enough for
        # the draws to matter
        self.speculate = (len(hosts) >= 256) if speculate is None else bool(speculate)
        self._job = None
        self.prepared_used = self.prepared_dropped = 0
        p0 = self.p0 = hosts[0].params_dict
        B, N = len(hosts), int(p0["num_oscillators"])
        self.B, self.N = B, N
        self.drift = bool(p0["temporal_drift"])
        self.stream = NumpyGlobalStream()
        i32 = lambda f: np.array([f(h) for h in hosts], dtype=np.int32)      # noqa: E731
        self.reset_count = i32(lambda h: h.reset_count)
        self.elec = np.array([h.elec_coords[0] for h in hosts], dtype=np.int32)
        self.rec = np.array([h.rec_coords[0] for h in hosts], dtype=np.int32)
        self.encaps = np.array([h.encapsulation_coeff for h in hosts], dtype=np.float64)
        self.spatial_feature = np.array([bool(h.params_dict["spatial_feature"]) for h in hosts])
        self.spatial_var_freq = i32(lambda h: h.spatial_var_freq)
        self.spatial_var_episode = i32(lambda h: h.spatial_var_episode)
        w_locus = np.array([h.params_dict["locus_without_w0"] for h in hosts], dtype=np.float64)
        lmask = np.array([h.params_dict["locus_mask"] for h in hosts], dtype=np.float64)
        # the two constant factors of apply_locus_mask (utils.py:902-906): w0 * (lmask * -1 + 1) + w_locus * lmask
        self.keep = lmask * -1 + 1
        self.locus_term = w_locus * lmask
        self.wl = np.array([h.w0_without_locus for h in hosts], dtype=np.float64)         # current w0_without_locus
        self.w0 = np.array([h.w0 for h in hosts], dtype=np.float64)
        self.init_state = np.array([h.init_state for h in hosts], dtype=np.float64)
        self.freq = np.zeros((B, 3), dtype=np.int32)
        self.M = 0
        if self.drift:
            self.elec_drift_episode = i32(lambda h: h.elec_drift_episode)
            self.elec_encaps_episode = i32(lambda h: h.elec_encaps_episode)
            self.plasticity_episode = i32(lambda h: h.plasticity_episode)
            self.count = i32(lambda h: h.plasticity_process_count)
            self.freq[:, 0] = [h.params_dict["electrode_drift_freq"] for h in hosts]
            self.freq[:, 1] = [h.params_dict["encapsulation_drift_freq"] for h in hosts]
            self.freq[:, 2] = [h.params_dict["plasticity_drift_freq"] for h in hosts]
            self.encaps_percent = np.array([h.encaps_precent for h in hosts], dtype=np.float64)
            self.step_scale = np.array([h.plasticity_percent * 0.01 for h in hosts], dtype=np.float64)
            self.regen_every = int(p0["reset_plasticity_episode"])
            self.M = 2 * self.regen_every
            self.wl_orig = np.array([h.w0_without_locus_ for h in hosts], dtype=np.float64)
            self.walk = np.array([h.w0_process for h in hosts], dtype=np.float64)         # [B, M + 1, N]
        self.table = stim_rec_table()
        self._electrodes = [None] * B
        self._last_key = np.full((B, 7), np.nan)       # (contacts, conduct_modifier) the device holds stim / rec vectors for
        self._w0_sent = np.zeros(B, dtype=bool)        # the device holds this environment's current w0
        self.lo, self.hi = 1, min(p0["grid_size"]) - 2
        if self.speculate:             # the first batched reset too (a VecEnv's first auto-reset, or the reset() right after construction)
            self._start_prepare(np.arange(B, dtype=np.int64), self.stream.pull().st)

    # ------------------------------------------------------------------------------------------------------
    def sync_to_host(self, i, h):
        """Write environment i's mutable fields back into its HostEnvState (attribute surface of env.py:339-386, :483-557)."""
        h.reset_count = int(self.reset_count[i])
        h.elec_coords = [self.elec[i].tolist()]
        h.rec_coords = [self.rec[i].tolist()]
        h.encapsulation_coeff = float(self.encaps[i])
        h.spatial_var_episode = int(self.spatial_var_episode[i])
        h.w0_without_locus = self.wl[i]
        h.w0 = self.w0[i]
        h.init_state = self.init_state[i]
        if self.drift:
            h.elec_drift_episode = int(self.elec_drift_episode[i])
            h.elec_encaps_episode = int(self.elec_encaps_episode[i])
            h.plasticity_episode = int(self.plasticity_episode[i])
            h.plasticity_process_count = int(self.count[i])
            h.w0_process = self.walk[i]

    # ------------------------------------------------------------------------------------------------------
    def begin_episodes(self, ids, changed_only=False):
        """env.py:467-598 for the listed environments, in that order.  Returns (w0, stim, rec, y0, electrodes): arrays
        [n, N] float64 and the ElectrodeModel of every environment.  With ``changed_only`` the tuple continues with two
        boolean masks (el_changed, w0_changed) over the listed environments and ``stim`` / ``rec`` hold only the rows of the
        environments whose electrode (contacts or conduct_modifier) differs from what the caller was given last time -- most
        resets change neither, and the conductance vectors are 2 x 4 KB per environment to build and upload.

        The draws of a reset depend only on the host state the previous reset left and on numpy's global stream, so after a
        reset of EVERY environment (synchronous episodes) the next one is prepared ahead of time on a worker thread
        (``speculate``): when the next call again lists every environment and ``np.random`` still is in the state the
        last reset left it in, its numbers are taken from there -- the same numbers, the same final stream state; in any
        other case the prepared reset is dropped and the call draws as usual."""
        ids = np.asarray(ids, dtype=np.int64)
        prep = self._take_prepared(ids)
        if prep is None:
            prep = self._prepare(ids, self.stream.pull().st)
        out = self._commit(prep, changed_only)
        if self.speculate and prep.everyone:
            self._start_prepare(ids, prep.end_state)
        return out

    # ---- speculation ---------------------------------------------------------------------------------------
    def _start_prepare(self, ids, state):
        import threading
        job = {"ids": ids, "state": _copy_state(state), "prep": None, "error": None}

        def work():
            try:
                job["prep"] = self._prepare(ids, _copy_state(job["state"]))
            except BaseException as e:          # (reported by the caller's thread when it takes the result)
                job["error"] = e
        job["thread"] = threading.Thread(target=work, name="dbsgym-reset-prepare", daemon=True)
        job["thread"].start()
        self._job = job

    def _take_prepared(self, ids):
        """The reset prepared ahead of time, if it is the one being asked for (same environments, numpy's global stream
        untouched since); waits for the worker either way, because the caller is about to change the state it reads."""
        job, self._job = self._job, None
        if job is None:
            return None
        job["thread"].join()
        if job["error"] is not None or job["prep"] is None or not np.array_equal(ids, job["ids"]):
            self.prepared_dropped += 1
            return None
        if not _same_state(self.stream.pull().st, job["state"]):
            self.prepared_dropped += 1
            return None
        self.prepared_used += 1
        return job["prep"]

    def close(self):
        job, self._job = self._job, None
        if job is not None:
            job["thread"].join()

    # ---- phase 1: everything up to and including the draws; does NOT modify the batch ---------------------
    def _prepare(self, ids, state):
        n, N = ids.size, self.N
        pr = _Prepared()
        pr.ids, pr.n = ids, n
        pr.everyone = everyone = n == self.B and np.array_equal(ids, np.arange(self.B))
        sel = slice(None) if everyone else ids                     # (views instead of gathered copies for a full reset)
        pr.rc = rc = self.reset_count[ids] + 1
        flags = np.zeros(n, dtype=np.uint8)
        pr.f_en = pr.f_pl = pr.f_rg = None
        if self.drift:
            f_el = self.elec_drift_episode[ids] == rc
            pr.f_en = f_en = self.elec_encaps_episode[ids] == rc
            pr.f_pl = f_pl = self.plasticity_episode[ids] == rc
            pr.f_rg = f_rg = (rc % self.regen_every) == 0
            flags |= (f_el * _capi.RESET_ELECTRODE_MOVE + f_en * _capi.RESET_ENCAPSULATION + f_pl * _capi.RESET_PLASTICITY +
                      f_rg * _capi.RESET_WALK_REGEN).astype(np.uint8)
        pr.f_sp = f_sp = self.spatial_feature[ids] & (self.spatial_var_episode[ids] == rc) & (rc > 2)
        flags |= (f_sp * _capi.RESET_SPATIAL).astype(np.uint8)

        # natural frequencies of this episode (known before any draw: the plasticity event takes an entry of the walk
        # generated EARLIER, a walk regeneration goes back to the original vector; env.py:519-541, :566)
        w0 = self.wl[sel] * self.keep[sel]                                # utils.py:902-906 apply_locus_mask
        w0 += self.locus_term[sel]
        pr.wl_rows, pr.wl_vals = np.empty(0, dtype=np.int64), None        # rows whose w0_without_locus changes, new vectors
        if self.drift and (f_pl.any() or f_rg.any()):
            rows = np.flatnonzero(f_pl | f_rg)
            who = ids[rows]
            vals = np.where(f_rg[rows][:, None], self.wl_orig[who], self.walk[who, np.minimum(self.count[who], self.M)])
            pr.wl_rows, pr.wl_vals = rows, vals
            fix = vals * self.keep[who]
            fix += self.locus_term[who]
            w0[rows] = fix
        bad = w0 <= 0.0
        n_fix = np.count_nonzero(bad, axis=1).astype(np.int32)

        # ---- every draw of this reset, in the reference's order, from numpy's global stream ----
        n_regen = int(f_rg.sum()) if self.drift else 0
        pr.elec = elec = np.ascontiguousarray(self.elec[ids])
        pr.inc = inc = np.zeros((n, 3), dtype=np.int32)
        pr.pick = pick = np.empty(n, dtype=np.int32)
        fix_noise = np.empty(max(int(n_fix.sum()), 1))
        pr.walk_noise = walk_noise = np.empty((max(n_regen, 1), max(self.M, 1), N)) if n_regen else None
        y0 = np.empty((n, N))
        cap_rows, cap_noise = 64, 512
        refix_env = np.zeros((cap_rows, 2), dtype=np.int32)
        refix_noise = np.empty(cap_noise)
        n_refix = C.c_int32(0)
        plan = _capi.DbsGymResetPlan()
        plan.struct_bytes = C.sizeof(_capi.DbsGymResetPlan)
        plan.n_envs, plan.n_osc, plan.walk_len = n, N, self.M
        plan.coord_lo, plan.coord_hi, plan.table_len = self.lo, self.hi, len(self.table)
        plan.random_freq_update = 1 if (self.drift and self.p0["random_freq_update"]) else 0
        plan.refix_cap_rows, plan.refix_cap_noise = cap_rows, cap_noise
        plan.init_mean, plan.init_sd = float(self.p0["init_state_mean"]), float(self.p0["init_state_sd"])
        freq = np.ascontiguousarray(self.freq[ids])
        rcode = self.stream.lib.dbsgym_np_reset_draws(
            C.byref(state), C.byref(plan), _capi.ptr(flags), _capi.ptr(freq), _capi.ptr(elec), _capi.ptr(inc),
            _capi.ptr(pick), _capi.ptr(n_fix), _capi.ptr(fix_noise), _capi.ptr(walk_noise) if n_regen else None,
            _capi.ptr(y0), _capi.ptr(refix_env), _capi.ptr(refix_noise), C.byref(n_refix))
        if rcode:
            raise _capi.DbsGymError(f"dbsgym_np_reset_draws failed ({rcode})")
        pr.end_state = state

        # ---- the arithmetic on them that touches only this reset's own arrays ----
        if n_fix.any():                                                   # utils.py:819-823 on w0
            rows = np.flatnonzero(n_fix)
            means = np.zeros(n)
            means[rows] = np.mean(w0[rows], axis=1)                       # (row-wise mean == np.mean of the 1-d vector, bit for bit)
            r_idx, c_idx = np.nonzero(bad)                                # row-major: the order the noise was drawn in
            w0[r_idx, c_idx] = np.abs(fix_noise[:r_idx.size] * 0.05) + means[r_idx]
        at = 0
        for r, k in refix_env[:n_refix.value]:                            # the same on the initial phases (env.py:598)
            row = y0[r]
            row[row <= 0.0] = np.abs(refix_noise[at:at + k] * 0.05) + np.mean(row)
            at += k
        pr.w0, pr.y0, pr.n_fix = w0, y0, n_fix
        return pr

    # ---- phase 2: apply a prepared reset to the batch --------------------------------------------------------
    def _commit(self, pr, changed_only):
        ids, n, N, rc = pr.ids, pr.n, self.N, pr.rc
        everyone, f_en, f_pl, f_rg, f_sp = pr.everyone, pr.f_en, pr.f_pl, pr.f_rg, pr.f_sp
        w0, y0, n_fix, inc = pr.w0, pr.y0, pr.n_fix, pr.inc
        self.stream.st = pr.end_state
        self.stream.push()
        self.reset_count[ids] = rc
        if self.drift:
            if f_pl.any():
                self.count[ids[f_pl]] += 1
            if f_rg.any():
                self.count[ids[f_rg]] = 0
            if pr.wl_rows.size:
                self.wl[ids[pr.wl_rows]] = pr.wl_vals
            self.elec_drift_episode[ids] += inc[:, 0]
            self.elec_encaps_episode[ids] += inc[:, 1]
            self.plasticity_episode[ids] += inc[:, 2]
            self.elec[ids] = pr.elec
            if f_en.any():
                self.encaps[ids[f_en]] += self.encaps_percent[ids[f_en]]          # added as an absolute amount
            if pr.walk_noise is not None:
                who = ids[f_rg]
                base = self.wl_orig[who]
                sigma = self.step_scale[who] * np.std(base, axis=1, ddof=1)     # env.py:21-57 generate_perturbations
                walk = np.empty((who.size, self.M + 1, N))
                walk[:, 0] = base
                for m in range(self.M):
                    walk[:, m + 1] = walk[:, m] + sigma[:, None] * pr.walk_noise[:, m]
                self.walk[who] = walk
        if f_sp.any():                                                    # env.py:544-552
            for r in np.flatnonzero(f_sp):
                i, row = int(ids[r]), self.table[int(pr.pick[r])]
                self.elec[i], self.rec[i] = row[0], row[1]
                self.spatial_var_episode[i] += self.spatial_var_freq[i]
                self.hosts[i].spatial_events.append([int(rc[r]), row])
        if everyone:
            self.w0, self.init_state = w0, y0
        else:
            self.w0[ids] = w0
            self.init_state[ids] = y0

        # ---- electrodes: one model per distinct (contacts, conduct_modifier), shared by the environments that have it ----
        keys = np.column_stack([self.elec[ids], self.rec[ids], self.encaps[ids]])
        el_changed = np.any(keys != self._last_key[ids], axis=1) if changed_only else np.ones(n, dtype=bool)
        rows = np.flatnonzero(el_changed)
        if rows.size:
            uniq, inv = np.unique(keys[rows], axis=0, return_inverse=True)
            inv = np.asarray(inv).reshape(-1)
            models = [cached_electrode(self.p0, self._cm(k[6]), [[int(v) for v in k[0:3]]], [[int(v) for v in k[3:6]]], 0) for k in uniq]
            stim = np.stack([m.stim_vector() for m in models])[inv]
            rec = np.stack([m.rec_vector() for m in models])[inv]
            for r, j in zip(rows, inv):
                self._electrodes[int(ids[r])] = models[j]
        else:
            stim = rec = np.empty((0, N))
        electrodes = [self._electrodes[int(i)] for i in ids]
        if self.B <= 256:            # small batches: keep the HostEnvState objects current for callers that hold on to one
            for i in ids:
                self.sync_to_host(int(i), self.hosts[int(i)])
        self._last_key[ids] = keys                 # (either way the caller now has the current vectors of these environments)
        if not changed_only:
            self._w0_sent[ids] = True
            return w0, stim, rec, y0, electrodes
        w0_changed = ~self._w0_sent[ids] | (n_fix > 0)
        if self.drift:
            w0_changed |= f_pl | f_rg
        self._w0_sent[ids] = True
        return w0, stim, rec, y0, electrodes, el_changed, w0_changed

    def _cm(self, v):
        """conduct_modifier as the per-environment path passes it (a Python float)."""
        return float(v)
