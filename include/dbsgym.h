/*
 * dbsgym.h -- C ABI of the B200-native DBS-Gym environment-step engine.
 *
 * The reference (NevVerVer/DBS-Gym) is pure Python: it has no FFI of its own.  The
 * boundary this library replaces is the body of the reference's gymnasium env
 * (paths relative to the reference checkout):
 *
 *   environment/env.py:415-454   SpatialKuramoto.step   -> dbsgym_step / dbsgym_step_host
 *   environment/env.py:605-612   transient part of reset -> dbsgym_transient
 *   environment/env.py:252-271   KuramotoJAX.dynamics / forward (diffrax Dopri5 +
 *                                PIDController(rtol=atol=1e-5), SaveAt(ts))      -> the step kernel
 *   environment/env.py:396-412   calc_naive_lfp / calc_distance_lfp             -> fused in the step kernel
 *   environment/env.py:447-452, :638-688, utils.py:21-27, :794-816
 *                                window slide, rewards R1/R2/R3                 -> the step kernel's fused tail
 *                                                                                  (R1/R3), the observation kernel (R2)
 *   aDBS_RL/evaluate_HF_DBS.py:122-135  calc_psd_for_simple_eval                -> dbsgym_trace_* / dbsgym_eval_bbpow
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success or a negative DBSGYM_E* code and
 *     never throws; dbsgym_last_error() gives the text of the last failure.
 *   - "host" pointers are ordinary CPU memory (pinned memory makes the copies faster);
 *     "device" pointers are CUDA device memory owned by the caller (e.g. a torch tensor's
 *     data_ptr()) and must stay alive until the stream work that uses them has finished.
 *   - per-oscillator parameter vectors always cross the boundary as float64, whatever the
 *     compute precision of the handle.
 *   - a handle is bound to one GPU and is not thread-safe; use one handle per GPU.
 *   - stream arguments are cudaStream_t values passed as void*: NULL is CUDA's legacy default
 *     stream as usual, DBSGYM_OWN_STREAM selects the handle's private non-blocking stream (the one
 *     the *_host calls use).  Calls issued on different streams are not ordered with respect to
 *     each other: synchronise when switching between them.
 */
#ifndef DBSGYM_H_
#define DBSGYM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBSGYM_ABI_VERSION 3
#define DBSGYM_OWN_STREAM ((void*)(intptr_t)-1)

typedef struct DbsGymHandle DbsGymHandle;

enum { DBSGYM_OK = 0, DBSGYM_EINVAL = -1, DBSGYM_ECUDA = -2, DBSGYM_ENOMEM = -3,
       DBSGYM_ESTATE = -4, DBSGYM_ESOLVER = -5 };

enum { DBSGYM_F32 = 0, DBSGYM_F64 = 1 };                 /* compute precision */
enum { DBSGYM_COUPLING_GRID = 0,                         /* alpha_ij = f(|dz|,|dx|,|dy|) on a regular grid */
       DBSGYM_COUPLING_DENSE = 1 };                      /* arbitrary symmetric alpha[N][N]            */
enum { DBSGYM_REWARD_BBPOW = 0,                          /* env.py:638-650  */
       DBSGYM_REWARD_TEMP_CONST = 1,                     /* env.py:653-666  */
       DBSGYM_REWARD_BBPOW_THRESH = 2 };                 /* env.py:669-688  */

typedef struct DbsGymConfig {
    uint32_t struct_bytes;      /* sizeof(DbsGymConfig), checked                                  */
    int32_t  device;            /* CUDA ordinal                                                   */
    int32_t  n_envs;            /* B: environments owned by this handle                           */
    int32_t  n_osc;             /* N: oscillators per environment (params_dict['num_oscillators']).
                                   GRID coupling: N <= 4096 one CTA per environment; 4096 < N <= 65536 (fp32, 8 x 8 x gz
                                   grids) one thread-block cluster of N/4096 CTAs per environment.  DENSE: N <= 8192. */
    int32_t  grid[3];           /* gx, gy, gz of utils.py:478-497 (row i = z*gx*gy + x*gy + y)    */
    int32_t  window;            /* W: observe_wind_idxs, env.py:296-297 (2340)                    */
    int32_t  precision;         /* DBSGYM_F32 | DBSGYM_F64                                        */
    int32_t  coupling;          /* DBSGYM_COUPLING_*                                              */
    int32_t  max_step_samples;  /* >= nI + nII - 1 of any step (env.py:444)                       */
    int32_t  max_steps;         /* diffrax max_steps per solve (4096)                             */
    double   K;                 /* params_dict['K']; the kernel applies K / N (env.py:264)        */
    double   rtol, atol, dt0;   /* env.py:249, :267                                               */
    double   safety, factor_min, factor_max;   /* diffrax PIDController defaults .9 / .2 / 10     */
    double   action_lo, action_hi;             /* dbs_action_bounds, env.py:389-393               */
    /* tuning / diagnostics (all 0 = library defaults; nothing in the library reads the environment) */
    int32_t  mw_mode;           /* multi-worker step kernel (8 x 8 x 8 grid, fp32): 0 = when the launch fills the
                                   GPU, 1 = never, 2 = always                                                  */
    int32_t  force_cluster;     /* > 1: integrate one environment with a cluster of that many CTAs even when
                                   n_osc <= 4096 (tests of the cluster path at small sizes)                    */
    int32_t  ctas_per_sm;       /* > 0: pad shared memory so that exactly this many CTAs share an SM           */
    uint32_t debug_flags;       /* DBSGYM_DBG_* (A/B switches used by tests and tuning runs)                   */
} DbsGymConfig;

enum { DBSGYM_DBG_NO_GEO1 = 1u,          /* skip the kernels specialised for the 8 x 8 x 8 grid               */
       DBSGYM_DBG_NO_SYM = 2u,           /* plain Toeplitz contraction instead of the parity sectors          */
       DBSGYM_DBG_NO_FSAL_REUSE = 4u,    /* evaluate the first stage of every segment (fp32 too)              */
       DBSGYM_DBG_NO_FUSED_OBS = 8u,     /* separate observation / reward kernel instead of the fused tail    */
       DBSGYM_DBG_NO_FAST_OBS = 16u,     /* generic observation kernel instead of the streamlined one         */
       DBSGYM_DBG_NO_WARP_KERNEL = 32u };/* spectral coupling: the 64-thread worker kernel instead of one warp per environment */

typedef struct DbsGymRewardSpec {
    uint32_t struct_bytes;
    int32_t  kind;              /* DBSGYM_REWARD_*                                                */
    int32_t  bin_lo, bin_hi;    /* rfft bins k with beta_a < k/(W*dt) < beta_b, inclusive range   */
    double   power_scale;       /* 1e4 (R1, R3)                                                   */
    double   action_cost;       /* 1e-2 (R1, R2), 1 (R3)                                          */
    double   threshold;         /* 20 (R3)                                                        */
    double   threshold_penalty; /* 5 (R3)                                                         */
    double   temp_scale;        /* 1e3 (R2)                                                       */
} DbsGymRewardSpec;

/* ---- lifetime ------------------------------------------------------------------------- */
int  dbsgym_abi_version(void);
/* compile-time switches of this build, for honest op counts in benchmarks: bit 0 = y-parity sector
 * contraction (DBSGYM_Y_PARITY), bit 1 = libm sincos instead of MUFU (DBSGYM_PRECISE_SINCOS) */
int  dbsgym_build_flags(void);
int  dbsgym_create(const DbsGymConfig* cfg, DbsGymHandle** out);
/* which step-kernel variant a launch over n_envs environments of this handle uses (for op counts in benchmarks):
 * 0 plain Toeplitz GRID, 1 DENSE, 2 parity-sector GRID_SYM (run-time extents), 3 GRID_SYM unrolled for the
 * 8 x 8 x 8 grid (one CTA per environment), 4 multi-worker (8 environments per CTA sharing precomputed sector
 * coefficients), 5 cluster mode (N > 4096), 6 GRID_SYM with gx = 8 fixed, 7 / 8 GRID_SYM with lines of 16 / 32 (gy = 16 / 32),
 * 9 spectral contraction (dbsgym_set_coupling_spectral; multi-worker hosting, 64 threads per environment),
 * 10 spectral contraction, one warp per environment with octant ownership (the default when the sector ranks fit a
 * compiled rank list), 11 DENSE operator in low-rank form (dbsgym_set_coupling_lowrank), 12 spectral contraction on the
 * 8 x 8 x 4 half grid (N = 256), one warp per environment with one octant point per lane, 13 sector form of the low-rank
 * operator with the eigenvectors in registers (1024 / 2048 / 4096 oscillators: one CTA per environment, 8192: a cluster of two; chosen by
 * dbsgym_set_coupling_lowrank_sectors when a compiled rank list covers the sectors' ranks); negative = error code */
int  dbsgym_step_variant(const DbsGymHandle* h, int32_t n_envs);
void dbsgym_destroy(DbsGymHandle* h);
/* text of the last error on this handle (h == NULL: last error of a failed create) */
const char* dbsgym_last_error(const DbsGymHandle* h);

/* ---- model set-up (host pointers, float64) ---------------------------------------------
 * GRID: table[(dz*gx + dx)*gy + dy] = alpha between two neurons whose grid offsets are
 *       (dx,dy,dz) -- i.e. cos(distance) or the wavelet kernel of env.py:219-229.
 * DENSE: alpha[i*N + j], must be symmetric.  Shared by every environment of the handle. */
int dbsgym_set_coupling_grid(DbsGymHandle* h, const double* table);
int dbsgym_set_coupling_dense(DbsGymHandle* h, const double* alpha);
/* Spectral form of the GRID operator (after dbsgym_set_coupling_grid; fp32 handles on the 8 x 8 x 8 grid): the
 * generalised mean-field identity
 *     sum_j alpha_ij sin(th_j - th_i) = sum_m lambda_m v_m[i] (cos th_i S_m - sin th_i C_m),  S_m = sum_j v_m[j] sin th_j, C_m likewise,
 * over the eigenpairs of alpha (env.py:219-223, :252-256) whose |lambda_m| exceeds the caller's truncation threshold.
 * alpha commutes with the three reflections of the grid, so the eigenvectors are given per parity sector
 * s = 4 [odd in y] + 2 [odd in z] + [odd in x]:  vecs[(s * 64 + a) * r_max + m] = unit eigenvector m of the 64 x 64
 * sector block at fundamental-octant point a = (zq * 4 + xq) * 4 + yq, vals[s * r_max + m] its eigenvalue (in the
 * unnormalised sector coordinates X_s[a] = sum_g chi_s(g) x[g a]); sector s uses its first ranks8[s] modes (0..9, at
 * most r_max).  The truncation error is the caller's responsibility (dbsgym_b200/geometry.py: spectral_factors returns
 * the spectral norm of what was dropped).  ranks8 == NULL switches back to the exact sector-block contraction.
 * The 8 x 8 x 4 half grid (n_osc = 256) is served the same way with 32-point sector blocks (a = (zq * 4 + xq) * 4 + yq,
 * zq < 2, vecs[(s * 32 + a) * r_max + m]); there the ranks must fit a compiled rank list ({5,4,4,2,4,4,2,1} or
 * {9,4,4,3,4,4,3,1}), otherwise DBSGYM_ESTATE and the handle keeps the exact contraction. */
int dbsgym_set_coupling_spectral(DbsGymHandle* h, const int32_t* ranks8, int32_t r_max,
                                 const double* vecs, const double* vals);

/* Low-rank form of the coupling operator (fp32 handles, DENSE or GRID: the neuron ordering does not matter, and GRID
 * handles bring the cluster mode for n_osc > 4096 with them): alpha ~ sum_m vals[m] v_m v_m^T with vecs[m * N + i] = v_m[i], the
 * eigenpairs of alpha (env.py:219-229; any neuron ordering, utils.py:483-497 shuffle=True) the caller keeps
 * (dbsgym_b200/geometry.py: lowrank_factors returns the spectral norm of what it dropped).  The step kernel then evaluates
 * the coupling sum of env.py:252-256 as sum_m v_m[i] (cos th_i S_m - sin th_i C_m) -- O(N rank) per evaluation instead of
 * O(N^2), rank * N floats of operator instead of N^2; dbsgym_set_coupling_dense need not be called at all.
 * rank <= 0 switches back to the full operator (matrix / grid table). */
int dbsgym_set_coupling_lowrank(DbsGymHandle* h, int32_t rank, const double* vecs, const double* vals);

/* Sector form of the low-rank operator for regular grids with even extents (fp32 GRID handles, right after
 * dbsgym_set_coupling_grid and before any dbsgym_set_env_params): the eigenpairs of the 8 parity-sector blocks of alpha
 * (dbsgym_b200/geometry.py: sector_block) over the N / 8 points of the fundamental octant, a = (zq * gx/2 + xq) * gy/2 + yq:
 * zvecs[m * N/8 + a], vals[m] (block eigenvalues), modes sorted by sector s = 4 [odd y] + 2 [odd z] + [odd x] with
 * soff9[s] = first mode of sector s, soff9[8] = number of modes, every sector padded to a multiple of 4 modes (zero rows).
 * 1/8 of the multiply-adds and of the eigenvector traffic of dbsgym_set_coupling_lowrank.  The library then stores the
 * oscillators in octant order internally; every entry point of this header keeps taking and returning the natural order
 * (dbsgym_get_state / dbsgym_set_state blobs are opaque and only valid for a handle in the same mode). */
int dbsgym_set_coupling_lowrank_sectors(DbsGymHandle* h, const int32_t* soff9, const double* zvecs, const double* vals);

/* Order of the oscillators (fp32 handles, right after dbsgym_create): order[d] = the caller's index of the oscillator that
 * sits at position d of the regular grid (row d = z * gx * gy + x * gy + y).  A grid whose neurons were shuffled
 * (utils.py:483-497 shuffle=True with n_neurons = the whole grid) is thereby stored in grid order inside the library and
 * runs the structured / spectral GRID kernels instead of the DENSE fallback; every entry point of this header keeps taking
 * and returning per-oscillator data in the caller's order.  order == NULL: identity. */
int dbsgym_set_oscillator_order(DbsGymHandle* h, const int32_t* order);

/* Per-environment vectors uploaded at reset (env.py:566-598): natural frequencies after
 * remove_negative_w0, stimulation conductance of the first contact (env.py:422-423), summed
 * recording conductance (env.py:410-411; NULL = 'naive' recording kernel) and the unwrapped
 * initial phases.  Arrays are [n][N]; env_ids == NULL means environments 0..n-1; any vector
 * pointer may be NULL to leave that quantity untouched. */
int dbsgym_set_env_params(DbsGymHandle* h, const int32_t* env_ids, int32_t n,
                          const double* w0, const double* stim_cond,
                          const double* rec_cond, const double* y0);
/* recording kernel: 0 = naive (observation = mean cos), 1 = conductance weighted */
int dbsgym_set_recording(DbsGymHandle* h, int32_t weighted);

/* The np.arange time grids of env.py:426-437 for step index 0..n_steps-1, computed on the
 * host with numpy itself (they depend on float64 rounding).
 * offs_I[k*max_I + j] = t_eval_step_I[j] - t_eval_step_I[0]; same for II. */
int dbsgym_set_schedule(DbsGymHandle* h, int32_t n_steps,
                        const int32_t* n_I, const int32_t* n_II,
                        const double* offs_I, int32_t max_I,
                        const double* offs_II, int32_t max_II);

/* reward definition; lin_functional (float64 [W], chronological order) is the vector g with
 * x_filt[-1] - mean(x_filt) == g . window for the filtfilt band-pass of utils.py:794-816
 * (needed for DBSGYM_REWARD_TEMP_CONST only, else NULL). */
int dbsgym_set_reward(DbsGymHandle* h, const DbsGymRewardSpec* spec, const double* lin_functional);

/* step counter and episode length (env.py:299-300, :450-451) of the listed environments */
int dbsgym_set_episode(DbsGymHandle* h, const int32_t* env_ids, int32_t n,
                       const int32_t* step_idx, const int32_t* episode_len);

/* ---- the hot path ----------------------------------------------------------------------
 * dbsgym_transient: env.py:605-612.  Integrates the listed environments from their current
 * phases over ts_offsets[0..n_ts-1] (one diffrax solve), fills the observation window with
 * the last W recorded LFP samples of ts[0..n_ts-2], leaves the phases at ts[n_ts-1], and
 * writes the reset observation obs[b*W .. ] (device float32, may be NULL).  env_ids is a host
 * pointer (NULL = all). Asynchronous on `stream`. */
int dbsgym_transient(DbsGymHandle* h, const int32_t* env_ids, int32_t n,
                     const double* ts_offsets, int32_t n_ts, float* obs_dev, void* stream);

/* dbsgym_step: env.py:415-454 for every environment of the handle.
 *   actions_dev  [B]     float32 in [-1, 1] (policy output, env.py:419)
 *   obs_dev      [B][W]  float32 observation window after the step
 *   reward_dev   [B]     float32 reward  (float64 copy: dbsgym_get_rewards)
 *   done_dev     [B]     uint8   current_step >= total_episode_counts
 * Asynchronous on `stream`; any output pointer may be NULL. */
int dbsgym_step(DbsGymHandle* h, const float* actions_dev, float* obs_dev,
                float* reward_dev, uint8_t* done_dev, void* stream);

/* Same with HOST buffers: copies actions in, runs the step, copies obs / reward / done out
 * and synchronises.  This is the call a gym VecEnv makes. */
int dbsgym_step_host(DbsGymHandle* h, const float* actions, float* obs,
                     float* reward, uint8_t* done);
/* Delta-transfer variant for hosts that mirror the observation window themselves: instead of the
 * [B][W] window only the step's NEW samples cross PCIe -- samples[b*max_step_samples + i], float32,
 * i < n_samples[b], exactly the values appended to the window (env.py:447) -- so the host can
 * slide its own copy (obs_{t+1} = concat(obs_t[n:], samples)).  reward / done as above. */
int dbsgym_step_host_samples(DbsGymHandle* h, const float* actions, float* samples,
                             int32_t* n_samples, float* reward, uint8_t* done);
/* Host mirror of the observation windows, written by the GPU itself: an append-only sample log per environment.
 * dbsgym_host_mirror() returns the CURRENT one of two pinned, device-mapped float32 arrays [B][*row_floats]
 * (row_floats = 2 * L, L = W + 256) owned by the library.  The step kernel stores every new window sample twice
 * through PCIe, at column c and at c + L (c = the environment's write position, advanced modulo L), so the
 * chronological window of environment b is always the CONTIGUOUS slice mirror[b][pos .. pos + W - 1] and no data
 * ever has to be moved or re-copied by the CPU.
 * LIFETIME of a window handed out this way (what lets a VecEnv return it without copying, like the arrays
 * DummyVecEnv returns): the next steps only append BEHIND it -- its first sample is overwritten after 256 further
 * samples, i.e. it stays intact for at least 13 more steps (<= 19 samples each) -- and a reset
 * (dbsgym_transient) switches to the other buffer and rewrites all windows there, so a window of the old buffer
 * stays intact until the reset after the next one.  Call dbsgym_host_mirror() again after every dbsgym_transient
 * to get the buffer that is current.
 * dbsgym_step_host_mirror() = one step; only actions (H2D) and reward / done / the new samples (D2H,
 * zero-copy) cross PCIe.  *pos receives the column of the oldest sample and *n_new the number of new
 * samples; if the environments are out of lockstep (different log positions, possible after a reset of a subset)
 * *n_new = -1 and the caller must read the windows back instead (dbsgym_get_obs_host). */
int dbsgym_host_mirror(DbsGymHandle* h, float** mirror, int32_t* row_floats);
int dbsgym_step_host_mirror(DbsGymHandle* h, const float* actions, int32_t* pos, int32_t* n_new,
                            float* reward, uint8_t* done);
/* reset observation (window as float32) of all environments to a host buffer [B][W] */
/* the same step in two halves: _begin copies the actions and launches the step kernel (returns immediately), _end
 * waits for it and returns what dbsgym_step_host_mirror returns -- host work can overlap the GPU in between */
int dbsgym_step_host_mirror_begin(DbsGymHandle* h, const float* actions);
int dbsgym_step_host_mirror_end(DbsGymHandle* h, int32_t* pos, int32_t* n_new, float* reward, uint8_t* done);
int dbsgym_get_obs_host(DbsGymHandle* h, float* obs);

/* ---- introspection (host buffers, synchronising) -------------------------------------- */
/* LFP samples of the last step: lfp_true = theta_mean (env.py:444), lfp_rec = theta_records
 * (env.py:445), each [B][max_step_samples] float64; n_samples [B]. Any may be NULL. */
int dbsgym_get_lfp(DbsGymHandle* h, double* lfp_true, double* lfp_rec, int32_t* n_samples);
int dbsgym_get_rewards(DbsGymHandle* h, double* reward, double* u);
/* unwrapped phases, float64 [n][N] */
int dbsgym_get_phases(DbsGymHandle* h, const int32_t* env_ids, int32_t n, double* y);
/* Snapshot / restore of the whole handle (the reference never checkpoints the environment; SB3's
 * CheckpointCallback saves only the agent, aDBS_RL/train_aDBS_RL.py:145-150).  The blob holds everything later steps
 * depend on -- phases + winding counts, w0 / conductances, observation rings + heads, running rfft bins, step counters,
 * done flags, the carried FSAL row, the last step's samples, reward, u and the solver counters -- so that
 * set_state(get_state()) followed by the same actions reproduces the same outputs bit for bit.  Coupling, schedule and
 * reward definition are configuration, not state: the restoring handle must have been set up the same way
 * (the header of the blob is checked against n_envs, n_osc, window, precision). */
int dbsgym_state_bytes(const DbsGymHandle* h, uint64_t* bytes);
int dbsgym_get_state(DbsGymHandle* h, void* blob, uint64_t bytes);
int dbsgym_set_state(DbsGymHandle* h, const void* blob, uint64_t bytes);
/* kernels launched through this handle since create (or the last call with reset != 0) */
int dbsgym_launch_count(DbsGymHandle* h, uint64_t* launches, int32_t reset);
/* observation window in chronological order, float64 [n][W] */
int dbsgym_get_window(DbsGymHandle* h, const int32_t* env_ids, int32_t n, double* window);
int dbsgym_set_window(DbsGymHandle* h, const int32_t* env_ids, int32_t n, const double* window);
int dbsgym_get_episode(DbsGymHandle* h, int32_t* step_idx, uint8_t* done);
/* totals since create (or the last call with reset != 0): accepted RK steps, rejected RK
 * steps, RHS evaluations, summed over environments; and the device status word (0 = ok). */
int dbsgym_counters(DbsGymHandle* h, uint64_t* accepted, uint64_t* rejected,
                    uint64_t* rhs_evals, int32_t* status, int32_t reset);
/* of the rhs_evals counted above (the evaluations the reference performs), how many the fp32 kernel did NOT execute
 * because the first stage of a segment was taken from the last stage of the previous accepted sub-step (same state,
 * only the pulse term differs: k1 = k7 + (amp_new - amp_old) * stim); reset together with dbsgym_counters */
int dbsgym_rhs_reused(DbsGymHandle* h, uint64_t* reused);

/* elapsed device time (ms) of the kernels of the most recent dbsgym_step*, measured with
 * CUDA events on the launching stream: [0] step kernel, [1] observation kernel */
int dbsgym_last_step_ms(DbsGymHandle* h, float* ms2);
/* enable / disable the per-kernel event timing above (off by default) */
int dbsgym_set_timing(DbsGymHandle* h, int32_t enabled);

/* ---- evaluation metric on the device (aDBS_RL/evaluate_HF_DBS.py:122-135) -----------------------
 * dbsgym_trace_begin: from now on every step appends its TRUE-LFP samples (theta_mean, env.py:441) to a per-
 * environment device trace of `capacity` float64 samples (lengths reset to 0); needs the beta-power reward
 * path (the fused observation tail).  dbsgym_trace_end stops recording (the trace stays readable).
 * dbsgym_trace_get copies trace [n_envs][capacity] and the lengths to the host (either may be NULL). */
int dbsgym_trace_begin(DbsGymHandle* h, int32_t capacity);
int dbsgym_trace_end(DbsGymHandle* h);
int dbsgym_trace_get(DbsGymHandle* h, double* trace, int32_t* len);

typedef struct DbsGymEvalSpec {
    uint32_t struct_bytes;
    int32_t  padlen;            /* filtfilt odd extension: 3 * max(len(a), len(b)) = 15                       */
    double   b[5], a[5];        /* butter(2, [12, 30] Hz, 'band') at fs = 1 / psd_dt (utils.py:794-816)       */
    double   zi[4];             /* scipy.signal.lfilter_zi(b, a)                                              */
    int32_t  k_lo, n_k;         /* rfft bins k_lo .. k_lo + n_k - 1 carry non-zero weight                     */
} DbsGymEvalSpec;
/* bbpow[e] = sum_k weights[k] * 2 |rfft(filtfilt(b, a, trace_e))[k_lo + k] / n|^2 for every environment; all
 * traces must have the same length n >= padlen + 2.  `weights` (float64 [n_k], host) folds the 12-tap spectrum
 * smoothing and the band sum of calc_psd_for_simple_eval (evaluate_HF_DBS.py:130-134). */
int dbsgym_eval_bbpow(DbsGymHandle* h, const DbsGymEvalSpec* spec, const double* weights, double* bbpow);

/* ---- host side of a batched reset: numpy's legacy global random stream, replayed natively ---------------------------
 * The reference draws everything reset() needs from the process-global np.random (env.py:291, :483-598; SURVEY.md
 * Appendix C).  To hand a batch of environments the numbers a sequential DummyVecEnv of reference environments would get,
 * the caller passes np.random.get_state() in, lets dbsgym_np_reset_draws consume the stream for all environments in
 * index order, and writes the state back with np.random.set_state().  No GPU involved. */
typedef struct DbsGymNpState {
    uint32_t key[624];          /* MT19937 state words                                                         */
    int32_t  pos;               /* 0..624                                                                      */
    int32_t  has_gauss;         /* the polar method's cached second deviate                                    */
    double   gauss;
} DbsGymNpState;
/* n draws of np.random.normal(loc, scale) (randn: loc 0, scale 1) */
int dbsgym_np_gauss(DbsGymNpState* st, int64_t n, double loc, double scale, double* out);
/* n draws of np.random.choice(pop_size) (== RandomState.randint(0, pop_size)) */
int dbsgym_np_choice(DbsGymNpState* st, int64_t n, uint32_t pop_size, int32_t* out);

enum { DBSGYM_RESET_ELECTRODE_MOVE = 1, DBSGYM_RESET_ENCAPSULATION = 2, DBSGYM_RESET_PLASTICITY = 4,
       DBSGYM_RESET_WALK_REGEN = 8, DBSGYM_RESET_SPATIAL = 16 };
typedef struct DbsGymResetPlan {
    uint32_t struct_bytes;
    int32_t  n_envs, n_osc;
    int32_t  walk_len;              /* M = 2 * reset_plasticity_episode vectors per regenerated plasticity walk (env.py:539) */
    int32_t  coord_lo, coord_hi;    /* an electrode move is redrawn until 1 <= coordinate <= min(grid_size) - 2 (env.py:487-496) */
    int32_t  table_len;             /* rows of the stim / rec / locus table a spatial re-draw picks from (env.py:546) */
    int32_t  random_freq_update;    /* params_dict['random_freq_update'] (env.py:457-464)                      */
    int32_t  refix_cap_rows, refix_cap_noise;   /* capacity of refix_env (pairs) / refix_noise                 */
    int32_t  reserved[2];
    double   init_mean, init_sd;    /* env.py:594-595                                                          */
} DbsGymResetPlan;
/* The draws of ONE reset of n_envs environments, in index order, exactly as env.py:483-598 makes them:
 *   flags[e]        DBSGYM_RESET_* events due at this reset
 *   freq[e][3]      electrode_drift_freq, encapsulation_drift_freq, plasticity_drift_freq
 *   elec_coords[e][3]  first stimulation contact, moved in place when DBSGYM_RESET_ELECTRODE_MOVE is set
 *   next_inc[e][3]  out: what to add to elec_drift_episode / elec_encaps_episode / plasticity_episode
 *   spatial_pick[e] out: row of the table (-1 = no re-draw)
 *   n_fix[e]        number of non-positive natural frequencies remove_negative_w0 replaces (utils.py:819-823);
 *                   fix_noise receives that many standard normals per environment, concatenated
 *   walk_noise      [number of environments with DBSGYM_RESET_WALK_REGEN][walk_len][n_osc] standard normals
 *   init_state      [n_envs][n_osc] = init_mean + init_sd * gauss
 *   refix_env / refix_noise / n_refix   environments whose init_state came out non-positive somewhere (pairs env, count)
 *                   and the standard normals remove_negative_w0(init_state) draws for them (env.py:598)
 * Returns DBSGYM_ESTATE when the refix buffers are too small (state untouched: copy it before the call to retry). */
int dbsgym_np_reset_draws(DbsGymNpState* st, const DbsGymResetPlan* plan, const uint8_t* flags, const int32_t* freq,
                          int32_t* elec_coords, int32_t* next_inc, int32_t* spatial_pick, const int32_t* n_fix,
                          double* fix_noise, double* walk_noise, double* init_state, int32_t* refix_env,
                          double* refix_noise, int32_t* n_refix);

/* FP32-FMA throughput micro-benchmark used for the roofline denominator: runs a dependent-
 * chain FFMA kernel on `device` for about `ms_target` ms; returns TFLOP/s in *tflops. */
int dbsgym_measure_fp32_peak(int32_t device, double ms_target, double* tflops);
/* special-function unit: dependent MUFU.SIN / MUFU.COS chains, result in 1e12 operations per second */
int dbsgym_measure_mufu_peak(int32_t device, double ms_target, double* tops);
/* the variants separately: packed == 0 scalar FFMA chains, packed == 1 FFMA2 (f32x2) chains, 2 = MUFU (as above) */
int dbsgym_measure_fp32_peak_mode(int32_t device, double ms_target, int32_t packed, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* DBSGYM_H_ */
