// C-ABI of the DBS-Gym step engine (see include/dbsgym.h for the contract).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/dbsgym.h"
#include "eval_kernel.cuh"
#include "obs_kernel.cuh"
#include "step_launch.h"

using namespace dbsgym;

namespace {

thread_local char g_create_err[512] = "";

// Host mirror log: samples that stay valid behind a returned window.  A window handed to the caller (W columns ending at the
// write position) is overwritten only after kMirrorGuard further samples have been appended: 13 steps of <= 19 samples.
constexpr int kMirrorGuard = 256;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct DbsGymHandle {
    DbsGymConfig cfg;
    int N = 0, Np = 0, B = 0, W = 0, smax = 0;
    bool f64 = false;
    size_t rb = 4;                       // bytes per real
    int nthreads = 0, tab = 0;
    cudaStream_t stream = nullptr;
    // model
    void* table = nullptr; void* alpha = nullptr;
    bool have_coupling = false;
    int weighted_rec = 0;
    // per-env state
    void *w0 = nullptr, *stim = nullptr, *rec = nullptr, *phase = nullptr, *ring = nullptr;
    int32_t *wind = nullptr, *head = nullptr, *n_samples = nullptr, *step_idx = nullptr, *episode_len = nullptr;
    double *lfp_true = nullptr, *lfp_rec = nullptr, *u = nullptr, *reward = nullptr;
    uint8_t* done = nullptr;
    // schedule
    int32_t *sched_nI = nullptr, *sched_nII = nullptr;
    double *sched_offI = nullptr, *sched_offII = nullptr;
    int maxI = 0, maxII = 0, n_sched = 0;
    // transient
    double* ts_dev = nullptr; int ts_cap = 0;
    int32_t* ids_dev = nullptr; int ids_cap = 0;
    // reward
    DbsGymRewardSpec rspec;
    bool have_reward = false;
    double *lin_g = nullptr, *tw_seed = nullptr;
    void* tw_inner = nullptr;
    double *spec = nullptr, *tw_full = nullptr;   // fused observation tail: running rfft bins [B][kTailBins][2], twiddles [W][nbins][2]
    bool fuse_tail = false;
    int obs_iters = 0;
    int nbins = 0;
    // bookkeeping
    unsigned long long* counters = nullptr;
    int32_t* status = nullptr;
    // host-API staging
    float *st_actions = nullptr, *st_obs = nullptr, *st_reward = nullptr, *st_samples = nullptr;
    // pinned + mapped window mirrors (dbsgym_host_mirror): two buffers [B][2 * mir_len], used alternately -- every
    // reset (dbsgym_transient) moves on to the other one, so windows handed out before it stay untouched
    float* mirror_host[2] = {nullptr, nullptr};
    float* mirror_dev[2] = {nullptr, nullptr};
    int mirror_cur = 0, mir_len = 0;
    int32_t* mpos = nullptr;             // [B] write column of every environment's mirror log
    int32_t* pin_ints = nullptr;         // pinned landing buffer: n_samples[B], head[B]
    // zero-copy control block of the host-mirror step (pinned + mapped): actions[B] f32 | reward[B] f32 |
    // n_samples[B] i32 | head[B] i32 | done[B] u8 -- read / written by the step kernel itself through PCIe
    unsigned char* ctl_host = nullptr; unsigned char* ctl_dev = nullptr;
    // evaluation trace (dbsgym_trace_begin): TRUE LFP of every step, [B][trace_cap] float64
    double* trace = nullptr; int32_t* trace_len = nullptr; int trace_cap = 0; bool trace_on = false;
    // FSAL carried across segments / launches (fp32): last stage derivative per environment + valid flags
    void* k_fsal = nullptr; int32_t* fsal_valid = nullptr; bool fsal_on = false;
    // grow-only staging pair for host -> device row uploads (pinned host side: true DMA, no cudaMalloc / cudaFree per call)
    void* stage_dev = nullptr; void* stage_host = nullptr; size_t stage_bytes = 0;
    bool mirror_on = false;              // obs kernel writes the mirror (set while a mirror step / reset runs)
    bool mirror_pending = false;         // between dbsgym_step_host_mirror_begin and _end
    uint8_t* st_done = nullptr;
    // timing
    int cluster = 1;                     // CTAs per environment (thread-block cluster; > 1 when N > 4096)
    void* cl_operand = nullptr; double* cl_scratch = nullptr;
    int ctas_per_sm = 0;                 // 0 = whatever fits
    int num_sms = 0;
    int mw_mode = -1;                    // multi-worker step kernel: -1 auto (full-occupancy batches), 0 never, 1 always
    bool no_geo1 = false, no_fast_obs = false, no_fused_obs = false;     // DbsGymConfig.debug_flags
    // spectral form of the coupling operator (dbsgym_set_coupling_spectral): eigenvector entries per worker thread and
    // eigenvalues x K / (8 N) per reduction row, and which compiled (RE, RO) pair serves them (0 = off)
    float* spec_v = nullptr; float* spec_lam = nullptr; int spec_re = 0, spec_ro = 0;
    // the same operator laid out for the one-warp-per-environment kernel (warp_kernel.cuh): [32 lanes][modes] float2 and
    // [modes] eigenvalues, for the compiled rank list warp_set (-1: the ranks fit none, or DBSGYM_DBG_NO_WARP_KERNEL)
    float* wspec_v = nullptr; float* wspec_lam = nullptr; int warp_set = -1; bool no_warp = false;
    // sector form of the low-rank operator on 1024 ... 4096 oscillators: tables of oct_kernel.cuh ([mode][N / 8] eigenvector
    // entries, eigenvalues x K / (8 N)) for the compiled rank list oct_set (-1: the block kernel of step_kernel.cuh runs)
    float* oct_v = nullptr; float* oct_lam = nullptr; int oct_set = -1;
    int half_set = -1;                   // 8 x 8 x 4 half grid: compiled rank list of warp1_kernel.cuh in use (-1: exact contraction)
    // low-rank form of a DENSE operator (dbsgym_set_coupling_lowrank): eigenvectors [lr_rank][Np], eigenvalues [lr_rank]
    float* lr_v = nullptr; float* lr_lam = nullptr; int lr_rank = 0;
    // sector form of the low-rank operator: oscillators stored in octant order (perm[d] = natural index of device position d)
    int32_t* lr_soff = nullptr; bool lr_sectors = false; bool params_set = false;
    // order of the oscillators on the device: perm[d] = the CALLER's index of the oscillator at device position d (empty =
    // identity).  Composed of the caller's own order (dbsgym_set_oscillator_order: a shuffled regular grid is stored in grid
    // order) and the octant order of the sector form; applied wherever per-oscillator data cross the ABI.
    std::vector<int32_t> perm, user_order, internal_order;
    void recompute_order() {
        perm.clear();
        if (user_order.empty() && internal_order.empty()) return;
        perm.resize((size_t)N);
        for (int d = 0; d < N; ++d) {
            const int nat = internal_order.empty() ? d : internal_order[d];
            perm[d] = user_order.empty() ? nat : user_order[nat];
        }
    }
    unsigned long long n_launches = 0;   // kernels launched by this handle (dbsgym_launch_count)
    // ordering between the private stream and caller streams (dbsgym_step / dbsgym_transient on a user stream)
    cudaEvent_t ev_own = nullptr, ev_user = nullptr; bool user_pending = false;
    bool grid_sym = false;               // GRID coupling: use the reflection-symmetry reduced contraction
    bool timing = false;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    float last_ms[2] = {0.f, 0.f};
    char err[512] = "";
};

namespace {

int fail(DbsGymHandle* h, int code, const char* fmt, ...) {
    char* dst = h ? h->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(h, call)                                                                            \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(h, DBSGYM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

template <typename T>
cudaError_t dalloc(T** p, size_t count) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
    return e;
}

__global__ void scatter_rows_kernel(unsigned char* dst, const unsigned char* src, const int32_t* ids,
                                    int n, size_t row_bytes) {
    const int r = blockIdx.x;
    if (r >= n) return;
    const int d = ids ? ids[r] : r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (size_t)r * row_bytes);
    uint32_t* o = reinterpret_cast<uint32_t*>(dst + (size_t)d * row_bytes);
    for (size_t i = threadIdx.x; i < row_bytes / 4; i += blockDim.x) o[i] = s[i];
}

__global__ void gather_rows_kernel(unsigned char* dst, const unsigned char* src, const int32_t* ids,
                                   int n, size_t row_bytes) {
    const int r = blockIdx.x;
    if (r >= n) return;
    const int d = ids ? ids[r] : r;
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src + (size_t)d * row_bytes);
    uint32_t* o = reinterpret_cast<uint32_t*>(dst + (size_t)r * row_bytes);
    for (size_t i = threadIdx.x; i < row_bytes / 4; i += blockDim.x) o[i] = s[i];
}

// dependent FFMA chains; 2*kChains*iters flops per thread
constexpr int kChains = 8;
__global__ void fma_peak_kernel(float* out, int iters, float a, float b) {
    float v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = (float)(threadIdx.x + c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) v[c] = fmaf(v[c], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += v[c];
    if (s == 123.456f) out[0] = s;
}

__global__ void fma2_peak_kernel(float* out, int iters, float a, float b) {
    float2 v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = make_float2((float)(threadIdx.x + c), (float)c);
    const float2 bb = make_float2(b, b * 0.5f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) v[c] = __ffma2_rn(make_float2(a, a), v[c], bb);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += v[c].x + v[c].y;
    if (s == 123.456f) out[0] = s;
}

// dependent MUFU.SIN / MUFU.COS chains: 2 * kChains * iters special-function operations per thread
__global__ void mufu_peak_kernel(float* out, int iters, float a) {
    float v[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) v[c] = 0.001f * (float)(threadIdx.x + c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) v[c] = __cosf(__sinf(v[c]) + a);
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += v[c];
    if (s == 123.456f) out[0] = s;
}

int upload_ids(DbsGymHandle* h, const int32_t* env_ids, int n, const int32_t** dev) {
    *dev = nullptr;
    if (!env_ids) return DBSGYM_OK;
    for (int i = 0; i < n; ++i)
        if (env_ids[i] < 0 || env_ids[i] >= h->B) return fail(h, DBSGYM_EINVAL, "env id %d out of range", env_ids[i]);
    if (n > h->ids_cap) {
        if (h->ids_dev) cudaFree(h->ids_dev);
        h->ids_dev = nullptr;
        CU(h, cudaMalloc(&h->ids_dev, sizeof(int32_t) * (size_t)n));
        h->ids_cap = n;
    }
    CU(h, cudaMemcpyAsync(h->ids_dev, env_ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));     // env_ids is caller memory
    *dev = h->ids_dev;
    return DBSGYM_OK;
}

int ensure_stage(DbsGymHandle* h, size_t bytes) {
    if (bytes <= h->stage_bytes) return DBSGYM_OK;
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->stage_dev) cudaFree(h->stage_dev);
    if (h->stage_host) cudaFreeHost(h->stage_host);
    h->stage_dev = h->stage_host = nullptr; h->stage_bytes = 0;
    const size_t cap = bytes + bytes / 4;
    CU(h, cudaMalloc(&h->stage_dev, cap));
    CU(h, cudaMallocHost(&h->stage_host, cap));
    h->stage_bytes = cap;
    return DBSGYM_OK;
}

// scatter n rows of `row_bytes` from host memory (h->stage_host itself, or any other buffer) into a [B][row] device array
int scatter_to_device(DbsGymHandle* h, void* dst, const void* host_rows, const int32_t* ids_dev, int n,
                      size_t row_bytes) {
    const size_t bytes = row_bytes * (size_t)n;
    if (host_rows != h->stage_host || bytes > h->stage_bytes) {
        int rc = ensure_stage(h, bytes);
        if (rc) return rc;
        memcpy(h->stage_host, host_rows, bytes);
    }
    cudaError_t e = cudaMemcpyAsync(h->stage_dev, h->stage_host, bytes, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        scatter_rows_kernel<<<n, 128, 0, h->stream>>>(static_cast<unsigned char*>(dst),
                                                      static_cast<const unsigned char*>(h->stage_dev), ids_dev, n, row_bytes);
        ++h->n_launches;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);      // the staging pair is reused by the next call
    if (e != cudaSuccess) return fail(h, DBSGYM_ECUDA, "scatter failed: %s", cudaGetErrorString(e));
    return DBSGYM_OK;
}

int gather_from_device(DbsGymHandle* h, void* host_rows, const void* src, const int32_t* ids_dev, int n,
                       size_t row_bytes) {
    void* tmp = nullptr;
    CU(h, cudaMalloc(&tmp, row_bytes * (size_t)n));
    gather_rows_kernel<<<n, 128, 0, h->stream>>>(static_cast<unsigned char*>(tmp),
                                                 static_cast<const unsigned char*>(src), ids_dev, n, row_bytes);
    ++h->n_launches;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_rows, tmp, row_bytes * (size_t)n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(h, DBSGYM_ECUDA, "gather failed: %s", cudaGetErrorString(e));
    return DBSGYM_OK;
}

template <typename real>
void to_real_rows(const double* src, int n, int N, int Np, std::vector<unsigned char>& out) {
    out.assign((size_t)n * Np * sizeof(real), 0);
    real* o = reinterpret_cast<real*>(out.data());
    for (int r = 0; r < n; ++r)
        for (int i = 0; i < N; ++i) o[(size_t)r * Np + i] = (real)src[(size_t)r * N + i];
}
// the same conversion straight into the pinned staging buffer
template <typename real>
void to_real_rows_stage(const double* src, int n, int N, int Np, void* stage) {
    real* o = reinterpret_cast<real*>(stage);
    for (int r = 0; r < n; ++r) {
        for (int i = 0; i < N; ++i) o[(size_t)r * Np + i] = (real)src[(size_t)r * N + i];
        for (int i = N; i < Np; ++i) o[(size_t)r * Np + i] = real(0);
    }
}

void fill_params(DbsGymHandle* h, StepParams& p) {
    const DbsGymConfig& c = h->cfg;
    p.N = h->N; p.Np = h->Np; p.B = h->B;
    p.GX = c.grid[0]; p.GZ = c.grid[2]; p.GY = c.grid[1];
    p.weighted_rec = h->weighted_rec;
    p.max_steps = c.max_steps;
    p.k_over_n = c.K / (double)h->N;
    p.rtol = c.rtol; p.atol = c.atol; p.dt0 = c.dt0;
    p.safety = c.safety; p.fmin = c.factor_min; p.fmax = c.factor_max;
    p.tol_end = 1e-10;
    p.act_lo = c.action_lo; p.act_hi = c.action_hi;
    p.table = h->table; p.alpha = h->alpha;
    p.phase = h->phase; p.wind = h->wind;
    p.w0 = h->w0; p.stim = h->stim; p.rec = h->rec;
    p.step_idx = h->step_idx;
    p.sched_nI = h->sched_nI; p.sched_nII = h->sched_nII;
    p.sched_offI = h->sched_offI; p.sched_offII = h->sched_offII;
    p.maxI = h->maxI; p.maxII = h->maxII; p.n_sched = h->n_sched;
    p.lfp_true = h->lfp_true; p.lfp_rec = h->lfp_rec; p.n_samples = h->n_samples;
    p.smax = h->smax; p.u_out = h->u;
    p.ring = h->ring; p.W = h->W; p.head = h->head;
    p.counters = h->counters; p.status = h->status;
    p.ts = nullptr; p.n_ts = 0; p.actions = nullptr; p.env_ids = nullptr; p.n_launch = 0; p.mode = MODE_STEP;
    p.cluster = h->cluster; p.cl_operand = h->cl_operand; p.cl_scratch = h->cl_scratch;
    p.tail_on = 0; p.tail_kind = h->rspec.kind; p.tail_nbins = h->nbins;
    p.spec = h->spec; p.tw_full = h->tw_full;
    p.samples_f = nullptr; p.mirror = nullptr; p.reward_f = nullptr; p.reward = h->reward;
    p.mir_len = h->mir_len; p.mpos = h->mpos;
    p.done_out = nullptr; p.done_dev = h->done;
    p.step_idx_rw = h->step_idx; p.episode_len = h->episode_len;
    p.nsamp_out = nullptr; p.head_out = nullptr;
    p.trace = nullptr; p.trace_len = nullptr; p.trace_cap = 0;
    p.fsal_on = h->fsal_on ? 1 : 0; p.k_fsal = h->k_fsal; p.fsal_valid = h->fsal_valid;
    const bool warp = (h->spec_re > 0 && h->warp_set >= 0) || h->half_set >= 0;
    p.spec_v = warp ? h->wspec_v : h->spec_v; p.spec_lam = warp ? h->wspec_lam : h->spec_lam;
    if (h->oct_set >= 0) { p.spec_v = h->oct_v; p.spec_lam = h->oct_lam; }
    p.lr_v = h->lr_v; p.lr_lam = h->lr_lam; p.lr_rank = h->lr_rank;
    p.lr_sectors = h->lr_sectors ? 1 : 0; p.lr_soff = h->lr_soff;
    p.power_scale = h->rspec.power_scale; p.action_cost = h->rspec.action_cost;
    p.threshold = h->rspec.threshold; p.threshold_penalty = h->rspec.threshold_penalty;
}

size_t plain_smem(DbsGymHandle* h, bool dense) {
    size_t smem = step_smem_bytes(h->Np, dense ? 0 : h->tab, h->nthreads, h->rb);
    if (h->ctas_per_sm > 0) {
        // occupancy knob: pad the dynamic shared memory so that exactly ctas_per_sm CTAs fit on an SM
        // (227 KB usable, 1 KB reserved per CTA) -- used to balance the waves of a launch
        const size_t want = (size_t)(227 * 1024) / (size_t)h->ctas_per_sm - 1024;
        if (want > smem) smem = want & ~(size_t)15;
    }
    return smem;
}

// which float32 GRID_SYM kernel family serves this handle: 1 = 8 x 8 x 8 (unrolled), 2 = gx 8, 3 / 4 = lines of 16 / 32, 0 = generic
int sym_geo(const DbsGymHandle* h, const StepParams& p) {
    if (p.GY == 2 * kRows) return 3;
    if (p.GY == 4 * kRows) return 4;
    if (h->nthreads == 64 && p.GZ == 8 && p.GX == 8 && !h->no_geo1) return 1;
    if (p.GX == 8 && h->nthreads <= 512) return 2;
    return 0;
}

// the step kernels live in their own translation units (step_*.cu, declared in step_launch.h)
cudaError_t launch_step(DbsGymHandle* h, const StepParams& p, cudaStream_t s) {
    ++h->n_launches;
    const int t = h->nthreads;
    if (h->oct_set >= 0) return launch_f32_oct(h->oct_set, h->num_sms, p, s);
    if (h->lr_rank > 0 && !h->f64) {                      // the operator in low-rank form (GRID and DENSE handles alike)
        if (h->cluster > 1) return launch_f32_lowrank_cluster(t, h->cluster, h->lr_rank, p, s);
        return launch_f32_lowrank(t, step_smem_bytes(h->Np, 2 * h->lr_rank, t, 4), p, s);
    }
    if (h->cluster > 1) return launch_f32_cluster(p.GY == 2 * kRows ? 3 : p.GY == 4 * kRows ? 4 : 2, t, h->cluster, p, s);
    const bool dense = h->cfg.coupling == DBSGYM_COUPLING_DENSE;
    const size_t smem = plain_smem(h, dense);
    if (h->f64) {
        if (dense) return launch_f64_dense(t, smem, p, s);
        return h->grid_sym ? launch_f64_sym(t, smem, p, s) : launch_f64_grid(t, smem, p, s);
    }
    if (dense) return launch_f32_dense(t, smem, p, s);
    if (!h->grid_sym) return launch_f32_grid(t, smem, p, s);
    if (h->half_set >= 0) return launch_f32_warp1(h->half_set, h->num_sms, p, s);
    if (h->spec_re > 0 && h->warp_set >= 0) return launch_f32_warp(h->warp_set, h->num_sms, p, s);
    if (h->spec_re > 0) return launch_f32_spectral(h->spec_ro, h->num_sms, p, s);
    const int geo = sym_geo(h, p);
    if (geo == 1) {
        // enough environments to fill every SM with kMwEnvs of them: share the sector-coefficient table
        const bool mw = kYParity && (h->mw_mode == 1 || (h->mw_mode < 0 && p.n_launch >= kMwEnvs * h->num_sms));
        if (mw) {
            int ctas = (p.n_launch + kMwEnvs - 1) / kMwEnvs;
            return launch_f32_mw(ctas > h->num_sms ? h->num_sms : ctas, p, s);
        }
    }
    return launch_f32_sym(geo, t, smem, p, s);
}

cudaError_t launch_obs(DbsGymHandle* h, float* obs, float* reward_f, uint8_t* done_out, int append,
                       const int32_t* ids_dev, int n, cudaStream_t s, float* samples_f = nullptr) {
    ObsParams o;
    o.samples_f = samples_f;
    o.mirror = h->mirror_on ? h->mirror_dev[h->mirror_cur] : nullptr;
    o.mir_len = h->mir_len; o.mpos = h->mpos;
    o.B = h->B; o.W = h->W; o.smax = h->smax;
    o.ring = h->ring; o.head = h->head;
    o.lfp_rec = h->lfp_rec; o.n_samples = h->n_samples;
    o.obs = obs; o.reward_f = reward_f; o.reward = h->reward; o.done_out = done_out; o.done_dev = h->done;
    o.step_idx = h->step_idx; o.episode_len = h->episode_len; o.u = h->u;
    o.kind = h->rspec.kind; o.nbins = h->nbins;
    o.power_scale = h->rspec.power_scale; o.action_cost = h->rspec.action_cost;
    o.threshold = h->rspec.threshold; o.threshold_penalty = h->rspec.threshold_penalty;
    o.temp_scale = h->rspec.temp_scale;
    o.lin_g = h->lin_g; o.tw_seed = h->tw_seed; o.tw_inner = h->tw_inner; o.iters = h->obs_iters;
    o.append = append; o.env_ids = ids_dev; o.n_launch = n;
    ++h->n_launches;
    if (append && o.kind != 1 && h->obs_iters == kFastIters && h->nbins <= kFastMaxBins && h->smax <= kObsThreads && !h->no_fast_obs) {
        const size_t smem = obs_fast_smem_bytes(h->nbins, h->rb);
        if (h->f64) obs_kernel_fast<double><<<n, kObsThreads, smem, s>>>(o);
        else obs_kernel_fast<float><<<n, kObsThreads, smem, s>>>(o);
        return cudaGetLastError();
    }
    const size_t smem = obs_smem_bytes(h->W, h->nbins, h->obs_iters, h->rb);
    if (h->f64) obs_kernel<double><<<n, kObsThreads, smem, s>>>(o);
    else obs_kernel<float><<<n, kObsThreads, smem, s>>>(o);
    return cudaGetLastError();
}

// running rfft bins of the given environments from their whole rings (fused-tail state)
cudaError_t launch_spec_init(DbsGymHandle* h, const int32_t* ids_dev, int n, cudaStream_t s) {
    if (!h->fuse_tail) return cudaSuccess;
    ++h->n_launches;
    if (h->f64) spec_init_kernel<double><<<n, kObsThreads, 0, s>>>(h->ring, h->tw_full, h->spec, h->W, h->nbins, kTailBins, ids_dev, n);
    else spec_init_kernel<float><<<n, kObsThreads, 0, s>>>(h->ring, h->tw_full, h->spec, h->W, h->nbins, kTailBins, ids_dev, n);
    return cudaGetLastError();
}

cudaError_t launch_obs_copy(DbsGymHandle* h, float* obs, cudaStream_t s) {
    ++h->n_launches;
    if (h->f64) obs_copy_kernel<double><<<h->B, kCopyThreads, 0, s>>>(h->ring, h->head, obs, h->W, h->B);
    else obs_copy_kernel<float><<<h->B, kCopyThreads, 0, s>>>(h->ring, h->head, obs, h->W, h->B);
    return cudaGetLastError();
}

// Ordering between the handle's private stream and caller streams.  Work the caller queued on ITS stream (dbsgym_step /
// dbsgym_transient with a stream argument) is not ordered against the private stream by CUDA; these two helpers add the
// missing edges with events: own-stream work waits for the last user-stream call, and a user-stream call waits for
// whatever the private stream had queued before it.
cudaError_t enter_own(DbsGymHandle* h) {
    if (!h->user_pending) return cudaSuccess;
    h->user_pending = false;
    return cudaStreamWaitEvent(h->stream, h->ev_user, 0);
}
cudaError_t enter_user(DbsGymHandle* h, cudaStream_t s) {
    if (s == h->stream) return enter_own(h);
    cudaError_t e = cudaEventRecord(h->ev_own, h->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(s, h->ev_own, 0);
    return e;
}
cudaError_t leave_user(DbsGymHandle* h, cudaStream_t s) {
    if (s == h->stream) return cudaSuccess;
    h->user_pending = true;
    return cudaEventRecord(h->ev_user, s);
}

int check_ready(DbsGymHandle* h, bool need_step) {
    if (!h) return DBSGYM_EINVAL;
    if (!h->have_coupling) return fail(h, DBSGYM_ESTATE, "coupling not set (dbsgym_set_coupling_*)");
    if (need_step) {
        if (h->n_sched <= 0) return fail(h, DBSGYM_ESTATE, "schedule not set (dbsgym_set_schedule)");
        if (!h->have_reward) return fail(h, DBSGYM_ESTATE, "reward not set (dbsgym_set_reward)");
    }
    return DBSGYM_OK;
}

cudaStream_t pick_stream(DbsGymHandle* h, void* stream) {
    return stream == DBSGYM_OWN_STREAM ? h->stream : static_cast<cudaStream_t>(stream);
}

int step_impl(DbsGymHandle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
              cudaStream_t s, float* samples_dev = nullptr, int32_t* nsamp_out = nullptr, int32_t* head_out = nullptr) {
    StepParams p;
    fill_params(h, p);
    p.nsamp_out = nsamp_out; p.head_out = head_out;
    p.mode = MODE_STEP; p.actions = actions_dev; p.n_launch = h->B;
    if (h->fuse_tail) {
        // ring append, host mirror, beta-power reward and episode bookkeeping happen in the step kernel's tail;
        // a separate (pure copy) kernel runs only when the caller wants the chronological window on the device
        p.tail_on = 1;
        p.samples_f = samples_dev; p.mirror = h->mirror_on ? h->mirror_dev[h->mirror_cur] : nullptr;
        p.reward_f = reward_dev; p.done_out = done_dev;
        if (h->trace_on) { p.trace = h->trace; p.trace_len = h->trace_len; p.trace_cap = h->trace_cap; }
    }
    if (h->timing) CU(h, cudaEventRecord(h->ev[0], s));
    CU(h, launch_step(h, p, s));
    if (h->timing) CU(h, cudaEventRecord(h->ev[1], s));
    if (!h->fuse_tail) CU(h, launch_obs(h, obs_dev, reward_dev, done_dev, 1, nullptr, h->B, s, samples_dev));
    else if (obs_dev) CU(h, launch_obs_copy(h, obs_dev, s));
    if (h->timing) CU(h, cudaEventRecord(h->ev[2], s));
    return DBSGYM_OK;
}

struct SnapshotHeader {
    uint32_t magic, abi;
    int32_t B, N, Np, W, smax, f64, fsal, nbins_pitch;
    uint64_t bytes;
};
struct SnapPart { void* dev; size_t bytes; };

SnapshotHeader snapshot_header(const DbsGymHandle* h, uint64_t bytes) {
    SnapshotHeader hd;
    memset(&hd, 0, sizeof(hd));
    hd.magic = 0x44425347u; hd.abi = DBSGYM_ABI_VERSION;
    hd.B = h->B; hd.N = h->N; hd.Np = h->Np; hd.W = h->W; hd.smax = h->smax; hd.f64 = h->f64 ? 1 : 0;
    hd.fsal = h->fsal_on ? 1 : 0; hd.nbins_pitch = kTailBins; hd.bytes = bytes;
    return hd;
}

std::vector<SnapPart> snapshot_parts(DbsGymHandle* h) {
    const size_t B = (size_t)h->B, BN = B * h->Np;
    std::vector<SnapPart> v = {
        {h->phase, BN * h->rb}, {h->wind, BN * 4}, {h->w0, BN * h->rb}, {h->stim, BN * h->rb}, {h->rec, BN * h->rb},
        {h->ring, B * h->W * h->rb}, {h->head, B * 4}, {h->spec, B * kTailBins * 2 * sizeof(double)},
        {h->step_idx, B * 4}, {h->episode_len, B * 4}, {h->done, B},
        {h->lfp_true, B * h->smax * 8}, {h->lfp_rec, B * h->smax * 8}, {h->n_samples, B * 4},
        {h->u, B * 8}, {h->reward, B * 8}, {h->counters, 4 * sizeof(unsigned long long)}, {h->status, 4}};
    if (h->fsal_on) { v.push_back({h->k_fsal, BN * h->rb}); v.push_back({h->fsal_valid, B * 4}); }
    return v;
}

// (re)write the windows of the listed environments (NULL = all) into the current mirror buffer
int mirror_refresh(DbsGymHandle* h, const int32_t* ids_dev, int n) {
    const bool was = h->mirror_on;
    h->mirror_on = true;
    cudaError_t e = launch_obs(h, nullptr, nullptr, nullptr, 0, ids_dev, n, h->stream);
    h->mirror_on = was;
    if (e != cudaSuccess) return fail(h, DBSGYM_ECUDA, "mirror refresh failed: %s", cudaGetErrorString(e));
    return DBSGYM_OK;
}

}  // namespace

// ============================================================================================
extern "C" {

int dbsgym_abi_version(void) { return DBSGYM_ABI_VERSION; }

int dbsgym_step_variant(const DbsGymHandle* h, int32_t n_envs) {
    if (!h) return DBSGYM_EINVAL;
    if (h->oct_set >= 0) return 13;
    if (h->lr_rank > 0 && !h->f64) return 11;
    if (h->cluster > 1) return 5;
    if (h->cfg.coupling == DBSGYM_COUPLING_DENSE) return 1;
    if (!h->grid_sym) return 0;
    if (h->f64) return 2;
    if (h->half_set >= 0) return 12;
    if (h->spec_re > 0) return h->warp_set >= 0 ? 10 : 9;
    if (h->cfg.grid[1] == 2 * kRows) return 7;
    if (h->cfg.grid[1] == 4 * kRows) return 8;
    const bool geo1 = h->nthreads == 64 && h->cfg.grid[2] == 8 && h->cfg.grid[0] == 8 && !h->no_geo1;
    if (geo1) {
        const bool mw = kYParity && (h->mw_mode == 1 || (h->mw_mode < 0 && n_envs >= kMwEnvs * h->num_sms));
        return mw ? 4 : 3;
    }
    return h->cfg.grid[0] == 8 ? 6 : 2;
}

int dbsgym_build_flags(void) {
    int f = kYParity ? 1 : 0;
#ifdef DBSGYM_PRECISE_SINCOS
    f |= 2;
#endif
    return f;
}

const char* dbsgym_last_error(const DbsGymHandle* h) { return h ? h->err : g_create_err; }

int dbsgym_create(const DbsGymConfig* cfg, DbsGymHandle** out) {
    if (!cfg || !out) return fail(nullptr, DBSGYM_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->struct_bytes != sizeof(DbsGymConfig))
        return fail(nullptr, DBSGYM_EINVAL, "DbsGymConfig size mismatch: got %u, library expects %zu",
                    cfg->struct_bytes, sizeof(DbsGymConfig));
    if (cfg->n_envs <= 0 || cfg->n_osc <= 0 || cfg->window <= 0 || cfg->max_step_samples <= 0)
        return fail(nullptr, DBSGYM_EINVAL, "n_envs, n_osc, window and max_step_samples must be positive");
    if (cfg->precision != DBSGYM_F32 && cfg->precision != DBSGYM_F64)
        return fail(nullptr, DBSGYM_EINVAL, "unknown precision %d", cfg->precision);
    if (cfg->coupling != DBSGYM_COUPLING_GRID && cfg->coupling != DBSGYM_COUPLING_DENSE)
        return fail(nullptr, DBSGYM_EINVAL, "unknown coupling mode %d", cfg->coupling);
    if (!(cfg->rtol > 0) || !(cfg->atol > 0) || !(cfg->dt0 > 0) || cfg->max_steps <= 0)
        return fail(nullptr, DBSGYM_EINVAL, "rtol, atol, dt0 and max_steps must be positive");
    // DENSE: pad to whole warps (padded oscillators are inert); GRID: whole z-planes of 8-lines
    const int gran = cfg->coupling == DBSGYM_COUPLING_DENSE ? 32 * kRows : kRows;
    int Np = (cfg->n_osc + gran - 1) / gran * gran;
    int dense_cluster = 1;
    if (cfg->coupling == DBSGYM_COUPLING_DENSE && Np / kRows > 1024) {
        // more than 8192 oscillators without a grid: no N x N matrix, the operator must come in low-rank form
        // (dbsgym_set_coupling_lowrank), integrated by a cluster of CTAs with 4096 oscillators each (padded ones are inert)
        while (dense_cluster * 4096 < cfg->n_osc) dense_cluster *= 2;
        if (dense_cluster > 16 || cfg->precision != DBSGYM_F32)
            return fail(nullptr, DBSGYM_EINVAL, "n_osc %d too large for DENSE coupling (fp32: at most 65536 in low-rank form; "
                        "fp64 and the full matrix: at most 8192)", cfg->n_osc);
        Np = dense_cluster * 4096;
    }
    if (cfg->coupling == DBSGYM_COUPLING_GRID) {
        const int gy = cfg->grid[1];
        if (gy != kRows && gy != 2 * kRows && gy != 4 * kRows)
            return fail(nullptr, DBSGYM_EINVAL, "GRID coupling needs grid[1] (gy) == %d, %d or %d, got %d", kRows, 2 * kRows, 4 * kRows, gy);
        if (cfg->grid[0] <= 0 || cfg->grid[2] <= 0 || cfg->n_osc % kRows != 0 ||
            cfg->n_osc > cfg->grid[0] * gy * cfg->grid[2] || cfg->n_osc % (cfg->grid[0] * gy) != 0)
            return fail(nullptr, DBSGYM_EINVAL, "GRID coupling needs n_osc to be whole z-planes of a gx*gy*gz grid");
        if (gy != kRows) {
            // lines of 16 / 32: fp32, mirror-symmetric contraction (even gx and populated gz); the chunk index of a thread
            // must be uniform per warp: a multiple of 8 fundamental (quarter-grid) lines
            const int gzu = cfg->n_osc / (cfg->grid[0] * gy);
            if (cfg->precision != DBSGYM_F32 || cfg->grid[0] % 2 != 0 || gzu % 2 != 0 || ((cfg->grid[0] / 2) * (gzu / 2)) % 8 != 0)
                return fail(nullptr, DBSGYM_EINVAL, "GRID coupling with gy = %d supports fp32, even gx and gz and a multiple of 8 "
                            "fundamental lines; use DENSE", gy);
        }
        if ((cfg->n_osc / kRows) % 32 != 0)
            return fail(nullptr, DBSGYM_EINVAL, "GRID coupling needs a multiple of 32 grid lines (n_osc %% 256 == 0); use DENSE");
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, DBSGYM_ECUDA, "no CUDA device available (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, DBSGYM_EINVAL, "device %d out of range", cfg->device);
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, DBSGYM_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));

    DbsGymHandle* h = new (std::nothrow) DbsGymHandle();
    if (!h) return fail(nullptr, DBSGYM_ENOMEM, "out of host memory");
    h->cfg = *cfg;
    h->N = cfg->n_osc; h->Np = Np; h->B = cfg->n_envs; h->W = cfg->window; h->smax = cfg->max_step_samples;
    h->f64 = cfg->precision == DBSGYM_F64;
    h->rb = h->f64 ? 8 : 4;
    h->nthreads = Np / kRows;
    if (dense_cluster > 1) { h->cluster = dense_cluster; h->nthreads = 512; }
    if (cfg->coupling == DBSGYM_COUPLING_GRID) {
        // more than 512 grid lines (N > 4096): one environment spans a thread-block cluster of 2..16 CTAs
        int want = cfg->force_cluster > 1 ? cfg->force_cluster : 1;             // (test hook: cluster mode at small N)
        while (h->nthreads / want > 512) want *= 2;
        if (want > 1) {
            const int lines = Np / kRows;
            if (want > 16 || lines % (want * 32) != 0 || (cfg->grid[1] == kRows && cfg->grid[0] != 8) || h->f64 ||
                (cfg->n_osc / (cfg->grid[0] * cfg->grid[1])) % 2 != 0) {
                fail(nullptr, DBSGYM_EINVAL, "n_osc %d needs cluster mode (%d CTAs per environment), which supports fp32, "
                     "8 x 8 x gz grids or lines of 16 / 32, even gz and at most 65536 oscillators", cfg->n_osc, want);
                delete h;
                return DBSGYM_EINVAL;
            }
            h->cluster = want;
            h->nthreads = lines / want;
        }
        // only the z-planes actually populated take part
        h->cfg.grid[2] = cfg->n_osc / (cfg->grid[0] * cfg->grid[1]);
        h->tab = h->cfg.grid[2] * cfg->grid[0] * cfg->grid[1];
        // mirror symmetry in z and x needs even extents; DBSGYM_DBG_NO_SYM keeps the plain Toeplitz kernel (A/B runs)
        h->grid_sym = h->cfg.grid[2] % 2 == 0 && cfg->grid[0] % 2 == 0 && !(cfg->debug_flags & DBSGYM_DBG_NO_SYM);
    }
    if (h->cluster == 1) {                               // does one environment's state fit the shared memory of an SM?
        int max_smem = 0;
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
        const size_t need = step_smem_bytes(Np, cfg->coupling == DBSGYM_COUPLING_DENSE ? 0 : h->tab, h->nthreads, h->rb);
        if (need > (size_t)max_smem) {
            fail(nullptr, DBSGYM_EINVAL, "n_osc %d in %s needs %zu bytes of shared memory per environment, the device allows %d",
                 cfg->n_osc, h->f64 ? "fp64" : "fp32", need, max_smem);
            delete h;
            return DBSGYM_EINVAL;
        }
    }
    memset(&h->rspec, 0, sizeof(h->rspec));
    h->ctas_per_sm = cfg->ctas_per_sm > 0 ? cfg->ctas_per_sm : 0;
    h->mw_mode = cfg->mw_mode == 1 ? 0 : cfg->mw_mode == 2 ? 1 : -1;
    h->no_geo1 = (cfg->debug_flags & DBSGYM_DBG_NO_GEO1) != 0;
    h->no_fast_obs = (cfg->debug_flags & DBSGYM_DBG_NO_FAST_OBS) != 0;
    h->no_fused_obs = (cfg->debug_flags & DBSGYM_DBG_NO_FUSED_OBS) != 0;
    h->no_warp = (cfg->debug_flags & DBSGYM_DBG_NO_WARP_KERNEL) != 0;
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
    const size_t BN = (size_t)h->B * Np;
    bool ok = true;
    ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    auto A = [&](void** p, size_t bytes) {
        if (!ok) return;
        ok = cudaMalloc(p, bytes) == cudaSuccess && cudaMemset(*p, 0, bytes) == cudaSuccess;
    };
    A(&h->w0, BN * h->rb); A(&h->stim, BN * h->rb); A(&h->rec, BN * h->rb); A(&h->phase, BN * h->rb);
    A((void**)&h->wind, BN * 4);
    A(&h->ring, (size_t)h->B * h->W * h->rb);
    A((void**)&h->head, (size_t)h->B * 4); A((void**)&h->n_samples, (size_t)h->B * 4);
    A((void**)&h->step_idx, (size_t)h->B * 4); A((void**)&h->episode_len, (size_t)h->B * 4);
    A((void**)&h->lfp_true, (size_t)h->B * h->smax * 8); A((void**)&h->lfp_rec, (size_t)h->B * h->smax * 8);
    A((void**)&h->u, (size_t)h->B * 8); A((void**)&h->reward, (size_t)h->B * 8);
    A((void**)&h->done, (size_t)h->B);
    A((void**)&h->counters, 4 * sizeof(unsigned long long)); A((void**)&h->status, 4);
    h->fsal_on = !h->f64 && !(cfg->debug_flags & DBSGYM_DBG_NO_FSAL_REUSE);
    if (h->fsal_on) { A(&h->k_fsal, BN * h->rb); A((void**)&h->fsal_valid, (size_t)h->B * 4); }
    A((void**)&h->spec, (size_t)h->B * kTailBins * 2 * sizeof(double));
    A((void**)&h->st_actions, (size_t)h->B * 4); A((void**)&h->st_obs, (size_t)h->B * h->W * 4);
    A((void**)&h->st_reward, (size_t)h->B * 4); A((void**)&h->st_done, (size_t)h->B);
    A((void**)&h->st_samples, (size_t)h->B * h->smax * 4);
    if (h->cluster > 1) {
        A(&h->cl_operand, (size_t)h->B * 2 * (2 * (size_t)Np + kScPad) * sizeof(float));
        A((void**)&h->cl_scratch, (size_t)h->B * 2 * h->cluster * kClSlots * sizeof(double));
    }
    A((void**)&h->mpos, (size_t)h->B * 4);
    h->mir_len = h->W + kMirrorGuard;
    ok = ok && cudaMallocHost(&h->pin_ints, (size_t)h->B * 8) == cudaSuccess;
    for (int i = 0; i < 3 && ok; ++i) ok = cudaEventCreate(&h->ev[i]) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_own, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
        // episode_len defaults to "never done"
        std::vector<int32_t> big((size_t)h->B, 0x7fffffff);
        ok = cudaMemcpy(h->episode_len, big.data(), big.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        fail(nullptr, DBSGYM_ENOMEM, "device allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        dbsgym_destroy(h);
        return DBSGYM_ENOMEM;
    }
    *out = h;
    return DBSGYM_OK;
}

void dbsgym_destroy(DbsGymHandle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void* bufs[] = {h->table, h->alpha, h->w0, h->stim, h->rec, h->phase, h->ring, h->wind, h->head, h->n_samples,
                    h->step_idx, h->episode_len, h->lfp_true, h->lfp_rec, h->u, h->reward, h->done, h->sched_nI,
                    h->sched_nII, h->sched_offI, h->sched_offII, h->ts_dev, h->ids_dev, h->lin_g, h->tw_seed,
                    h->tw_inner, h->spec, h->tw_full, h->k_fsal, h->fsal_valid, h->counters, h->status, h->st_actions, h->st_obs, h->st_reward, h->st_done, h->st_samples,
                    h->cl_operand, h->cl_scratch, h->mpos, h->spec_v, h->spec_lam, h->wspec_v, h->wspec_lam, h->lr_v, h->lr_lam, h->lr_soff, h->oct_v, h->oct_lam};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (h->pin_ints) cudaFreeHost(h->pin_ints);
    for (float* m : h->mirror_host)
        if (m) cudaFreeHost(m);
    if (h->ev_own) cudaEventDestroy(h->ev_own);
    if (h->ev_user) cudaEventDestroy(h->ev_user);
    if (h->ctl_host) cudaFreeHost(h->ctl_host);
    if (h->stage_dev) cudaFree(h->stage_dev);
    if (h->stage_host) cudaFreeHost(h->stage_host);
    if (h->trace) cudaFree(h->trace);
    if (h->trace_len) cudaFree(h->trace_len);
    for (int i = 0; i < 3; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int dbsgym_set_coupling_grid(DbsGymHandle* h, const double* table) {
    if (!h || !table) return fail(h, DBSGYM_EINVAL, "null argument");
    if (h->cfg.coupling != DBSGYM_COUPLING_GRID) return fail(h, DBSGYM_ESTATE, "handle was created with DENSE coupling");
    CU(h, cudaSetDevice(h->cfg.device));
    const int GX = h->cfg.grid[0], GZ = h->cfg.grid[2], GY = h->cfg.grid[1], NC = GZ * GX;
    const int n = h->f64 ? 2 : 4;                      // elements per 16 bytes
    std::vector<unsigned char> buf((size_t)h->tab * h->rb);
    for (int c = 0; c < NC; ++c)
        for (int dy = 0; dy < GY; ++dy) {
            const double v = table[(size_t)c * GY + dy];
            const size_t at = ((size_t)(dy / n) * NC + c) * n + dy % n;
            if (h->f64) reinterpret_cast<double*>(buf.data())[at] = v;
            else reinterpret_cast<float*>(buf.data())[at] = (float)v;
        }
    if (!h->table) CU(h, cudaMalloc(&h->table, buf.size()));
    CU(h, cudaMemcpy(h->table, buf.data(), buf.size(), cudaMemcpyHostToDevice));
    h->have_coupling = true;
    if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
    return DBSGYM_OK;
}

// 8 x 8 x 4 half grid (N = 256): tables of warp1_kernel.cuh -- lane l owns the octant point a = l = (zq * 4 + xq) * 4 + yq;
// [mode][lane] eigenvector entries, modes sector after sector, padded to the compiled rank list
static int set_coupling_spectral_half(DbsGymHandle* h, const int32_t* ranks8, int32_t r_max, const double* vecs, const double* vals) {
    int r8[8], compiled[8];
    for (int s8 = 0; s8 < 8; ++s8) {
        if (ranks8[s8] < 0 || ranks8[s8] > r_max) return fail(h, DBSGYM_EINVAL, "spectral rank %d of sector %d outside 0..r_max", ranks8[s8], s8);
        r8[s8] = ranks8[s8];
    }
    const int wset = h->no_warp ? -1 : warp1_kernel_rank_set(r8, compiled);
    if (wset < 0) return fail(h, DBSGYM_ESTATE, "8 x 8 x 4 grid: no compiled rank list covers these sector ranks");
    int nm = 0;
    for (int s8 = 0; s8 < 8; ++s8) nm += compiled[s8];
    const double scale = h->cfg.K / (8.0 * (double)h->N);
    std::vector<float> wv((size_t)32 * nm, 0.f), wlam((size_t)nm, 0.f);
    int off = 0;
    for (int s8 = 0; s8 < 8; ++s8) {
        for (int m = 0; m < ranks8[s8]; ++m) {
            wlam[off + m] = (float)(vals[(size_t)s8 * r_max + m] * scale);
            for (int l = 0; l < 32; ++l) wv[(size_t)(off + m) * 32 + l] = (float)vecs[((size_t)s8 * 32 + l) * r_max + m];
        }
        off += compiled[s8];
    }
    for (float** q : {&h->wspec_v, &h->wspec_lam}) { if (*q) cudaFree(*q); *q = nullptr; }
    CU(h, cudaMalloc(&h->wspec_v, wv.size() * sizeof(float)));
    CU(h, cudaMemcpy(h->wspec_v, wv.data(), wv.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMalloc(&h->wspec_lam, wlam.size() * sizeof(float)));
    CU(h, cudaMemcpy(h->wspec_lam, wlam.data(), wlam.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->half_set = wset;
    if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
    return DBSGYM_OK;
}

int dbsgym_set_coupling_spectral(DbsGymHandle* h, const int32_t* ranks8, int32_t r_max, const double* vecs, const double* vals) {
    if (!h) return DBSGYM_EINVAL;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    if (!ranks8) {                                    // back to the exact sector-block contraction
        h->spec_re = h->spec_ro = 0;
        h->warp_set = h->half_set = -1;
        if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
        return DBSGYM_OK;
    }
    if (!vecs || !vals) return fail(h, DBSGYM_EINVAL, "null argument");
    if (h->lr_rank > 0) return fail(h, DBSGYM_ESTATE, "the handle already runs the low-rank form of its operator");
    const bool grid_ok = h->cfg.coupling == DBSGYM_COUPLING_GRID && !h->f64 && h->grid_sym && h->cluster <= 1 &&
                         h->cfg.grid[0] == 8 && h->cfg.grid[1] == 8;
    if (grid_ok && h->cfg.grid[2] == 4 && h->nthreads == 32) return set_coupling_spectral_half(h, ranks8, r_max, vecs, vals);
    if (!grid_ok || h->nthreads != 64 || h->cfg.grid[2] != 8)
        return fail(h, DBSGYM_ESTATE, "the spectral contraction serves fp32 GRID handles on the 8 x 8 x 8 and 8 x 8 x 4 grids");
    int r_even = 1, r_odd = 1;
    for (int s8 = 0; s8 < 8; ++s8) {
        if (ranks8[s8] < 0 || ranks8[s8] > 9 || ranks8[s8] > r_max)
            return fail(h, DBSGYM_EINVAL, "spectral rank %d of sector %d outside 0..min(9, r_max)", ranks8[s8], s8);
        if (s8 < 4) r_even = std::max(r_even, (int)ranks8[s8]); else r_odd = std::max(r_odd, (int)ranks8[s8]);
    }
    const double scale = h->cfg.K / (8.0 * (double)h->N);
    // ---- tables of the 64-thread worker kernel (step_kernel<CPL_SPECTRAL>): ranks padded to (9, 4) or (9, 9) ----
    const int RE = 9, RO = r_odd <= 4 ? 4 : 9, R = RE + RO;
    std::vector<float> v((size_t)kMwThreads * 4 * R, 0.f), lam((size_t)4 * R, 0.f);
    for (int t = 0; t < kMwThreads; ++t) {
        int zq, xq, sec;
        mw_decode(t, zq, xq, sec);
        const int q = zq * 4 + xq;
        for (int j = 0; j < 4; ++j)
            for (int m = 0; m < R; ++m) {
                const int s8 = m < RE ? sec : 4 + sec, mm = m < RE ? m : m - RE;
                if (mm < ranks8[s8]) v[((size_t)t * 4 + j) * R + m] = (float)vecs[((size_t)s8 * 64 + q * 4 + j) * r_max + mm];
            }
    }
    for (int sg = 0; sg < 4; ++sg)
        for (int m = 0; m < R; ++m) {
            const int s8 = m < RE ? sg : 4 + sg, mm = m < RE ? m : m - RE;
            if (mm < ranks8[s8]) lam[(size_t)sg * R + m] = (float)(vals[(size_t)s8 * r_max + mm] * scale);
        }
    // ---- tables of the one-warp-per-environment kernel (warp_kernel.cuh): lane l owns the octant points
    //      a_p = (zq * 4 + xq) * 4 + 2 * yp + p, zq = l >> 3, xq = (l >> 1) & 3, yp = l & 1; modes sector after sector ----
    int compiled[8];
    const int wset = h->no_warp ? -1 : warp_kernel_rank_set(ranks8, compiled);
    std::vector<float> wv, wlam;
    if (wset >= 0) {
        int nm = 0;
        for (int s8 = 0; s8 < 8; ++s8) nm += compiled[s8];
        wv.assign((size_t)32 * nm * 2, 0.f);
        wlam.assign((size_t)nm, 0.f);
        int off = 0;
        for (int s8 = 0; s8 < 8; ++s8) {
            for (int m = 0; m < compiled[s8] && m < ranks8[s8]; ++m) {
                wlam[off + m] = (float)(vals[(size_t)s8 * r_max + m] * scale);
                for (int l = 0; l < 32; ++l)
                    for (int pt = 0; pt < 2; ++pt) {
                        const int a = ((l >> 3) * 4 + ((l >> 1) & 3)) * 4 + 2 * (l & 1) + pt;
                        const int mode = off + m;                      // [mode pair][lane](V0[m], V1[m], V0[m + 1], V1[m + 1])
                        wv[(((size_t)(mode >> 1) * 32 + l) * 2 + (mode & 1)) * 2 + pt] = (float)vecs[((size_t)s8 * 64 + a) * r_max + m];
                    }
            }
            off += compiled[s8];
        }
    }
    for (float** q : {&h->spec_v, &h->spec_lam, &h->wspec_v, &h->wspec_lam}) { if (*q) cudaFree(*q); *q = nullptr; }
    auto upload = [&](float** dst, const std::vector<float>& src) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, src.size() * sizeof(float));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice);
    };
    CU(h, upload(&h->spec_v, v));
    CU(h, upload(&h->spec_lam, lam));
    if (wset >= 0) { CU(h, upload(&h->wspec_v, wv)); CU(h, upload(&h->wspec_lam, wlam)); }
    h->spec_re = RE; h->spec_ro = RO; h->warp_set = wset;
    if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
    return DBSGYM_OK;
}

int dbsgym_set_coupling_dense(DbsGymHandle* h, const double* alpha) {
    if (!h || !alpha) return fail(h, DBSGYM_EINVAL, "null argument");
    if (h->cfg.coupling != DBSGYM_COUPLING_DENSE) return fail(h, DBSGYM_ESTATE, "handle was created with GRID coupling");
    if (h->cluster > 1) return fail(h, DBSGYM_ESTATE, "n_osc %d: the full matrix is not kept above 8192 oscillators, use dbsgym_set_coupling_lowrank", h->N);
    CU(h, cudaSetDevice(h->cfg.device));
    const int N = h->N, Np = h->Np;
    std::vector<unsigned char> buf((size_t)Np * Np * h->rb, 0);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {                  // stored transposed: [j][i]
            const double v = alpha[(size_t)i * N + j];
            if (h->f64) reinterpret_cast<double*>(buf.data())[(size_t)j * Np + i] = v;
            else reinterpret_cast<float*>(buf.data())[(size_t)j * Np + i] = (float)v;
        }
    if (!h->alpha) CU(h, cudaMalloc(&h->alpha, buf.size()));
    CU(h, cudaMemcpy(h->alpha, buf.data(), buf.size(), cudaMemcpyHostToDevice));
    h->have_coupling = true;
    if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
    return DBSGYM_OK;
}

int dbsgym_set_coupling_lowrank(DbsGymHandle* h, int32_t rank, const double* vecs, const double* vals) {
    if (!h) return DBSGYM_EINVAL;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    if (h->lr_sectors) return fail(h, DBSGYM_ESTATE, "the handle stores its oscillators in octant order (sector form): create a new handle");
    for (float** q : {&h->lr_v, &h->lr_lam}) { if (*q) cudaFree(*q); *q = nullptr; }
    h->lr_rank = 0;
    if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
    const bool dense = h->cfg.coupling == DBSGYM_COUPLING_DENSE;
    if (rank <= 0) {                                  // back to the full operator (DENSE: needs dbsgym_set_coupling_dense)
        h->have_coupling = dense ? h->alpha != nullptr : h->table != nullptr;
        return DBSGYM_OK;
    }
    if (!vecs || !vals) return fail(h, DBSGYM_EINVAL, "null argument");
    if (h->f64) return fail(h, DBSGYM_ESTATE, "the low-rank contraction serves fp32 handles (fp64 parity mode evaluates the full sum)");
    const int R4 = (rank + 3) / 4 * 4, N = h->N, Np = h->Np;
    if (R4 > 1024) return fail(h, DBSGYM_EINVAL, "rank %d too large (max 1024)", rank);
    if (Np % 256 != 0) return fail(h, DBSGYM_ESTATE, "the low-rank kernel needs whole warps of 8-oscillator threads (padded n_osc %d)", Np);
    const size_t need = h->cluster > 1 ? step_smem_bytes_cluster_lr(h->nthreads, R4) : step_smem_bytes(Np, 2 * R4, h->nthreads, h->rb);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->cfg.device);
    if (need > (size_t)max_smem) return fail(h, DBSGYM_EINVAL, "rank %d needs %zu bytes of shared memory, the device allows %d", rank, need, max_smem);
    std::vector<float> v((size_t)R4 * Np, 0.f), lam((size_t)R4, 0.f);
    for (int m = 0; m < rank; ++m) {
        lam[m] = (float)vals[m];
        for (int i = 0; i < N; ++i) v[(size_t)m * Np + i] = (float)vecs[(size_t)m * N + (h->perm.empty() ? i : h->perm[i])];
    }
    CU(h, cudaMalloc(&h->lr_v, v.size() * sizeof(float)));
    CU(h, cudaMalloc(&h->lr_lam, lam.size() * sizeof(float)));
    CU(h, cudaMemcpy(h->lr_v, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->lr_lam, lam.data(), lam.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->lr_rank = R4;
    h->have_coupling = true;
    return DBSGYM_OK;
}

int dbsgym_set_oscillator_order(DbsGymHandle* h, const int32_t* order) {
    if (!h) return DBSGYM_EINVAL;
    if (h->params_set || h->lr_rank > 0)
        return fail(h, DBSGYM_ESTATE, "set the oscillator order right after dbsgym_create, before any per-oscillator data are uploaded");
    if (h->f64) return fail(h, DBSGYM_ESTATE, "oscillator orders are applied on fp32 handles (fp64 parity mode keeps the caller's order)");
    if (h->N != h->Np) return fail(h, DBSGYM_ESTATE, "a padded handle keeps the caller's order");
    h->user_order.clear();
    if (order) {
        std::vector<uint8_t> seen((size_t)h->N, 0);
        for (int d = 0; d < h->N; ++d) {
            if (order[d] < 0 || order[d] >= h->N || seen[order[d]]) return fail(h, DBSGYM_EINVAL, "order is not a permutation of 0 .. n_osc - 1");
            seen[order[d]] = 1;
        }
        h->user_order.assign(order, order + h->N);
    }
    h->recompute_order();
    return DBSGYM_OK;
}

int dbsgym_set_coupling_lowrank_sectors(DbsGymHandle* h, const int32_t* soff9, const double* zvecs, const double* vals) {
    if (!h || !soff9 || !zvecs || !vals) return fail(h, DBSGYM_EINVAL, "null argument");
    if (h->cfg.coupling != DBSGYM_COUPLING_GRID || h->f64)
        return fail(h, DBSGYM_ESTATE, "the sector form of the low-rank operator serves fp32 GRID handles");
    if (h->params_set || h->lr_rank > 0)
        return fail(h, DBSGYM_ESTATE, "set the sector form right after dbsgym_set_coupling_grid, before any environment vectors are "
                                      "uploaded (it changes the order the oscillators are stored in)");
    const int gx = h->cfg.grid[0], gy = h->cfg.grid[1], gz = h->cfg.grid[2], N = h->N, Np = h->Np;
    if (gx % 2 || gy % 2 || gz % 2 || N != gx * gy * gz || Np != N || Np % 256 != 0)
        return fail(h, DBSGYM_ESTATE, "the sector form needs a full grid with even extents and whole warps of octant points");
    const int R = soff9[8];
    if (soff9[0] != 0 || R <= 0 || R > 1024) return fail(h, DBSGYM_EINVAL, "mode offsets out of range (at most 1024 modes)");
    for (int s8 = 0; s8 < 8; ++s8)
        if (soff9[s8 + 1] < soff9[s8] || (soff9[s8 + 1] - soff9[s8]) % 4 != 0)
            return fail(h, DBSGYM_EINVAL, "the modes of every sector must be padded to a multiple of 4");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    const size_t need = h->cluster > 1 ? step_smem_bytes_cluster_lr(h->nthreads, R) : step_smem_bytes(Np, 2 * R, h->nthreads, h->rb);
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->cfg.device);
    if (need > (size_t)max_smem) return fail(h, DBSGYM_EINVAL, "%d modes need %zu bytes of shared memory, the device allows %d", R, need, max_smem);
    // octant order: device position 8 a + g holds the image g = 4 my + 2 mz + mx of octant point a = (zq * gx/2 + xq) * gy/2 + yq
    const int hx = gx / 2, hy = gy / 2, hz = gz / 2, P8 = N / 8;
    h->internal_order.assign((size_t)N, 0);
    for (int zq = 0; zq < hz; ++zq)
        for (int xq = 0; xq < hx; ++xq)
            for (int yq = 0; yq < hy; ++yq) {
                const int a = (zq * hx + xq) * hy + yq;
                for (int g = 0; g < 8; ++g) {
                    const int z = (g & 2) ? gz - 1 - zq : zq, x = (g & 1) ? gx - 1 - xq : xq, y = (g & 4) ? gy - 1 - yq : yq;
                    h->internal_order[(size_t)a * 8 + g] = (z * gx + x) * gy + y;
                }
            }
    h->recompute_order();
    std::vector<float> z((size_t)R * P8), lam((size_t)R);
    for (int m = 0; m < R; ++m) {
        lam[m] = (float)(vals[m] / 8.0);               // (alpha x)[g a] = 1/8 sum_s chi_s(g) (block_s X_s)[a]
        for (int a = 0; a < P8; ++a) z[(size_t)m * P8 + a] = (float)zvecs[(size_t)m * P8 + a];
    }
    CU(h, cudaMalloc(&h->lr_v, z.size() * sizeof(float)));
    CU(h, cudaMalloc(&h->lr_lam, lam.size() * sizeof(float)));
    CU(h, cudaMalloc(&h->lr_soff, 9 * sizeof(int32_t)));
    CU(h, cudaMemcpy(h->lr_v, z.data(), z.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->lr_lam, lam.data(), lam.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->lr_soff, soff9, 9 * sizeof(int32_t), cudaMemcpyHostToDevice));
    h->lr_rank = R;
    h->lr_sectors = true;
    h->have_coupling = true;
    // 1024 ... 4096 oscillators in one CTA, 8192 in a cluster of 2: the register-resident kernel of oct_kernel.cuh when a compiled rank list covers
    // the sectors' ranks (modes with a non-zero eigenvalue; the padding of every sector's block is zero)
    h->oct_set = -1;
    const bool one_cta = h->cluster <= 1 && (N == 1024 || N == 2048 || N == 4096) && h->nthreads == P8;
    const bool clustered = N == 8192 && h->cluster == 2 && h->nthreads == 512;     // a cluster of 2 CTAs of 4096 oscillators
    if (!h->no_warp && (one_cta || clustered)) {
        int ranks8[8], compiled[8];
        for (int s8 = 0; s8 < 8; ++s8) {
            ranks8[s8] = 0;
            for (int m = soff9[s8]; m < soff9[s8 + 1]; ++m)
                if (vals[m] != 0.0) ranks8[s8] = m - soff9[s8] + 1;
        }
        const int oset = oct_kernel_rank_set(N, ranks8, compiled);
        if (oset >= 0) {
            int nm = 0;
            for (int s8 = 0; s8 < 8; ++s8) nm += compiled[s8];
            const double scale = h->cfg.K / (8.0 * (double)N);
            std::vector<float> ov((size_t)nm * P8, 0.f), olam((size_t)nm, 0.f);
            int off = 0;
            for (int s8 = 0; s8 < 8; ++s8) {
                for (int j = 0; j < ranks8[s8]; ++j) {
                    const int m = soff9[s8] + j;
                    olam[off + j] = (float)(vals[m] * scale);
                    for (int a = 0; a < P8; ++a) ov[(size_t)(off + j) * P8 + a] = (float)zvecs[(size_t)m * P8 + a];
                }
                off += compiled[s8];
            }
            CU(h, cudaMalloc(&h->oct_v, ov.size() * sizeof(float)));
            CU(h, cudaMalloc(&h->oct_lam, olam.size() * sizeof(float)));
            CU(h, cudaMemcpy(h->oct_v, ov.data(), ov.size() * sizeof(float), cudaMemcpyHostToDevice));
            CU(h, cudaMemcpy(h->oct_lam, olam.data(), olam.size() * sizeof(float), cudaMemcpyHostToDevice));
            h->oct_set = oset;
        }
    }
    if (h->fsal_valid) CU(h, cudaMemset(h->fsal_valid, 0, (size_t)h->B * 4));
    return DBSGYM_OK;
}

int dbsgym_set_recording(DbsGymHandle* h, int32_t weighted) {
    if (!h) return DBSGYM_EINVAL;
    h->weighted_rec = weighted ? 1 : 0;
    return DBSGYM_OK;
}

int dbsgym_set_env_params(DbsGymHandle* h, const int32_t* env_ids, int32_t n, const double* w0,
                          const double* stim_cond, const double* rec_cond, const double* y0) {
    if (!h) return DBSGYM_EINVAL;
    if (n <= 0 || n > h->B) return fail(h, DBSGYM_EINVAL, "n=%d out of range", n);
    CU(h, cudaSetDevice(h->cfg.device));
    const int32_t* ids = nullptr;
    CU(h, enter_own(h));
    int rc = upload_ids(h, env_ids, n, &ids);
    if (rc) return rc;
    if (h->fsal_valid) {                            // the carried stage derivative belongs to the old vectors
        std::vector<int32_t> z((size_t)n, 0);
        rc = scatter_to_device(h, h->fsal_valid, z.data(), ids, n, 4);
        if (rc) return rc;
    }
    h->params_set = true;
    // sector form of the low-rank operator: the device keeps the oscillators in octant order
    std::vector<double> pv[4];
    if (!h->perm.empty()) {
        const double** srcs[4] = {&w0, &stim_cond, &rec_cond, &y0};
        for (int k = 0; k < 4; ++k) {
            if (!*srcs[k]) continue;
            pv[k].resize((size_t)n * h->N);
            for (int r = 0; r < n; ++r) {
                const double* a = *srcs[k] + (size_t)r * h->N;
                double* b = pv[k].data() + (size_t)r * h->N;
                for (int d = 0; d < h->N; ++d) b[d] = a[h->perm[d]];
            }
            *srcs[k] = pv[k].data();
        }
    }
    std::vector<unsigned char> buf;
    const size_t row = (size_t)h->Np * h->rb;
    struct { const double* src; void* dst; } vecs[3] = {{w0, h->w0}, {stim_cond, h->stim}, {rec_cond, h->rec}};
    rc = ensure_stage(h, row * (size_t)n);
    if (rc) return rc;
    for (auto& v : vecs) {
        if (!v.src) continue;
        if (h->f64) to_real_rows_stage<double>(v.src, n, h->N, h->Np, h->stage_host);
        else to_real_rows_stage<float>(v.src, n, h->N, h->Np, h->stage_host);
        rc = scatter_to_device(h, v.dst, h->stage_host, ids, n, row);
        if (rc) return rc;
    }
    if (y0) {
        if (h->f64) {
            to_real_rows<double>(y0, n, h->N, h->Np, buf);
            rc = scatter_to_device(h, h->phase, buf.data(), ids, n, row);
            if (rc) return rc;
        } else {
            // fp32 state: wrapped phase + integer winding count, y = phase + 2*pi*wind.  Both arrays are written straight into
            // the pinned staging buffer (phases in its first half, winding counts in the second), one copy, two scatters.
            const size_t half = row * (size_t)n;
            rc = ensure_stage(h, 2 * half);
            if (rc) return rc;
            float* ph = reinterpret_cast<float*>(h->stage_host);
            int32_t* wd = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(h->stage_host) + half);
            for (int r = 0; r < n; ++r) {
                const double* yr = y0 + (size_t)r * h->N;
                float* pr = ph + (size_t)r * h->Np;
                int32_t* wr = wd + (size_t)r * h->Np;
                for (int i = 0; i < h->N; ++i) {
                    const double y = yr[i];
                    if (y >= 0.0 && y < kTwoPi) { pr[i] = (float)y; wr[i] = 0; }          // (a freshly drawn phase: N(pi, 0.6))
                    else {
                        const double k = std::floor(y / kTwoPi);
                        pr[i] = (float)(y - k * kTwoPi);
                        wr[i] = (int32_t)k;
                    }
                }
                for (int i = h->N; i < h->Np; ++i) { pr[i] = 0.f; wr[i] = 0; }
            }
            cudaError_t e = cudaMemcpyAsync(h->stage_dev, h->stage_host, 2 * half, cudaMemcpyHostToDevice, h->stream);
            if (e == cudaSuccess) {
                scatter_rows_kernel<<<n, 128, 0, h->stream>>>(reinterpret_cast<unsigned char*>(h->phase),
                                                              static_cast<const unsigned char*>(h->stage_dev), ids, n, row);
                scatter_rows_kernel<<<n, 128, 0, h->stream>>>(reinterpret_cast<unsigned char*>(h->wind),
                                                              static_cast<const unsigned char*>(h->stage_dev) + half, ids, n, row);
                h->n_launches += 2;
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);      // the staging pair is reused by the next call
            if (e != cudaSuccess) return fail(h, DBSGYM_ECUDA, "scatter failed: %s", cudaGetErrorString(e));
        }
    }
    return DBSGYM_OK;
}

int dbsgym_set_schedule(DbsGymHandle* h, int32_t n_steps, const int32_t* n_I, const int32_t* n_II,
                        const double* offs_I, int32_t max_I, const double* offs_II, int32_t max_II) {
    if (!h || !n_I || !n_II || !offs_I || !offs_II) return fail(h, DBSGYM_EINVAL, "null argument");
    if (n_steps <= 0 || max_I <= 0 || max_II <= 0) return fail(h, DBSGYM_EINVAL, "bad schedule sizes");
    for (int k = 0; k < n_steps; ++k) {
        if (n_I[k] < 2 || n_I[k] > max_I || n_II[k] < 2 || n_II[k] > max_II)
            return fail(h, DBSGYM_EINVAL, "schedule step %d: segment lengths (%d,%d) out of range", k, n_I[k], n_II[k]);
        if (n_I[k] + n_II[k] - 1 > h->smax)
            return fail(h, DBSGYM_EINVAL, "schedule step %d needs %d samples > max_step_samples %d", k,
                        n_I[k] + n_II[k] - 1, h->smax);
        if (offs_I[(size_t)k * max_I] != 0.0 || offs_II[(size_t)k * max_II] != 0.0)
            return fail(h, DBSGYM_EINVAL, "schedule offsets must start at 0");
    }
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    for (void* b : {(void*)h->sched_nI, (void*)h->sched_nII, (void*)h->sched_offI, (void*)h->sched_offII})
        if (b) cudaFree(b);
    h->sched_nI = h->sched_nII = nullptr; h->sched_offI = h->sched_offII = nullptr; h->n_sched = 0;
    CU(h, cudaMalloc(&h->sched_nI, (size_t)n_steps * 4));
    CU(h, cudaMalloc(&h->sched_nII, (size_t)n_steps * 4));
    CU(h, cudaMalloc(&h->sched_offI, (size_t)n_steps * max_I * 8));
    CU(h, cudaMalloc(&h->sched_offII, (size_t)n_steps * max_II * 8));
    CU(h, cudaMemcpy(h->sched_nI, n_I, (size_t)n_steps * 4, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->sched_nII, n_II, (size_t)n_steps * 4, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->sched_offI, offs_I, (size_t)n_steps * max_I * 8, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->sched_offII, offs_II, (size_t)n_steps * max_II * 8, cudaMemcpyHostToDevice));
    h->maxI = max_I; h->maxII = max_II; h->n_sched = n_steps;
    return DBSGYM_OK;
}

int dbsgym_set_reward(DbsGymHandle* h, const DbsGymRewardSpec* spec, const double* lin_functional) {
    if (!h || !spec) return fail(h, DBSGYM_EINVAL, "null argument");
    if (spec->struct_bytes != sizeof(DbsGymRewardSpec)) return fail(h, DBSGYM_EINVAL, "DbsGymRewardSpec size mismatch");
    if (spec->kind < 0 || spec->kind > 2) return fail(h, DBSGYM_EINVAL, "unknown reward kind %d", spec->kind);
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    const int W = h->W;
    if (spec->kind == DBSGYM_REWARD_TEMP_CONST) {
        if (!lin_functional) return fail(h, DBSGYM_EINVAL, "TEMP_CONST reward needs the linear functional");
        if (!h->lin_g) CU(h, cudaMalloc(&h->lin_g, (size_t)W * 8));
        CU(h, cudaMemcpy(h->lin_g, lin_functional, (size_t)W * 8, cudaMemcpyHostToDevice));
        h->nbins = 0;
    } else {
        const int nb = spec->bin_hi - spec->bin_lo + 1;
        if (spec->bin_lo < 0 || nb <= 0 || nb > kMaxBins || spec->bin_hi > W / 2)
            return fail(h, DBSGYM_EINVAL, "bad rfft bin range [%d,%d] (at most %d bins)", spec->bin_lo, spec->bin_hi, kMaxBins);
        const int iters = ((W + kObsThreads - 1) / kObsThreads + 1) & ~1;   // even: table rows stay 16-byte aligned
        std::vector<double> seed((size_t)nb * kObsThreads * 2);
        std::vector<unsigned char> inner((size_t)nb * iters * 2 * h->rb);
        const double two_pi = kTwoPi;
        for (int b = 0; b < nb; ++b) {
            const long long k = spec->bin_lo + b;
            for (int m = 0; m < kObsThreads; ++m) {
                const long long q = (k * m) % W;       // exact integer phase reduction
                seed[((size_t)b * kObsThreads + m) * 2] = std::cos(two_pi * (double)q / W);
                seed[((size_t)b * kObsThreads + m) * 2 + 1] = std::sin(two_pi * (double)q / W);
            }
            for (int i = 0; i < iters; ++i) {
                const long long q = (k * (long long)kObsThreads * i) % W;
                const double c = std::cos(two_pi * (double)q / W), sn = std::sin(two_pi * (double)q / W);
                const size_t at = ((size_t)b * iters + i) * 2;
                if (h->f64) { reinterpret_cast<double*>(inner.data())[at] = c; reinterpret_cast<double*>(inner.data())[at + 1] = sn; }
                else { reinterpret_cast<float*>(inner.data())[at] = (float)c; reinterpret_cast<float*>(inner.data())[at + 1] = (float)sn; }
            }
        }
        if (h->tw_seed) cudaFree(h->tw_seed);
        if (h->tw_inner) cudaFree(h->tw_inner);
        h->tw_seed = nullptr; h->tw_inner = nullptr;
        CU(h, cudaMalloc(&h->tw_seed, seed.size() * 8));
        CU(h, cudaMalloc(&h->tw_inner, inner.size()));
        CU(h, cudaMemcpy(h->tw_seed, seed.data(), seed.size() * 8, cudaMemcpyHostToDevice));
        CU(h, cudaMemcpy(h->tw_inner, inner.data(), inner.size(), cudaMemcpyHostToDevice));
        h->obs_iters = iters;
        h->nbins = nb;
        // fused observation tail: twiddles of every ring position, [W][nb][2] float64 (exact integer phase reduction)
        std::vector<double> full((size_t)W * nb * 2);
        for (int m = 0; m < W; ++m)
            for (int b = 0; b < nb; ++b) {
                const long long q = ((long long)(spec->bin_lo + b) * m) % W;
                full[((size_t)m * nb + b) * 2] = std::cos(two_pi * (double)q / W);
                full[((size_t)m * nb + b) * 2 + 1] = std::sin(two_pi * (double)q / W);
            }
        if (h->tw_full) cudaFree(h->tw_full);
        h->tw_full = nullptr;
        CU(h, cudaMalloc(&h->tw_full, full.size() * 8));
        CU(h, cudaMemcpy(h->tw_full, full.data(), full.size() * 8, cudaMemcpyHostToDevice));
    }
    h->rspec = *spec;
    h->have_reward = true;
    h->fuse_tail = spec->kind != DBSGYM_REWARD_TEMP_CONST && h->nbins <= kTailBins && h->smax <= 32 && h->smax <= W && !h->no_fused_obs;
    if (h->fuse_tail) {                              // the bins of whatever the rings hold now
        CU(h, launch_spec_init(h, nullptr, h->B, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return DBSGYM_OK;
}

int dbsgym_set_episode(DbsGymHandle* h, const int32_t* env_ids, int32_t n, const int32_t* step_idx,
                       const int32_t* episode_len) {
    if (!h) return DBSGYM_EINVAL;
    if (n <= 0 || n > h->B) return fail(h, DBSGYM_EINVAL, "n=%d out of range", n);
    CU(h, cudaSetDevice(h->cfg.device));
    const int32_t* ids = nullptr;
    CU(h, enter_own(h));
    int rc = upload_ids(h, env_ids, n, &ids);
    if (rc) return rc;
    if (step_idx) { rc = scatter_to_device(h, h->step_idx, step_idx, ids, n, 4); if (rc) return rc; }
    if (episode_len) { rc = scatter_to_device(h, h->episode_len, episode_len, ids, n, 4); if (rc) return rc; }
    if (step_idx) {
        std::vector<uint8_t> z((size_t)n, 0);
        // done flags of re-armed environments are cleared (1-byte rows cannot use the 4-byte scatter)
        std::vector<uint8_t> all((size_t)h->B);
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaMemcpy(all.data(), h->done, (size_t)h->B, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i) all[env_ids ? env_ids[i] : i] = 0;
        CU(h, cudaMemcpy(h->done, all.data(), (size_t)h->B, cudaMemcpyHostToDevice));
    }
    return DBSGYM_OK;
}

int dbsgym_transient(DbsGymHandle* h, const int32_t* env_ids, int32_t n, const double* ts_offsets, int32_t n_ts,
                     float* obs_dev, void* stream) {
    int rc = check_ready(h, false);
    if (rc) return rc;
    if (!ts_offsets || n_ts < 2) return fail(h, DBSGYM_EINVAL, "need at least two sample times");
    if (n <= 0 || n > h->B) return fail(h, DBSGYM_EINVAL, "n=%d out of range", n);
    if (ts_offsets[0] != 0.0) return fail(h, DBSGYM_EINVAL, "ts_offsets must start at 0");
    if (n_ts - 1 < h->W) return fail(h, DBSGYM_EINVAL, "transient gives %d samples < window %d (env.py:303-304)", n_ts - 1, h->W);
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = pick_stream(h, stream);
    const int32_t* ids = nullptr;
    rc = upload_ids(h, env_ids, n, &ids);
    if (rc) return rc;
    if (n_ts > h->ts_cap) {
        CU(h, cudaStreamSynchronize(s));
        if (h->ts_dev) cudaFree(h->ts_dev);
        h->ts_dev = nullptr;
        CU(h, cudaMalloc(&h->ts_dev, (size_t)n_ts * 8));
        h->ts_cap = n_ts;
    }
    CU(h, cudaMemcpyAsync(h->ts_dev, ts_offsets, (size_t)n_ts * 8, cudaMemcpyHostToDevice, s));
    CU(h, cudaStreamSynchronize(s));               // ts_offsets is caller memory
    StepParams p;
    fill_params(h, p);
    p.mode = MODE_TRANSIENT; p.env_ids = ids; p.n_launch = n; p.ts = h->ts_dev; p.n_ts = n_ts;
    CU(h, enter_user(h, s));
    CU(h, launch_step(h, p, s));
    CU(h, launch_spec_init(h, ids, n, s));
    const bool mir = h->mirror_on;
    h->mirror_on = false;
    cudaError_t eo = obs_dev ? launch_obs(h, obs_dev, nullptr, nullptr, 0, ids, n, s) : cudaSuccess;
    h->mirror_on = mir;
    CU(h, eo);
    if (mir) {
        // A reset replaces whole windows.  Windows handed out earlier must survive it (a caller may still hold the
        // observation of the previous step), so the mirror moves on to its other buffer: every log restarts at column 0
        // there and ALL environments' windows are written afresh (which also re-aligns their positions).
        h->mirror_cur ^= 1;
        CU(h, cudaMemsetAsync(h->mpos, 0, (size_t)h->B * 4, s));
        CU(h, launch_obs(h, nullptr, nullptr, nullptr, 0, nullptr, h->B, s));
    }
    CU(h, leave_user(h, s));
    if (ids) CU(h, cudaStreamSynchronize(s));      // ids_dev is reused by the next call
    return DBSGYM_OK;
}

int dbsgym_step(DbsGymHandle* h, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                void* stream) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!actions_dev) return fail(h, DBSGYM_EINVAL, "actions_dev is null");
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = pick_stream(h, stream);
    CU(h, enter_user(h, s));
    rc = step_impl(h, actions_dev, obs_dev, reward_dev, done_dev, s);
    if (rc) return rc;
    CU(h, leave_user(h, s));
    return DBSGYM_OK;
}

int dbsgym_step_host(DbsGymHandle* h, const float* actions, float* obs, float* reward, uint8_t* done) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!actions) return fail(h, DBSGYM_EINVAL, "actions is null");
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->stream;
    CU(h, enter_own(h));
    CU(h, cudaMemcpyAsync(h->st_actions, actions, (size_t)h->B * 4, cudaMemcpyHostToDevice, s));
    rc = step_impl(h, h->st_actions, obs ? h->st_obs : nullptr, h->st_reward, h->st_done, s);
    if (rc) return rc;
    if (obs) CU(h, cudaMemcpyAsync(obs, h->st_obs, (size_t)h->B * h->W * 4, cudaMemcpyDeviceToHost, s));
    if (reward) CU(h, cudaMemcpyAsync(reward, h->st_reward, (size_t)h->B * 4, cudaMemcpyDeviceToHost, s));
    if (done) CU(h, cudaMemcpyAsync(done, h->st_done, (size_t)h->B, cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    return DBSGYM_OK;
}

int dbsgym_step_host_samples(DbsGymHandle* h, const float* actions, float* samples, int32_t* n_samples,
                             float* reward, uint8_t* done) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!actions || !samples || !n_samples) return fail(h, DBSGYM_EINVAL, "null argument");
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->stream;
    CU(h, enter_own(h));
    CU(h, cudaMemcpyAsync(h->st_actions, actions, (size_t)h->B * 4, cudaMemcpyHostToDevice, s));
    rc = step_impl(h, h->st_actions, nullptr, h->st_reward, h->st_done, s, h->st_samples);
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(samples, h->st_samples, (size_t)h->B * h->smax * 4, cudaMemcpyDeviceToHost, s));
    CU(h, cudaMemcpyAsync(n_samples, h->n_samples, (size_t)h->B * 4, cudaMemcpyDeviceToHost, s));
    if (reward) CU(h, cudaMemcpyAsync(reward, h->st_reward, (size_t)h->B * 4, cudaMemcpyDeviceToHost, s));
    if (done) CU(h, cudaMemcpyAsync(done, h->st_done, (size_t)h->B, cudaMemcpyDeviceToHost, s));
    CU(h, cudaStreamSynchronize(s));
    return DBSGYM_OK;
}

int dbsgym_host_mirror(DbsGymHandle* h, float** mirror, int32_t* row_floats) {
    if (!h || !mirror) return fail(h, DBSGYM_EINVAL, "null argument");
    CU(h, cudaSetDevice(h->cfg.device));
    if (!h->mirror_host[0]) {
        const size_t bytes = (size_t)h->B * 2 * h->mir_len * sizeof(float);
        for (int i = 0; i < 2; ++i) {
            CU(h, cudaHostAlloc(reinterpret_cast<void**>(&h->mirror_host[i]), bytes, cudaHostAllocMapped | cudaHostAllocPortable));
            CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->mirror_dev[i]), h->mirror_host[i], 0));
        }
        // fill the first buffer from the current device windows, every log starting at column 0
        CU(h, enter_own(h));
        CU(h, cudaMemsetAsync(h->mpos, 0, (size_t)h->B * 4, h->stream));
        h->mirror_cur = 0;
        int rc = mirror_refresh(h, nullptr, h->B);
        if (rc) return rc;
        CU(h, cudaStreamSynchronize(h->stream));
    }
    h->mirror_on = true;
    CU(h, enter_own(h));
    CU(h, cudaStreamSynchronize(h->stream));       // whatever a reset queued for the current buffer has landed
    *mirror = h->mirror_host[h->mirror_cur];
    if (row_floats) *row_floats = 2 * h->mir_len;
    return DBSGYM_OK;
}

// The host-mirror step in two halves: _begin copies the actions and launches (returns at once, the GPU works),
// _end waits and hands back reward / done / ring position.  dbsgym_step_host_mirror = _begin + _end.
int dbsgym_step_host_mirror_begin(DbsGymHandle* h, const float* actions) {
    int rc = check_ready(h, true);
    if (rc) return rc;
    if (!actions) return fail(h, DBSGYM_EINVAL, "null argument");
    if (!h->mirror_host[0]) return fail(h, DBSGYM_ESTATE, "no host mirror: call dbsgym_host_mirror first");
    if (h->mirror_pending) return fail(h, DBSGYM_ESTATE, "a host-mirror step is already in flight (call dbsgym_step_host_mirror_end)");
    const int B = h->B;
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->stream;
    CU(h, enter_own(h));
    h->mirror_on = true;
    if (h->fuse_tail) {
        // zero-copy control block: the kernel reads the actions from, and its tail writes reward / done / n_samples /
        // ring head to, pinned mapped host memory -- no memcpy nodes on the stream, one launch and one sync per step
        if (!h->ctl_host) {
            CU(h, cudaHostAlloc(reinterpret_cast<void**>(&h->ctl_host), (size_t)B * 17 + 64, cudaHostAllocMapped | cudaHostAllocPortable));
            CU(h, cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->ctl_dev), h->ctl_host, 0));
        }
        auto at = [&](unsigned char* base, int k) { return base + (size_t)k * B * 4; };
        memcpy(at(h->ctl_host, 0), actions, (size_t)B * 4);
        rc = step_impl(h, reinterpret_cast<const float*>(at(h->ctl_dev, 0)), nullptr, reinterpret_cast<float*>(at(h->ctl_dev, 1)),
                       at(h->ctl_dev, 4), s, nullptr, reinterpret_cast<int32_t*>(at(h->ctl_dev, 2)),
                       reinterpret_cast<int32_t*>(at(h->ctl_dev, 3)));
        if (rc) return rc;
    } else {
        CU(h, cudaMemcpyAsync(h->st_actions, actions, (size_t)B * 4, cudaMemcpyHostToDevice, s));
        rc = step_impl(h, h->st_actions, nullptr, h->st_reward, h->st_done, s, nullptr);
        if (rc) return rc;
        CU(h, cudaMemcpyAsync(h->pin_ints, h->n_samples, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
        CU(h, cudaMemcpyAsync(h->pin_ints + B, h->mpos, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    }
    h->mirror_pending = true;
    return DBSGYM_OK;
}

int dbsgym_step_host_mirror_end(DbsGymHandle* h, int32_t* pos, int32_t* n_new, float* reward, uint8_t* done) {
    if (!h || !pos || !n_new) return fail(h, DBSGYM_EINVAL, "null argument");
    if (!h->mirror_pending) return fail(h, DBSGYM_ESTATE, "no host-mirror step in flight (call dbsgym_step_host_mirror_begin)");
    const int B = h->B;
    CU(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = h->stream;
    h->mirror_pending = false;
    const int32_t *ns = nullptr, *hd = nullptr;
    if (h->fuse_tail) {
        auto at = [&](unsigned char* base, int k) { return base + (size_t)k * B * 4; };
        CU(h, cudaStreamSynchronize(s));
        if (reward) memcpy(reward, at(h->ctl_host, 1), (size_t)B * 4);
        if (done) memcpy(done, at(h->ctl_host, 4), (size_t)B);
        ns = reinterpret_cast<const int32_t*>(at(h->ctl_host, 2));
        hd = reinterpret_cast<const int32_t*>(at(h->ctl_host, 3));
    } else {
        if (reward) CU(h, cudaMemcpyAsync(reward, h->st_reward, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
        if (done) CU(h, cudaMemcpyAsync(done, h->st_done, (size_t)B, cudaMemcpyDeviceToHost, s));
        CU(h, cudaStreamSynchronize(s));
        ns = h->pin_ints; hd = h->pin_ints + B;
    }
    const int n = ns[0], hd0 = hd[0];
    bool uniform = true;
    for (int b = 1; b < B && uniform; ++b) uniform = ns[b] == n && hd[b] == hd0;
    *pos = (hd0 - h->W + h->mir_len) % h->mir_len;      // column of the oldest sample of the window
    *n_new = uniform ? n : -1;
    return DBSGYM_OK;
}

int dbsgym_step_host_mirror(DbsGymHandle* h, const float* actions, int32_t* pos, int32_t* n_new, float* reward,
                            uint8_t* done) {
    if (!pos || !n_new) return fail(h, DBSGYM_EINVAL, "null argument");
    int rc = dbsgym_step_host_mirror_begin(h, actions);
    if (rc) return rc;
    return dbsgym_step_host_mirror_end(h, pos, n_new, reward, done);
}

int dbsgym_get_obs_host(DbsGymHandle* h, float* obs) {
    int rc = check_ready(h, false);
    if (rc) return rc;
    if (!obs) return fail(h, DBSGYM_EINVAL, "obs is null");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, enter_own(h));
    {
        const bool mir = h->mirror_on;             // (a plain read-back: the mirror is not touched)
        h->mirror_on = false;
        cudaError_t eo = launch_obs(h, h->st_obs, nullptr, nullptr, 0, nullptr, h->B, h->stream);
        h->mirror_on = mir;
        CU(h, eo);
    }
    CU(h, cudaMemcpyAsync(obs, h->st_obs, (size_t)h->B * h->W * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return DBSGYM_OK;
}

int dbsgym_get_lfp(DbsGymHandle* h, double* lfp_true, double* lfp_rec, int32_t* n_samples) {
    if (!h) return DBSGYM_EINVAL;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    const size_t nb = (size_t)h->B * h->smax * 8;
    if (lfp_true) CU(h, cudaMemcpy(lfp_true, h->lfp_true, nb, cudaMemcpyDeviceToHost));
    if (lfp_rec) CU(h, cudaMemcpy(lfp_rec, h->lfp_rec, nb, cudaMemcpyDeviceToHost));
    if (n_samples) CU(h, cudaMemcpy(n_samples, h->n_samples, (size_t)h->B * 4, cudaMemcpyDeviceToHost));
    return DBSGYM_OK;
}

int dbsgym_get_rewards(DbsGymHandle* h, double* reward, double* u) {
    if (!h) return DBSGYM_EINVAL;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    if (reward) CU(h, cudaMemcpy(reward, h->reward, (size_t)h->B * 8, cudaMemcpyDeviceToHost));
    if (u) CU(h, cudaMemcpy(u, h->u, (size_t)h->B * 8, cudaMemcpyDeviceToHost));
    return DBSGYM_OK;
}

int dbsgym_get_phases(DbsGymHandle* h, const int32_t* env_ids, int32_t n, double* y) {
    if (!h || !y) return fail(h, DBSGYM_EINVAL, "null argument");
    if (n <= 0 || n > h->B) return fail(h, DBSGYM_EINVAL, "n=%d out of range", n);
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    const int32_t* ids = nullptr;
    int rc = upload_ids(h, env_ids, n, &ids);
    if (rc) return rc;
    std::vector<unsigned char> ph((size_t)n * h->Np * h->rb);
    rc = gather_from_device(h, ph.data(), h->phase, ids, n, (size_t)h->Np * h->rb);
    if (rc) return rc;
    if (!h->perm.empty()) {                          // octant order on the device (sector form of the low-rank operator, fp32)
        std::vector<int32_t> wd((size_t)n * h->Np);
        rc = gather_from_device(h, wd.data(), h->wind, ids, n, (size_t)h->Np * 4);
        if (rc) return rc;
        const float* p = reinterpret_cast<const float*>(ph.data());
        for (int r = 0; r < n; ++r)
            for (int d = 0; d < h->N; ++d)
                y[(size_t)r * h->N + h->perm[d]] = (double)p[(size_t)r * h->Np + d] + kTwoPi * (double)wd[(size_t)r * h->Np + d];
        return DBSGYM_OK;
    }
    if (h->f64) {
        const double* p = reinterpret_cast<const double*>(ph.data());
        for (int r = 0; r < n; ++r)
            for (int i = 0; i < h->N; ++i) y[(size_t)r * h->N + i] = p[(size_t)r * h->Np + i];
    } else {
        std::vector<int32_t> wd((size_t)n * h->Np);
        rc = gather_from_device(h, wd.data(), h->wind, ids, n, (size_t)h->Np * 4);
        if (rc) return rc;
        const float* p = reinterpret_cast<const float*>(ph.data());
        for (int r = 0; r < n; ++r)
            for (int i = 0; i < h->N; ++i)
                y[(size_t)r * h->N + i] = (double)p[(size_t)r * h->Np + i] + kTwoPi * (double)wd[(size_t)r * h->Np + i];
    }
    return DBSGYM_OK;
}

// ---- whole-handle snapshot ---------------------------------------------------------------------------
// Everything a later step depends on, as one opaque blob: phases + winding counts, per-oscillator vectors, observation
// rings + heads, running rfft bins, episode counters, the carried FSAL row, the last step's samples and the counters.
int dbsgym_state_bytes(const DbsGymHandle* h, uint64_t* bytes) {
    if (!h || !bytes) return DBSGYM_EINVAL;
    uint64_t total = sizeof(SnapshotHeader);
    for (const SnapPart& sp : snapshot_parts(const_cast<DbsGymHandle*>(h))) total += sp.bytes;
    *bytes = total;
    return DBSGYM_OK;
}

int dbsgym_get_state(DbsGymHandle* h, void* blob, uint64_t bytes) {
    if (!h || !blob) return fail(h, DBSGYM_EINVAL, "null argument");
    uint64_t need = 0;
    dbsgym_state_bytes(h, &need);
    if (bytes < need) return fail(h, DBSGYM_EINVAL, "snapshot buffer too small: %llu < %llu bytes",
                                  (unsigned long long)bytes, (unsigned long long)need);
    if (h->mirror_pending) return fail(h, DBSGYM_ESTATE, "a host-mirror step is in flight");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    unsigned char* out = static_cast<unsigned char*>(blob);
    SnapshotHeader hd = snapshot_header(h, need);
    memcpy(out, &hd, sizeof(hd));
    out += sizeof(hd);
    for (const SnapPart& sp : snapshot_parts(h)) {
        CU(h, cudaMemcpy(out, sp.dev, sp.bytes, cudaMemcpyDeviceToHost));
        out += sp.bytes;
    }
    return DBSGYM_OK;
}

int dbsgym_set_state(DbsGymHandle* h, const void* blob, uint64_t bytes) {
    if (!h || !blob) return fail(h, DBSGYM_EINVAL, "null argument");
    uint64_t need = 0;
    dbsgym_state_bytes(h, &need);
    if (bytes < need) return fail(h, DBSGYM_EINVAL, "snapshot is %llu bytes, this handle needs %llu",
                                  (unsigned long long)bytes, (unsigned long long)need);
    SnapshotHeader hd, want = snapshot_header(h, need);
    memcpy(&hd, blob, sizeof(hd));
    if (memcmp(&hd, &want, sizeof(hd)) != 0)
        return fail(h, DBSGYM_EINVAL, "snapshot does not belong to a handle of this shape (n_envs, n_osc, window, precision, ...)");
    if (h->mirror_pending) return fail(h, DBSGYM_ESTATE, "a host-mirror step is in flight");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    const unsigned char* in = static_cast<const unsigned char*>(blob) + sizeof(hd);
    for (const SnapPart& sp : snapshot_parts(h)) {
        CU(h, cudaMemcpy(sp.dev, in, sp.bytes, cudaMemcpyHostToDevice));
        in += sp.bytes;
    }
    if (h->mirror_on) {                             // the host mirror follows the restored rings
        int rc = mirror_refresh(h, nullptr, h->B);
        if (rc) return rc;
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return DBSGYM_OK;
}

int dbsgym_launch_count(DbsGymHandle* h, uint64_t* launches, int32_t reset) {
    if (!h || !launches) return DBSGYM_EINVAL;
    *launches = h->n_launches;
    if (reset) h->n_launches = 0;
    return DBSGYM_OK;
}

int dbsgym_get_window(DbsGymHandle* h, const int32_t* env_ids, int32_t n, double* window) {
    if (!h || !window) return fail(h, DBSGYM_EINVAL, "null argument");
    if (n <= 0 || n > h->B) return fail(h, DBSGYM_EINVAL, "n=%d out of range", n);
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    const int32_t* ids = nullptr;
    int rc = upload_ids(h, env_ids, n, &ids);
    if (rc) return rc;
    const int W = h->W;
    std::vector<unsigned char> rg((size_t)n * W * h->rb);
    std::vector<int32_t> hd((size_t)n);
    rc = gather_from_device(h, rg.data(), h->ring, ids, n, (size_t)W * h->rb);
    if (rc) return rc;
    rc = gather_from_device(h, hd.data(), h->head, ids, n, 4);
    if (rc) return rc;
    for (int r = 0; r < n; ++r)
        for (int m = 0; m < W; ++m) {
            int c = m - hd[r];
            if (c < 0) c += W;
            window[(size_t)r * W + c] = h->f64 ? reinterpret_cast<const double*>(rg.data())[(size_t)r * W + m]
                                               : (double)reinterpret_cast<const float*>(rg.data())[(size_t)r * W + m];
        }
    return DBSGYM_OK;
}

int dbsgym_set_window(DbsGymHandle* h, const int32_t* env_ids, int32_t n, const double* window) {
    if (!h || !window) return fail(h, DBSGYM_EINVAL, "null argument");
    if (n <= 0 || n > h->B) return fail(h, DBSGYM_EINVAL, "n=%d out of range", n);
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    const int32_t* ids = nullptr;
    int rc = upload_ids(h, env_ids, n, &ids);
    if (rc) return rc;
    std::vector<unsigned char> buf;
    if (h->f64) to_real_rows<double>(window, n, h->W, h->W, buf);
    else to_real_rows<float>(window, n, h->W, h->W, buf);
    rc = scatter_to_device(h, h->ring, buf.data(), ids, n, (size_t)h->W * h->rb);
    if (rc) return rc;
    std::vector<int32_t> z((size_t)n, 0);
    rc = scatter_to_device(h, h->head, z.data(), ids, n, 4);
    if (rc) return rc;
    CU(h, launch_spec_init(h, ids, n, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (h->mirror_on) {                             // keep the host mirror in sync with the overwritten rings
        rc = mirror_refresh(h, ids, n);
        if (rc) return rc;
        CU(h, cudaStreamSynchronize(h->stream));
    }
    return DBSGYM_OK;
}

int dbsgym_get_episode(DbsGymHandle* h, int32_t* step_idx, uint8_t* done) {
    if (!h) return DBSGYM_EINVAL;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    if (step_idx) CU(h, cudaMemcpy(step_idx, h->step_idx, (size_t)h->B * 4, cudaMemcpyDeviceToHost));
    if (done) CU(h, cudaMemcpy(done, h->done, (size_t)h->B, cudaMemcpyDeviceToHost));
    return DBSGYM_OK;
}

int dbsgym_counters(DbsGymHandle* h, uint64_t* accepted, uint64_t* rejected, uint64_t* rhs_evals, int32_t* status,
                    int32_t reset) {
    if (!h) return DBSGYM_EINVAL;
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    unsigned long long c[4];
    int32_t st = 0;
    CU(h, cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
    CU(h, cudaMemcpy(&st, h->status, 4, cudaMemcpyDeviceToHost));
    if (accepted) *accepted = c[0];
    if (rejected) *rejected = c[1];
    if (rhs_evals) *rhs_evals = c[2];
    if (status) *status = st;
    if (reset) {
        CU(h, cudaMemset(h->counters, 0, sizeof(c)));
        CU(h, cudaMemset(h->status, 0, 4));
    }
    return DBSGYM_OK;
}

int dbsgym_rhs_reused(DbsGymHandle* h, uint64_t* reused) {
    if (!h || !reused) return fail(h, DBSGYM_EINVAL, "null argument");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    unsigned long long v = 0;
    CU(h, cudaMemcpy(&v, h->counters + 3, sizeof(v), cudaMemcpyDeviceToHost));
    *reused = v;
    return DBSGYM_OK;
}

int dbsgym_set_timing(DbsGymHandle* h, int32_t enabled) {
    if (!h) return DBSGYM_EINVAL;
    h->timing = enabled != 0;
    return DBSGYM_OK;
}

int dbsgym_last_step_ms(DbsGymHandle* h, float* ms2) {
    if (!h || !ms2) return fail(h, DBSGYM_EINVAL, "null argument");
    if (!h->timing) return fail(h, DBSGYM_ESTATE, "timing is off (dbsgym_set_timing)");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaEventSynchronize(h->ev[2]));
    CU(h, cudaEventElapsedTime(&ms2[0], h->ev[0], h->ev[1]));
    CU(h, cudaEventElapsedTime(&ms2[1], h->ev[1], h->ev[2]));
    return DBSGYM_OK;
}

int dbsgym_trace_begin(DbsGymHandle* h, int32_t capacity) {
    if (!h) return DBSGYM_EINVAL;
    if (capacity <= 0) return fail(h, DBSGYM_EINVAL, "trace capacity must be positive");
    if (!h->fuse_tail) return fail(h, DBSGYM_ESTATE, "the evaluation trace needs a beta-power reward (fused observation tail)");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, enter_own(h));
    CU(h, cudaStreamSynchronize(h->stream));
    if (capacity != h->trace_cap) {                 // the row stride of the trace IS the capacity (header: [n_envs][capacity])
        if (h->trace) cudaFree(h->trace);
        h->trace = nullptr; h->trace_cap = 0;
        CU(h, cudaMalloc(&h->trace, (size_t)h->B * capacity * sizeof(double)));
        h->trace_cap = capacity;
    }
    if (!h->trace_len) CU(h, cudaMalloc(&h->trace_len, (size_t)h->B * 4));
    CU(h, cudaMemset(h->trace_len, 0, (size_t)h->B * 4));
    h->trace_on = true;
    return DBSGYM_OK;
}

int dbsgym_trace_end(DbsGymHandle* h) {
    if (!h) return DBSGYM_EINVAL;
    h->trace_on = false;
    return DBSGYM_OK;
}

int dbsgym_trace_get(DbsGymHandle* h, double* trace, int32_t* len) {
    if (!h) return DBSGYM_EINVAL;
    if (!h->trace) return fail(h, DBSGYM_ESTATE, "no trace recorded (dbsgym_trace_begin)");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    if (trace) CU(h, cudaMemcpy(trace, h->trace, (size_t)h->B * h->trace_cap * sizeof(double), cudaMemcpyDeviceToHost));
    if (len) CU(h, cudaMemcpy(len, h->trace_len, (size_t)h->B * 4, cudaMemcpyDeviceToHost));
    return DBSGYM_OK;
}

int dbsgym_eval_bbpow(DbsGymHandle* h, const DbsGymEvalSpec* spec, const double* weights, double* bbpow) {
    if (!h || !spec || !weights || !bbpow) return fail(h, DBSGYM_EINVAL, "null argument");
    if (spec->struct_bytes != sizeof(DbsGymEvalSpec)) return fail(h, DBSGYM_EINVAL, "DbsGymEvalSpec size mismatch");
    if (!h->trace) return fail(h, DBSGYM_ESTATE, "no trace recorded (dbsgym_trace_begin)");
    if (spec->n_k <= 0 || spec->k_lo < 0 || spec->padlen < 0) return fail(h, DBSGYM_EINVAL, "bad bin range / padlen");
    if (spec->a[0] != 1.0) return fail(h, DBSGYM_EINVAL, "filter must be normalised (a[0] == 1)");
    CU(h, cudaSetDevice(h->cfg.device));
    CU(h, cudaDeviceSynchronize());
    std::vector<int32_t> len((size_t)h->B);
    CU(h, cudaMemcpy(len.data(), h->trace_len, (size_t)h->B * 4, cudaMemcpyDeviceToHost));
    const int n = len[0];
    for (int b = 1; b < h->B; ++b)
        if (len[b] != n) return fail(h, DBSGYM_ESTATE, "traces have different lengths (%d vs %d): evaluate in lockstep", len[b], n);
    if (n < spec->padlen + 2) return fail(h, DBSGYM_ESTATE, "trace of %d samples is shorter than the filter padding", n);
    if (spec->k_lo + spec->n_k - 1 > n / 2) return fail(h, DBSGYM_EINVAL, "bin range exceeds the rfft length");
    EvalParams p;
    p.trace = h->trace; p.cap = h->trace_cap; p.n = n; p.B = h->B; p.pad = spec->padlen;
    for (int i = 0; i < 5; ++i) { p.b[i] = spec->b[i]; p.a[i] = spec->a[i]; }
    for (int i = 0; i < 4; ++i) p.zi[i] = spec->zi[i];
    p.k_lo = spec->k_lo; p.n_k = spec->n_k;
    double *scratch = nullptr, *w = nullptr, *out = nullptr;
    cudaError_t e = cudaMalloc(&scratch, (size_t)h->B * (n + 2 * spec->padlen) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&w, (size_t)spec->n_k * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&out, (size_t)h->B * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(w, weights, (size_t)spec->n_k * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        p.scratch = scratch; p.weights = w; p.out = out;
        eval_filtfilt_kernel<<<(h->B + 31) / 32, 32, 0, h->stream>>>(p);
        eval_band_power_kernel<<<h->B, kEvalThreads, 0, h->stream>>>(p);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(bbpow, out, (size_t)h->B * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(scratch); cudaFree(w); cudaFree(out);
    if (e != cudaSuccess) return fail(h, DBSGYM_ECUDA, "evaluation kernels failed: %s", cudaGetErrorString(e));
    return DBSGYM_OK;
}

int dbsgym_measure_fp32_peak(int32_t device, double ms_target, double* tflops) {
    // best of the scalar FFMA chain and the packed FFMA2 chain
    if (!tflops) return DBSGYM_EINVAL;
    double a = 0.0, b = 0.0;
    int rc = dbsgym_measure_fp32_peak_mode(device, ms_target, 0, &a);
    if (rc) return rc;
    rc = dbsgym_measure_fp32_peak_mode(device, ms_target, 1, &b);
    if (rc) return rc;
    *tflops = a > b ? a : b;
    return DBSGYM_OK;
}

int dbsgym_measure_mufu_peak(int32_t device, double ms_target, double* tops) {
    return dbsgym_measure_fp32_peak_mode(device, ms_target, 2, tops);
}

int dbsgym_measure_fp32_peak_mode(int32_t device, double ms_target, int32_t packed, double* tflops) {
    if (!tflops) return DBSGYM_EINVAL;
    DbsGymHandle* h = nullptr;
    CU(h, cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(h, cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 4, threads = 512;
    float* out = nullptr;
    CU(h, cudaMalloc(&out, 4));
    cudaEvent_t a, b;
    CU(h, cudaEventCreate(&a));
    CU(h, cudaEventCreate(&b));
    int iters = 1 << 14;
    double best = 0.0;
    for (int rep = 0; rep < 8; ++rep) {
        cudaEventRecord(a);
        if (packed == 2) mufu_peak_kernel<<<blocks, threads>>>(out, iters, 1e-3f);
        else if (packed) fma2_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
        else fma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(b);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) { cudaFree(out); return fail(h, DBSGYM_ECUDA, "peak kernel: %s", cudaGetErrorString(e)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        // packed == 2: special-function operations (one MUFU each; the FADD between them runs on another pipe)
        const double fl = (packed == 2 ? 2.0 : packed ? 4.0 : 2.0) * kChains * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
        if (ms < ms_target && iters < (1 << 24)) iters *= 2;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    *tflops = best;
    return DBSGYM_OK;
}

}  // extern "C"
