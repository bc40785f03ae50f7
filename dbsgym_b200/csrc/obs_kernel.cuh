// Observation / reward kernels (sm_100a), HBM-bound.
//
// For the beta-power rewards the per-step work (window append, reward, episode bookkeeping) is fused into the step
// kernel's tail (step_kernel.cuh: obs_tail, with incrementally updated rfft bins); this file holds its companions
// (spec_init_kernel, obs_copy_kernel) and the stand-alone observation kernel that still serves the R2 reward, window
// emission at reset and the A/B switch DBSGYM_NO_FUSED_OBS:
//
// Per environment: append the step's recorded LFP samples to the observation ring
// (reference environment/env.py:447-448), emit the window in chronological order as the
// float32 observation (env.py:454), and evaluate the reward of env.py:638-688:
//   R1/R3  beta-band power  sum_k 2 |X_k / W|^2 over the rfft bins inside (12.5, 21) Hz
//          (utils.py:21-27).  |X_k| is invariant under a circular shift when the DFT length
//          equals the ring length, so the bins are accumulated in ring STORAGE order.  Sample
//          m = 128 i + t is handled by thread t: it first sums x[128 i + t] e^{-j w_k 128 i} over i
//          (a [bins][iters] twiddle table in shared memory, arithmetic in the ring precision) and
//          then multiplies by its own seed e^{-j w_k t} in float64 -- 42 instead of 114 FMAs per
//          thread and bin compared with a rotation recurrence, and no sincos.
//   R2     -1e3 (x_f[-1] - mean x_f)^2 with x_f = filtfilt(butter(2,[12,30] Hz)) (env.py:653-666,
//          utils.py:794-816).  filtfilt (odd padding + lfilter_zi start) is linear in the
//          window, so the bracket is a dot product g . window with g precomputed on the host.
// One CTA per environment; the window passes through shared memory once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dbsgym {

constexpr int kObsThreads = 128;
constexpr int kMaxBins = 32;

struct ObsParams {
    int B, W, smax;
    void* ring; int32_t* head;
    const double* lfp_rec; const int32_t* n_samples;
    float* obs; float* samples_f; float* mirror; float* reward_f; double* reward; uint8_t* done_out; uint8_t* done_dev;
    int mir_len; int32_t* mpos;     // host mirror log: row = 2 * mir_len floats, per-environment write column (dbsgym.h)
    int32_t* step_idx; const int32_t* episode_len; const double* u;
    int kind, nbins;
    double power_scale, action_cost, threshold, threshold_penalty, temp_scale;
    const double* lin_g;      // [W] chronological
    const double* tw_seed;    // [nbins][kObsThreads][2]  cos, sin of 2*pi*k*t/W for t < kObsThreads
    const void* tw_inner;     // [nbins][iters][2] (real) cos, sin of 2*pi*k*(kObsThreads*i)/W
    int iters;                // ceil(W / kObsThreads)
    int append;               // 1: full step; 0: only emit the observation of the current window
    const int32_t* env_ids; int n_launch;
};

template <int CNT>
__device__ __forceinline__ void loadv(const float* __restrict__ p, float* o) {
#pragma unroll
    for (int q = 0; q < CNT / 4; ++q) { const float4 v = reinterpret_cast<const float4*>(p)[q]; o[4*q] = v.x; o[4*q+1] = v.y; o[4*q+2] = v.z; o[4*q+3] = v.w; }
}
template <int CNT>
__device__ __forceinline__ void loadv(const double* __restrict__ p, double* o) {
#pragma unroll
    for (int q = 0; q < CNT / 2; ++q) { const double2 v = reinterpret_cast<const double2*>(p)[q]; o[2*q] = v.x; o[2*q+1] = v.y; }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename real>
__global__ void __launch_bounds__(kObsThreads) obs_kernel(const ObsParams p) {
    extern __shared__ __align__(16) unsigned char obs_smem[];
    real* xs = reinterpret_cast<real*>(obs_smem);                       // [W] window, ring storage order
    real* twi = xs + ((p.W + 3) & ~3);                                  // [nbins][iters][2]
    __shared__ double part[kMaxBins + 1][kObsThreads / 32][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = blockIdx.x;
    if (slot >= p.n_launch) return;
    const int env = p.env_ids ? p.env_ids[slot] : slot;
    const int W = p.W;
    real* ring = reinterpret_cast<real*>(p.ring) + (size_t)env * W;
    const int head = p.head[env];
    const int S = p.append ? p.n_samples[env] : 0;

    for (int m = tid; m < W; m += kObsThreads) xs[m] = ring[m];
    if (p.append && p.kind != 1) {
        const real* src = reinterpret_cast<const real*>(p.tw_inner);
        for (int i = tid; i < p.nbins * p.iters * 2; i += kObsThreads) twi[i] = src[i];
    }
    __syncthreads();
    for (int i = tid; i < S; i += kObsThreads) {
        int pos = head + i;
        if (pos >= W) pos -= W;
        const real v = real(p.lfp_rec[(size_t)env * p.smax + i]);
        ring[pos] = v;
        xs[pos] = v;
        if (p.samples_f) p.samples_f[(size_t)env * p.smax + i] = (float)v;
        if (p.mirror) {                               // zero-copy store into the pinned host mirror (both copies)
            float* mr = p.mirror + (size_t)env * 2 * p.mir_len;
            int c = p.mpos[env] + i;
            if (c >= p.mir_len) c -= p.mir_len;
            mr[c] = (float)v;
            mr[c + p.mir_len] = (float)v;
        }
    }
    int new_head = head + S;
    if (new_head >= W) new_head -= W;
    __syncthreads();

    if (p.obs) {
        float* o = p.obs + (size_t)env * W;
        for (int m = tid; m < W; m += kObsThreads) {
            int n = m - new_head;
            if (n < 0) n += W;
            o[n] = (float)xs[m];
        }
    }
    if (!p.append) {
        if (p.mirror) {                               // reset / refresh: the whole window, ending at the write column
            float* mr = p.mirror + (size_t)env * 2 * p.mir_len;
            const int mp = p.mpos[env];
            for (int m = tid; m < W; m += kObsThreads) {
                int n = m - new_head;                 // chronological index of ring slot m
                if (n < 0) n += W;
                int c = mp - W + n;
                if (c < 0) c += p.mir_len;
                const float v = (float)xs[m];
                mr[c] = v; mr[c + p.mir_len] = v;
            }
        }
        return;
    }

    if (p.kind == 1) {                                // R2: g . window (chronological order)
        double acc = 0.0;
        for (int m = tid; m < W; m += kObsThreads) {
            int n = m - new_head;
            if (n < 0) n += W;
            acc = fma(p.lin_g[n], (double)xs[m], acc);
        }
        acc = warp_sum_d(acc);
        if (lane == 0) part[kMaxBins][warp][0] = acc;
    } else {
        // this thread's samples x[128 i + tid], i < iters, kept in registers across the bins
        constexpr int kMaxIters = 24;                 // W <= 3072 in registers; longer windows re-read shared memory
        real xr[kMaxIters];
        const bool in_regs = p.iters <= kMaxIters;
        if (in_regs) {
#pragma unroll
            for (int i = 0; i < kMaxIters; ++i) {
                const int m = i * kObsThreads + tid;
                xr[i] = (i < p.iters && m < W) ? xs[m] : real(0);
            }
        }
        for (int kb = 0; kb < p.nbins; ++kb) {
            const real* tw = twi + (size_t)kb * p.iters * 2;
            real re = real(0), im = real(0);
            if (in_regs) {
                // the table row is read with 128-bit loads (broadcast): two iterations per LDS for float
                constexpr int kPer = 16 / (2 * sizeof(real));       // (cos,sin) pairs per 16 bytes
#pragma unroll
                for (int i = 0; i < kMaxIters; i += kPer) {
                    if (i < p.iters) {
                        real t[2 * kPer];
                        loadv<2 * kPer>(tw + 2 * i, t);
#pragma unroll
                        for (int e = 0; e < kPer; ++e) { re += xr[i + e] * t[2 * e]; im += xr[i + e] * t[2 * e + 1]; }
                    }
                }
            } else {
                for (int i = 0; i < p.iters; ++i) {
                    const int m = i * kObsThreads + tid;
                    const real x = m < W ? xs[m] : real(0);
                    re += x * tw[2 * i]; im += x * tw[2 * i + 1];
                }
            }
            // (re + j im) * (cs + j sn): the seed twiddle of this thread's offset, in float64
            const double cs = p.tw_seed[(kb * kObsThreads + tid) * 2];
            const double sn = p.tw_seed[(kb * kObsThreads + tid) * 2 + 1];
            double fr = (double)re * cs - (double)im * sn;
            double fi = (double)re * sn + (double)im * cs;
            fr = warp_sum_d(fr);
            fi = warp_sum_d(fi);
            if (lane == 0) { part[kb][warp][0] = fr; part[kb][warp][1] = fi; }
        }
    }
    __syncthreads();
    if (tid == 0) {
        const double au = fabs(p.u[env]);
        double r;
        if (p.kind == 1) {
            double d = 0.0;
            for (int w = 0; w < kObsThreads / 32; ++w) d += part[kMaxBins][w][0];
            r = -p.temp_scale * d * d - p.action_cost * au;
        } else {
            double pw = 0.0;
            for (int kb = 0; kb < p.nbins; ++kb) {
                double re = 0.0, im = 0.0;
                for (int w = 0; w < kObsThreads / 32; ++w) { re += part[kb][w][0]; im += part[kb][w][1]; }
                re /= (double)W; im /= (double)W;
                pw += (re * re + im * im) * 2.0;
            }
            if (p.kind == 0) r = -p.power_scale * pw - p.action_cost * au;
            else r = -((p.power_scale * pw > p.threshold) ? p.threshold_penalty : 0.0) - p.action_cost * au;
        }
        p.reward[env] = r;
        if (p.reward_f) p.reward_f[env] = (float)r;
        const int k = p.step_idx[env] + 1;
        p.step_idx[env] = k;
        const uint8_t dn = k >= p.episode_len[env] ? 1 : 0;
        p.done_dev[env] = dn;
        if (p.done_out) p.done_out[env] = dn;
        p.head[env] = new_head;
        if (p.mirror) {
            int nm = p.mpos[env] + S;
            if (nm >= p.mir_len) nm -= p.mir_len;
            p.mpos[env] = nm;
        }
    }
}

// ---- streamlined variant for the common case (append, beta-power reward, W <= 20 * 128) ----------------
// No shared-memory staging of the window: the new samples are written to the ring first, then every
// thread loads its 20 strided samples straight into registers (coalesced), writes the chronological
// observation from them, runs the 2-level DFT with compile-time trip counts and hands its (re, im)
// partials to a shared-memory tree (20 stores + 32 loads per thread instead of 20 float64 shuffle
// butterflies).  ~3 k instead of ~9 k warp instructions per environment.
constexpr int kFastIters = 20;
constexpr int kFastMaxBins = 16;
constexpr int kRedPitch = kObsThreads + 4;

template <typename real>
__global__ void __launch_bounds__(kObsThreads) obs_kernel_fast(const ObsParams p) {
    extern __shared__ __align__(16) unsigned char obs_smem[];
    real* twi = reinterpret_cast<real*>(obs_smem);                       // [nbins][kFastIters][2]
    real* red = twi + p.nbins * kFastIters * 2;                          // [2 * nbins][kRedPitch]
    __shared__ double fin[2 * kFastMaxBins];
    const int tid = threadIdx.x, lane = tid & 31;
    const int slot = blockIdx.x;
    if (slot >= p.n_launch) return;
    const int env = p.env_ids ? p.env_ids[slot] : slot;
    const int W = p.W;
    real* ring = reinterpret_cast<real*>(p.ring) + (size_t)env * W;
    const int head = p.head[env];
    const int S = p.n_samples[env];

    if (tid < S) {                                   // env.py:447: append the step's samples (oldest are overwritten)
        int pos = head + tid;
        if (pos >= W) pos -= W;
        const real v = real(p.lfp_rec[(size_t)env * p.smax + tid]);
        ring[pos] = v;
        if (p.samples_f) p.samples_f[(size_t)env * p.smax + tid] = (float)v;
        if (p.mirror) {
            float* mr = p.mirror + (size_t)env * 2 * p.mir_len;
            int c = p.mpos[env] + tid;
            if (c >= p.mir_len) c -= p.mir_len;
            mr[c] = (float)v;
            mr[c + p.mir_len] = (float)v;
        }
    }
    {
        const real* src = reinterpret_cast<const real*>(p.tw_inner);
        for (int i = tid; i < p.nbins * kFastIters * 2; i += kObsThreads) twi[i] = src[i];
    }
    __syncthreads();                                 // ring writes of this block are visible to its loads below
    int new_head = head + S;
    if (new_head >= W) new_head -= W;

    real x[kFastIters];
#pragma unroll
    for (int i = 0; i < kFastIters; ++i) {
        const int m = i * kObsThreads + tid;
        x[i] = m < W ? ring[m] : real(0);
    }
    if (p.obs) {
        float* o = p.obs + (size_t)env * W;
#pragma unroll
        for (int i = 0; i < kFastIters; ++i) {
            const int m = i * kObsThreads + tid;
            if (m < W) {
                int n = m - new_head;
                if (n < 0) n += W;
                o[n] = (float)x[i];
            }
        }
    }
    constexpr int kPer = 16 / (2 * sizeof(real));
    for (int kb = 0; kb < p.nbins; ++kb) {
        const real* tw = twi + kb * kFastIters * 2;
        real re = real(0), im = real(0);
#pragma unroll
        for (int i = 0; i < kFastIters; i += kPer) {
            real t[2 * kPer];
            loadv<2 * kPer>(tw + 2 * i, t);
#pragma unroll
            for (int e = 0; e < kPer; ++e) { re += x[i + e] * t[2 * e]; im += x[i + e] * t[2 * e + 1]; }
        }
        const real cs = real(p.tw_seed[(kb * kObsThreads + tid) * 2]);
        const real sn = real(p.tw_seed[(kb * kObsThreads + tid) * 2 + 1]);
        red[(2 * kb) * kRedPitch + tid] = re * cs - im * sn;
        red[(2 * kb + 1) * kRedPitch + tid] = re * sn + im * cs;
    }
    __syncthreads();
    {
        const int v = tid >> 2, sub = tid & 3;       // 4 threads per value, 32 entries each (skewed against bank conflicts)
        double acc = 0.0;
        if (v < 2 * p.nbins) {
            const real* row = red + v * kRedPitch + sub * 32;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) acc += (double)row[(k + lane) & 31];
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (sub == 0 && v < 2 * p.nbins) fin[v] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        const double au = fabs(p.u[env]);
        double pw = 0.0;
        for (int kb = 0; kb < p.nbins; ++kb) {
            const double re = fin[2 * kb] / (double)W, im = fin[2 * kb + 1] / (double)W;
            pw += (re * re + im * im) * 2.0;
        }
        double r;
        if (p.kind == 0) r = -p.power_scale * pw - p.action_cost * au;
        else r = -((p.power_scale * pw > p.threshold) ? p.threshold_penalty : 0.0) - p.action_cost * au;
        p.reward[env] = r;
        if (p.reward_f) p.reward_f[env] = (float)r;
        const int k = p.step_idx[env] + 1;
        p.step_idx[env] = k;
        const uint8_t dn = k >= p.episode_len[env] ? 1 : 0;
        p.done_dev[env] = dn;
        if (p.done_out) p.done_out[env] = dn;
        p.head[env] = new_head;
        if (p.mirror) {
            int nm = p.mpos[env] + S;
            if (nm >= p.mir_len) nm -= p.mir_len;
            p.mpos[env] = nm;
        }
    }
}

// ---- companions of the fused observation tail (step_kernel.cuh: obs_tail) -------------------------------
// spec_init_kernel: the running rfft bins of an environment, recomputed from its whole ring in float64
// (after a reset transient, dbsgym_set_window or dbsgym_set_reward).  X_k = sum_m ring[m] e^{j 2 pi k m / W}
// in ring storage order; one CTA per environment.
template <typename real>
__global__ void __launch_bounds__(kObsThreads) spec_init_kernel(const void* ring_, const double* tw_full, double* spec,
                                                                int W, int nbins, int spec_pitch,
                                                                const int32_t* env_ids, int n_launch) {
    __shared__ double part[kObsThreads / 32][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = blockIdx.x;
    if (slot >= n_launch) return;
    const int env = env_ids ? env_ids[slot] : slot;
    const real* ring = reinterpret_cast<const real*>(ring_) + (size_t)env * W;
    const double2* tw = reinterpret_cast<const double2*>(tw_full);
    for (int kb = 0; kb < nbins; ++kb) {
        double re = 0.0, im = 0.0;
        for (int m = tid; m < W; m += kObsThreads) {
            const double x = (double)ring[m];
            const double2 w = tw[(size_t)m * nbins + kb];
            re = fma(x, w.x, re);
            im = fma(x, w.y, im);
        }
        re = warp_sum_d(re); im = warp_sum_d(im);
        if (lane == 0) { part[warp][0] = re; part[warp][1] = im; }
        __syncthreads();
        if (tid == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < kObsThreads / 32; ++w) { a += part[w][0]; b += part[w][1]; }
            spec[((size_t)env * spec_pitch + kb) * 2] = a;
            spec[((size_t)env * spec_pitch + kb) * 2 + 1] = b;
        }
        __syncthreads();
    }
}

// obs_copy_kernel: the chronological float32 observation (env.py:454) of every environment from its ring,
// obs[b][n] = ring[b][(head_b + n) mod W].  Pure HBM streaming: each thread has kCopyPer independent loads
// in flight; reads are coalesced and writes are coalesced up to the rotation.
constexpr int kCopyThreads = 256;
constexpr int kCopyPer = 10;       // 256 * 10 >= W = 2340 in one pass
template <typename real>
__global__ void __launch_bounds__(kCopyThreads) obs_copy_kernel(const void* ring_, const int32_t* head, float* obs, int W, int B) {
    const int env = blockIdx.x;
    if (env >= B) return;
    const real* ring = reinterpret_cast<const real*>(ring_) + (size_t)env * W;
    float* o = obs + (size_t)env * W;
    const int hd = head[env];
    for (int base = 0; base < W; base += kCopyThreads * kCopyPer) {
        real x[kCopyPer];
#pragma unroll
        for (int i = 0; i < kCopyPer; ++i) {
            const int n = base + i * kCopyThreads + threadIdx.x;      // chronological index
            int m = n + hd;
            if (m >= W) m -= W;
            x[i] = n < W ? ring[m] : real(0);
        }
#pragma unroll
        for (int i = 0; i < kCopyPer; ++i) {
            const int n = base + i * kCopyThreads + threadIdx.x;
            if (n < W) o[n] = (float)x[i];
        }
    }
}

inline size_t obs_fast_smem_bytes(int nbins, size_t real_bytes) {
    return (size_t)(nbins * kFastIters * 2 + 2 * nbins * kRedPitch) * real_bytes;
}

inline size_t obs_smem_bytes(int W, int nbins, int iters, size_t real_bytes) {
    return (size_t)(((W + 3) & ~3) + nbins * iters * 2) * real_bytes;
}

}  // namespace dbsgym
