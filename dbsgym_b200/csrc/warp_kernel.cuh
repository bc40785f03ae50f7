// Spectral step kernel, ONE WARP PER ENVIRONMENT, octant ownership (sm_100a; float32, the shipped 8 x 8 x 8 grid).
//
// Same mathematics as step_kernel<CPL_SPECTRAL> (step_kernel.cuh): adaptive Dopri5 + I-controller + dense output as the
// reference calls diffrax (environment/env.py:247-271), the coupling sum of env.py:252-256 through the generalised
// mean-field identity over the eigenmodes of alpha (geometry.py: spectral_factors), LFP samples (env.py:396-412) and the
// fused observation tail (env.py:447-454, :638-650, :669-688).  What changes is the mapping onto the machine:
//
//   * a lane owns TWO points of the fundamental octant (zq, xq, yq = 2 yp + {0, 1}) and ALL 8 mirror images of each:
//     16 oscillators.  The three reflections of the grid therefore act inside a thread -- the parity-sector transform
//     is an 8-point Walsh-Hadamard butterfly on registers (12 packed add / sub per point and direction) and needs NO
//     shuffle (the 64-thread kernels spend 128 SHFL + 128 FMA-pipe slots per thread and RHS evaluation on it);
//   * every sector uses exactly its own number of modes (compile-time rank list, e.g. 7 + 4 x 6 + 1 = 32): projection and
//     expansion are 2 FFMA2 per mode and lane each; the only cross-lane step is the sum of the 32 lane partials of every
//     mode, done through shared memory by all 32 lanes in parallel (one half row of 16 partials each per round);
//   * one environment = one warp: every reduction (error norm, LFP samples, mode sums) stays inside the warp --
//     __syncwarp instead of named barriers -- and the control code (step-size controller, segment programme) runs once
//     per environment.  16 independent chains per lane hide FFMA / MUFU / LDS latency from within the warp, so 8-10
//     resident warps per SM are enough.
//
// Lane l: zq = l >> 3, xq = (l >> 1) & 3, yp = l & 1.  Register r = 8 p + g: point p (yq = 2 yp + p), image
// g = 4 my + 2 mz + mx (bit set = mirrored coordinate); after the butterfly index g is the sector
// s = 4 [odd in y] + 2 [odd in z] + [odd in x] of geometry.sector_blocks.
#pragma once
#include <utility>
#include "step_kernel.cuh"

// mode sums in one pass (all partials in shared memory at once) or in two (half the buffer, two more __syncwarp)
#ifndef DBSGYM_WARP_PASSES
#define DBSGYM_WARP_PASSES 1
#endif
// eigenvector entries: 0 = registers of the lane (2 x modes of them), 1 = one table per CTA in shared memory (LDS.128 per
// mode pair in projection and expansion; frees 64 registers per thread).  Measured on B200 at 8 warps per SM: the table
// costs 11 % (0.281 -> 0.311 ms per 4096-env step), and the register file allocates warps in fours, so the next
// occupancy step after 8 warps x 255 registers is 12 x 168, which the 12 KB of stage derivatives per warp do not allow.
#ifndef DBSGYM_WARP_VSMEM
#define DBSGYM_WARP_VSMEM 0
#endif
#ifndef DBSGYM_WARP_MAXNREG
#define DBSGYM_WARP_MAXNREG 255      // (step_f32_warp.cu derives it from the warps per CTA)
#endif
#ifndef DBSGYM_WARP_MIRROR_STORES
#define DBSGYM_WARP_MIRROR_STORES 1  // (0: A/B builds that measure what the zero-copy sample stores cost; the host log is then wrong)
#endif
// A/B switches of the per-environment prologue / epilogue
#ifndef DBSGYM_WARP_PREFETCH
#define DBSGYM_WARP_PREFETCH 1       // pull the next environment's rows into L2 while this one is integrated
#endif
#ifndef DBSGYM_WARP_HOIST
#define DBSGYM_WARP_HOIST 0          // fetch w0 / stim together with the state instead of at the first segment (measured: 0.245 -> 0.271 ms, register pressure)
#endif
#ifndef DBSGYM_WARP_TAIL2
#define DBSGYM_WARP_TAIL2 1          // warp-kernel observation tail (samples from shared memory, spread twiddle loop)
#endif

namespace dbsgym {

constexpr int kWR = 16;            // oscillators per lane
template <int V> struct IC { static constexpr int value = V; };
constexpr int kWarpTs = 32;        // save times of a step() segment staged in shared memory (longer lists are read from global)
template <int... Rs> struct RankSet {
    static_assert(sizeof...(Rs) == 8, "one rank per parity sector");
    __host__ __device__ static constexpr int get(int s) { const int r[8] = {Rs...}; return r[s]; }
    __host__ __device__ static constexpr int off(int s) {             // first mode of sector s in the lane's eigenvector registers (off(8) = all modes)
        const int r[8] = {Rs...};
        int o = 0;
        for (int i = 0; i < s; ++i) o += r[i];
        return o;
    }
    __host__ __device__ static constexpr int sector_of(int m) {       // the sector mode m belongs to
        for (int s = 0; s < 8; ++s)
            if (m < off(s + 1)) return s;
        return 7;
    }
    __host__ __device__ static constexpr bool first_of_sector(int m) { return off(sector_of(m)) == m; }
};

template <class F, int... I> __device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) { (f(IC<I>{}), ...); }
template <int N, class F> __device__ __forceinline__ void static_for(F&& f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

template <class RK> struct WarpLayout {
    static constexpr int NM = RK::off(8);                       // modes (even: they are handled in pairs)
    static constexpr int NP = NM / 2;                           // mode pairs
    static constexpr int PASSES = DBSGYM_WARP_PASSES;
    static constexpr int PP = (NP + PASSES - 1) / PASSES;       // mode pairs per pass
    static constexpr int RS = 36;                               // words per half row: 16 float2 partials + 4 (conflict-free both ways)
    static constexpr int HR = 4 * PP;                           // half rows of a pass: (mode, lanes 0-15 / 16-31)
    static constexpr int ROUNDS = (HR + 31) / 32;
    static constexpr int p_floats = HR * RS;
    static constexpr int c_floats = 2 * NM;
    static constexpr bool VSMEM = DBSGYM_WARP_VSMEM != 0;
    static constexpr size_t v_bytes = VSMEM ? (size_t)NP * 32 * 16 : 0;      // eigenvector table of the CTA
    static_assert(NM % 2 == 0 && c_floats % 4 == 0, "modes come in pairs");
    static_assert(p_floats >= 8 * 68, "the partials buffer also holds the lane sums of 8 LFP samples");
    // bytes of one warp's shared memory: K slots, partials + coefficients, winding counts, w0 + pulse, tail scratch
    static constexpr size_t bytes = (size_t)(kSlots * 512 + p_floats + c_floats + 512 + 512 + 32) * 4 + (32 + 2 * kWarpTs) * 8 + 36 * 4 + 8;
    static constexpr size_t bytes_aligned = (bytes + 15) & ~(size_t)15;
    static constexpr size_t cta_bytes(int warps) { return v_bytes + (size_t)warps * bytes_aligned; }
};

// thread-private rows of 16 floats in shared memory: piece q (4 floats) of lane l at float4 index q * 32 + l
__device__ __forceinline__ void wload16(const float* __restrict__ row, int lane, float (&o)[kWR]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) unpack(reinterpret_cast<const float4*>(row)[q * 32 + lane], o + 4 * q);
}
__device__ __forceinline__ void wstore16(float* __restrict__ row, int lane, const float (&o)[kWR]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(row)[q * 32 + lane] = pack4(o + 4 * q);
}

// the lane's 16 entries of a natural-order [512] vector (index (z * 8 + x) * 8 + y): 8 accesses of two adjacent y
struct OctLane {
    int line[4];                   // first oscillator of the line (z, x) of image c = 2 mz + mx
    int ylo, yhi;                  // y = 2 yp (points 0, 1 unmirrored), y = 6 - 2 yp (point 1, point 0 mirrored)
    __device__ __forceinline__ explicit OctLane(int lane) {
        const int zq = lane >> 3, xq = (lane >> 1) & 3, yp = lane & 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int z = (c & 2) ? 7 - zq : zq, x = (c & 1) ? 7 - xq : xq;
            line[c] = (z * 8 + x) * 8;
        }
        ylo = 2 * yp; yhi = 6 - 2 * yp;
    }
};
template <typename T, typename T2>
__device__ __forceinline__ void oct_load(const T* __restrict__ g, const OctLane& L, T (&o)[kWR]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const T2 a = *reinterpret_cast<const T2*>(g + L.line[c] + L.ylo);
        const T2 b = *reinterpret_cast<const T2*>(g + L.line[c] + L.yhi);
        o[c] = a.x; o[8 + c] = a.y;
        o[4 + c] = b.y; o[12 + c] = b.x;
    }
}
template <typename T, typename T2>
__device__ __forceinline__ void oct_store(T* __restrict__ g, const OctLane& L, const T (&o)[kWR]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        T2 a, b;
        a.x = o[c]; a.y = o[8 + c];
        b.y = o[4 + c]; b.x = o[12 + c];
        *reinterpret_cast<T2*>(g + L.line[c] + L.ylo) = a;
        *reinterpret_cast<T2*>(g + L.line[c] + L.yhi) = b;
    }
}

// y0 + sum_j (dt coef[j]) K[slot[j]]: the stage argument directly (dt folded into the tableau row once per warp)
template <int NJ>
__device__ __forceinline__ void wlincomb(const float* __restrict__ Kb, int lane, const int (&slot)[NJ], const double (&coef)[NJ],
                                         float dt, const float (&start)[kWR], float (&out)[kWR]) {
#pragma unroll
    for (int r = 0; r < kWR; ++r) out[r] = start[r];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float kj[kWR];
        wload16(Kb + slot[j] * 512, lane, kj);
        const float a = dt * float(coef[j]);
#pragma unroll
        for (int r = 0; r < kWR; ++r) out[r] = fmaf(a, kj[r], out[r]);
    }
}

__device__ __forceinline__ void wstage_argument(int s, const float* __restrict__ Kb, int lane, float dt, const float (&y0)[kWR],
                                                float (&y)[kWR]) {
    switch (s) {
        case 1: { constexpr int sl[] = {0}; constexpr double cf[] = {1.0 / 5}; wlincomb<1>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 2: { constexpr int sl[] = {0, 1}; constexpr double cf[] = {3.0 / 40, 9.0 / 40}; wlincomb<2>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 3: { constexpr int sl[] = {0, 1, 2}; constexpr double cf[] = {44.0 / 45, -56.0 / 15, 32.0 / 9};
                  wlincomb<3>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 4: { constexpr int sl[] = {0, 1, 2, 3};
                  constexpr double cf[] = {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729};
                  wlincomb<4>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 5: { constexpr int sl[] = {0, 1, 2, 3, 4};
                  constexpr double cf[] = {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656};
                  wlincomb<5>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 6: { constexpr int sl[] = {0, 2, 3, 4, 5};
                  constexpr double cf[] = {35.0 / 384, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
                  wlincomb<5>(Kb, lane, sl, cf, dt, y0, y); break; }
        default: {
#pragma unroll
            for (int r = 0; r < kWR; ++r) y[r] = y0[r];
        }
    }
}

// 8-point Walsh-Hadamard butterfly on (sin, cos) pairs, in place on x[o .. o + 7]: image index -> sector index (and,
// applied again, sector -> image; the 1/8 is folded into the eigenvalues)
__device__ __forceinline__ void wht8(float2 (&x)[kWR], int o) {
    const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int h = 1; h < 8; h <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((i & h) == 0) {
                const float2 a = x[o + i], b = x[o + i + h];
                x[o + i] = __fadd2_rn(a, b);
                x[o + i + h] = __ffma2_rn(b, m1, a);
            }
        }
    }
}

__device__ __forceinline__ float2 bcast2(float v) { return make_float2(v, v); }

// sin / cos of a wrapped phase plus a stage increment (|x| well below 2^22): nearest multiple of 2 pi by the magic-number
// trick -- two FMA-pipe instructions instead of FMUL + FRND (FRND runs on the 16-lane XU pipe next to the MUFUs) -- one
// Cody-Waite constant (|k| <= 1 here: the dropped low part is k * 1.7e-7 rad, below MUFU.SIN's own 5e-7), then MUFU
#ifndef DBSGYM_PRECISE_SINCOS
__device__ __forceinline__ float wreduce(float x) {
    const float t = fmaf(x, 0.15915494309189535f, 12582912.0f);          // 1.5 * 2^23: the sum is rounded to an integer
    const float k = t - 12582912.0f;
#ifdef DBSGYM_WARP_CW2
    return fmaf(-k, -1.7484555314695172e-07f, fmaf(-k, 6.2831854820251465f, x));
#else
    return fmaf(-k, 6.2831854820251465f, x);
#endif
}
__device__ __forceinline__ void wsincos(float x, float* s, float* c) { const float r = wreduce(x); *s = __sinf(r); *c = __cosf(r); }
__device__ __forceinline__ float wcos(float x) { return __cosf(wreduce(x)); }
#else
__device__ __forceinline__ void wsincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ float wcos(float x) { return cosf(x); }
#endif

// Observation tail of the warp kernel: obs_tail (step_kernel.cuh; env.py:447-454, :638-650, :669-688) with the step's samples
// taken from shared memory instead of read back from global memory, and the twiddle products of the running rfft bins
// spread over all 32 lanes (lane = (sample group, bin): with 10 bins three groups share the samples) instead of one lane
// per bin walking all samples -- the tail is a pure latency chain, so its length is what it costs.
__device__ __forceinline__ void wobs_tail(const StepParams& p, int env, int lane, int S, double* t_delta, int* t_pos,
                                          const float* LF, double u) {
    const int W = p.W, nb = p.tail_nbins;
    float* ring = reinterpret_cast<float*>(p.ring) + (size_t)env * W;
    const int head = t_pos[32];
    const double2* tw = reinterpret_cast<const double2*>(p.tw_full);
    if (lane < S) {
        const int pos = t_pos[lane];
        const float v = LF[lane];
        ring[pos] = v;
        if (p.samples_f) p.samples_f[(size_t)env * p.smax + lane] = v;
        // (the pinned host log received the samples when they were formed, see the dense-output part of the kernel)
        t_delta[lane] = (double)v - t_delta[lane];
        if (p.trace) {                                // evaluation trace: theta_mean of the step (env.py:441)
            const int at = p.trace_len[env] + lane;
            if (at < p.trace_cap) p.trace[(size_t)env * p.trace_cap + at] = p.lfp_true[(size_t)env * p.smax + lane];
        }
    }
    __syncwarp();
    const int G = nb > 0 ? 32 / nb : 1;               // sample groups
    const int g = nb > 0 ? lane / nb : 0, b = nb > 0 ? lane - g * nb : 0;
    double re = 0.0, im = 0.0;
    if (g < G) {
#pragma unroll 4
        for (int j = g; j < S; j += G) {
            const double2 w = tw[(size_t)t_pos[j] * nb + b];
            re = fma(t_delta[j], w.x, re);
            im = fma(t_delta[j], w.y, im);
        }
    }
    for (int k = 1; k < G; ++k) {                     // (uniform trip count; lanes >= nb do not use the result)
        const double r2 = __shfl_sync(0xffffffffu, re, (lane + k * nb) & 31), i2 = __shfl_sync(0xffffffffu, im, (lane + k * nb) & 31);
        if (lane < nb) { re += r2; im += i2; }
    }
    double pw = 0.0;
    if (lane < nb) {
        double2* sp = reinterpret_cast<double2*>(p.spec) + (size_t)env * kTailBins + lane;
        double2 X = *sp;
        X.x += re; X.y += im;
        *sp = X;
        const double a = X.x / (double)W, c = X.y / (double)W;
        pw = (a * a + c * c) * 2.0;                   // utils.py:21-27: one-sided power of bin k
    }
    pw = warp_sum(pw);
    if (lane == 0) {
        const double au = fabs(u);
        double r;
        if (p.tail_kind == 0) r = -p.power_scale * pw - p.action_cost * au;                                   // env.py:638-650
        else r = -((p.power_scale * pw > p.threshold) ? p.threshold_penalty : 0.0) - p.action_cost * au;      // env.py:669-688
        p.reward[env] = r;
        if (p.reward_f) p.reward_f[env] = (float)r;
        const int k = p.step_idx_rw[env] + 1;
        p.step_idx_rw[env] = k;
        const uint8_t dn = k >= p.episode_len[env] ? 1 : 0;
        p.done_dev[env] = dn;
        if (p.done_out) p.done_out[env] = dn;
        int nh = head + S;
        if (nh >= W) nh -= W;
        p.head[env] = nh;
        if (p.trace) p.trace_len[env] = min(p.trace_len[env] + S, p.trace_cap);
        int nm = nh;                                  // reported position: the mirror's write column when it is on
        if (p.mirror) {
            nm = t_pos[33] + S;
            if (nm >= p.mir_len) nm -= p.mir_len;
            p.mpos[env] = nm;
        }
        if (p.nsamp_out) p.nsamp_out[env] = S;
        if (p.head_out) p.head_out[env] = nm;
    }
}

template <class RK>
__global__ void __maxnreg__(DBSGYM_WARP_MAXNREG) warp_step_kernel(const StepParams p) {
    using L = WarpLayout<RK>;
    constexpr int NM = L::NM;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = (int)(threadIdx.x & 31), wid = (int)(threadIdx.x >> 5), nwarp = (int)(blockDim.x >> 5);
    unsigned char* wsm = smem_raw + L::v_bytes + (size_t)wid * L::bytes_aligned;
    float* K = reinterpret_cast<float*>(wsm);                 // [kSlots][512], thread-private interleaved rows
    float* Pw = K + kSlots * 512;                             // [HR][RS] projection partials of one pass
    float* Cw = Pw + L::p_floats;                             // [NM] float2 mode coefficients x lambda
    int* WD = reinterpret_cast<int*>(Cw + L::c_floats);       // [16][32] winding counts
    float* C0 = reinterpret_cast<float*>(WD + 512);           // [16][32] w0 + pulse of the segment (interleaved like K)
    double* t_delta = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(C0 + 512) + 7) & ~uintptr_t(7));
    double* TS = t_delta + 32;                                // [2][kWarpTs] save times of the two segments of a step
    int* t_pos = reinterpret_cast<int*>(TS + 2 * kWarpTs);
    float* LF = reinterpret_cast<float*>(t_pos + 36);         // [32] recorded LFP samples of the step (what the window receives)

    const OctLane OL(lane);

    // eigenvector entries of this lane's two points, per mode pair (V0[m], V1[m], V0[m + 1], V1[m + 1]): registers for
    // the whole launch, or one table per CTA in shared memory; and the eigenvalues of the half rows this lane sums
    // (already multiplied by K / (8 N))
    constexpr int NP = L::NP, PP = L::PP;
    float4 Vr[L::VSMEM ? 1 : NP];
    const float4* Vs4 = reinterpret_cast<const float4*>(smem_raw) + lane;
    if constexpr (L::VSMEM) {
        for (int i = (int)threadIdx.x; i < NP * 32; i += (int)blockDim.x)
            reinterpret_cast<float4*>(smem_raw)[i] = __ldg(reinterpret_cast<const float4*>(p.spec_v) + i);
        __syncthreads();
    } else {
#pragma unroll
        for (int m2 = 0; m2 < NP; ++m2) Vr[m2] = __ldg(reinterpret_cast<const float4*>(p.spec_v) + m2 * 32 + lane);
    }
    auto vpair = [&](auto m2) -> float4 {
        if constexpr (L::VSMEM) return Vs4[decltype(m2)::value * 32];
        else return Vr[decltype(m2)::value];
    };
    float lam_r[L::PASSES][L::ROUNDS];
#pragma unroll
    for (int q = 0; q < L::PASSES; ++q) {
#pragma unroll
        for (int rd = 0; rd < L::ROUNDS; ++rd) {
            const int m = 2 * q * PP + ((lane + 32 * rd) >> 1);
            lam_r[q][rd] = (lane + 32 * rd < L::HR && m < NM) ? __ldg(p.spec_lam + m) : 0.f;
        }
    }

    const float rtol = (float)p.rtol, atol = (float)p.atol;
    const float two_pi_r = (float)kTwoPi;
    const float safety_f = (float)p.safety;
    const float inv_n = 1.0f / (float)p.N;

#pragma unroll 1
    for (int slot = (int)blockIdx.x + wid * (int)gridDim.x; slot < p.n_launch; slot += (int)gridDim.x * nwarp) {
    const int env = p.env_ids ? p.env_ids[slot] : slot;
    const size_t base = (size_t)env * p.Np;

    // every load that does not depend on another one is issued here, before anything waits: the per-environment scalars, then
    // the four state vectors of the lane; the rows of the NEXT environment of this warp are pulled into L2 meanwhile
    const bool step_mode = p.mode == MODE_STEP;
    const int k_idx = step_mode ? p.step_idx[env] : 0;
    const float act = step_mode ? p.actions[env] : 0.f;
    const bool fsal_in = p.fsal_on && step_mode && p.fsal_valid[env] != 0;
    float y0[kWR], c0v[kWR], stimv[kWR];
    oct_load<float, float2>(reinterpret_cast<const float*>(p.phase) + base, OL, y0);
    {
        int wd[kWR];
        oct_load<int, int2>(p.wind + base, OL, wd);
        if (DBSGYM_WARP_HOIST) {
            oct_load<float, float2>(reinterpret_cast<const float*>(p.w0) + base, OL, c0v);
            oct_load<float, float2>(reinterpret_cast<const float*>(p.stim) + base, OL, stimv);
        }
#pragma unroll
        for (int r = 0; r < kWR; ++r) WD[r * 32 + lane] = wd[r];
    }
    {
        const int nslot = slot + (int)gridDim.x * nwarp;
        if (DBSGYM_WARP_PREFETCH && nslot < p.n_launch && lane < 16) {
            const int nenv = p.env_ids ? p.env_ids[nslot] : nslot;
            const size_t nb_ = (size_t)nenv * p.Np * 4 + (size_t)lane * 128;          // 512 floats = 16 lines of 128 bytes per row
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.phase) + nb_));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.wind) + nb_));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.w0) + nb_));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.stim) + nb_));
            if (p.fsal_on) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.k_fsal) + nb_));
            if (p.weighted_rec) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.rec) + nb_));
        }
    }
    unsigned int n_acc = 0, n_rej = 0, n_rhs = 0, n_reuse = 0;
    bool k0_valid = false;                                // K slot 0 holds f(y0) for pulse amplitude amp_k0
    float amp_k0 = 0.f;
    if (fsal_in) {
        float k[kWR];
        wload16(reinterpret_cast<const float*>(p.k_fsal) + base, lane, k);      // (kept in this kernel's private layout)
        wstore16(K, lane, k);
        k0_valid = true;
    }
    int status = 0;

    // ---- segment programme (env.py:415-441 / :605-612) ----
    int nseg;
    double u_step = 0.0;
    const double* seg_ts[2];
    int seg_nts[2], seg_nrec[2], seg_from[2], seg_out[2];
    float seg_amp[2];
    if (step_mode) {
        int k = k_idx;
        if (k < 0 || k >= p.n_sched) { status |= STATUS_SCHEDULE; k = k < 0 ? 0 : p.n_sched - 1; }
        const int nI = p.sched_nI[k], nII = p.sched_nII[k];
        const double a = (double)act;                                       // env.py:389-393 rescale_action, env.py:419
        const double u = p.act_lo + ((p.act_hi - p.act_lo) * (a - (-1.0))) / (1.0 - (-1.0));
        u_step = u;
        if (lane == 0) { p.u_out[env] = u; p.n_samples[env] = nI + nII - 1; }
        nseg = 2;
        seg_ts[0] = p.sched_offI + (size_t)k * p.maxI;   seg_nts[0] = nI;  seg_nrec[0] = nI;
        seg_ts[1] = p.sched_offII + (size_t)k * p.maxII; seg_nts[1] = nII; seg_nrec[1] = nII - 1;
        seg_from[0] = seg_from[1] = 0;
        seg_out[0] = 0; seg_out[1] = nI;
        seg_amp[0] = (float)u; seg_amp[1] = 0.f;
        if (nI <= kWarpTs && nII <= kWarpTs) {               // the save times are consulted all the time: keep them on chip
            if (lane < nI) TS[lane] = seg_ts[0][lane];
            if (lane < nII) TS[kWarpTs + lane] = seg_ts[1][lane];
            seg_ts[0] = TS; seg_ts[1] = TS + kWarpTs;
        }
    } else {
        nseg = 1;
        seg_ts[0] = p.ts; seg_nts[0] = p.n_ts; seg_nrec[0] = p.n_ts - 1;
        seg_from[0] = seg_nrec[0] > p.W ? seg_nrec[0] - p.W : 0;
        seg_out[0] = 0; seg_amp[0] = 0.f;
        seg_ts[1] = nullptr; seg_nts[1] = seg_nrec[1] = seg_from[1] = seg_out[1] = 0; seg_amp[1] = 0.f;
    }
    const bool tail = p.tail_on && p.mode == MODE_STEP;
    if (tail) obs_tail_prefetch<float>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos);
    __syncwarp();

#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
        const double* __restrict__ ts = seg_ts[sg];
        const int n_ts = seg_nts[sg], n_rec = seg_nrec[sg], rec_from = seg_from[sg], out_base = seg_out[sg];
        const float amp = seg_amp[sg];
        {                                     // w0 + pulse, constant over the segment (env.py:254-255, :421-424)
            if (sg > 0 || !DBSGYM_WARP_HOIST) {  // (the first segment's vectors were fetched with the state; these hit L1 / L2)
                oct_load<float, float2>(reinterpret_cast<const float*>(p.w0) + base, OL, c0v);
                oct_load<float, float2>(reinterpret_cast<const float*>(p.stim) + base, OL, stimv);
            }
#pragma unroll
            for (int r = 0; r < kWR; ++r) c0v[r] = c0v[r] + amp * stimv[r];
            wstore16(C0, lane, c0v);
            if (k0_valid) {                      // k1 of this segment from the carried k7: only the pulse term changes
                float k[kWR];
                wload16(K, lane, k);
                const float da = amp - amp_k0;
#pragma unroll
                for (int r = 0; r < kWR; ++r) k[r] = fmaf(da, stimv[r], k[r]);
                wstore16(K, lane, k);
            }
        }
        const double T_end = ts[n_ts - 1];
        double t = 0.0;
        double tnext = fmin(p.dt0, T_end);
        int save_idx = 0;
        int attempts = 0;
        bool have_f0 = k0_valid;              // (still one logical RHS evaluation of the reference)
        if (have_f0) { ++n_rhs; ++n_reuse; }
        k0_valid = false;

        while (t < T_end) {
            if (++attempts > p.max_steps) { status |= STATUS_MAX_STEPS; break; }
            const double dt_d = tnext - t;
            const float dt = (float)dt_d;

#pragma unroll 1
            for (int s = have_f0 ? 1 : 0; s < 7; ++s) {
                float sv[kWR], cv[kWR];
                {
                    float ys[kWR];
                    if (s == 6) {                 // y1 = y0 + d1 with d1 summed on its own: k7 = f(y1) exactly (FSAL)
                        float zero[kWR];
#pragma unroll
                        for (int r = 0; r < kWR; ++r) zero[r] = 0.f;
                        wstage_argument(6, K, lane, dt, zero, ys);
#pragma unroll
                        for (int r = 0; r < kWR; ++r) ys[r] += y0[r];
                    } else wstage_argument(s, K, lane, dt, y0, ys);
#pragma unroll
                    for (int r = 0; r < kWR; ++r) wsincos(ys[r], &sv[r], &cv[r]);
                }
                float2 X[kWR];
#pragma unroll
                for (int r = 0; r < kWR; ++r) X[r] = make_float2(sv[r], cv[r]);
                wht8(X, 0);
                wht8(X, 8);
                // ---- mode sums, PASSES passes of PP mode pairs: projection partials P[mode][lane] of this lane, then one
                //      lane per half row adds 16 of them and the two halves of a mode meet by one shuffle ----
                static_for<L::PASSES>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    float* prow = Pw + (lane >> 4) * L::RS + 2 * (lane & 15);
                    static_for<PP>([&](auto jj) {
                        constexpr int j = decltype(jj)::value, m2 = q * PP + j;
                        if constexpr (m2 < NP) {
                            constexpr int sa = RK::sector_of(2 * m2), sb = RK::sector_of(2 * m2 + 1);
                            const float4 v = vpair(IC<m2>{});
                            float2 a = __fmul2_rn(bcast2(v.x), X[sa]), b = __fmul2_rn(bcast2(v.z), X[sb]);
                            a = __ffma2_rn(bcast2(v.y), X[8 + sa], a);
                            b = __ffma2_rn(bcast2(v.w), X[8 + sb], b);
                            *reinterpret_cast<float2*>(prow + (2 * j) * 2 * L::RS) = a;
                            *reinterpret_cast<float2*>(prow + (2 * j + 1) * 2 * L::RS) = b;
                        }
                    });
                    __syncwarp();
                    constexpr int live_rows = 4 * ((NP - q * PP) < PP ? (NP - q * PP) : PP);      // half rows written in this pass
                    // all rounds side by side (loads, then the add trees, then the shuffles): independent chains
                    float2 v[L::ROUNDS][16];
                    bool live[L::ROUNDS];
#pragma unroll
                    for (int rd = 0; rd < L::ROUNDS; ++rd) {
                        const int hg = lane + 32 * rd;
                        live[rd] = (live_rows >= 32 * (rd + 1)) || hg < live_rows;
                        const float4* r4 = reinterpret_cast<const float4*>(Pw + (live[rd] ? hg : 0) * L::RS);
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float4 x = r4[i]; v[rd][2 * i] = make_float2(x.x, x.y); v[rd][2 * i + 1] = make_float2(x.z, x.w); }
                    }
#pragma unroll
                    for (int w = 8; w > 0; w >>= 1) {
#pragma unroll
                        for (int rd = 0; rd < L::ROUNDS; ++rd) {
#pragma unroll
                            for (int i = 0; i < w; ++i) v[rd][i] = __fadd2_rn(v[rd][i], v[rd][i + w]);
                        }
                    }
                    float2 oth[L::ROUNDS];
#pragma unroll
                    for (int rd = 0; rd < L::ROUNDS; ++rd)
                        oth[rd] = make_float2(__shfl_xor_sync(FULL, v[rd][0].x, 1), __shfl_xor_sync(FULL, v[rd][0].y, 1));
#pragma unroll
                    for (int rd = 0; rd < L::ROUNDS; ++rd) {
                        const float2 tot = __fmul2_rn(bcast2(lam_r[q][rd]), __fadd2_rn(v[rd][0], oth[rd]));
                        if (live[rd] && !(lane & 1)) reinterpret_cast<float2*>(Cw)[2 * q * PP + ((lane + 32 * rd) >> 1)] = tot;
                    }
                    __syncwarp();
                });
                // ---- expansion back to the lane's sector coordinates, then sectors -> images ----
                static_for<8>([&](auto ss) {
                    constexpr int sec = decltype(ss)::value;
                    if constexpr (RK::get(sec) == 0) { X[sec] = make_float2(0.f, 0.f); X[8 + sec] = make_float2(0.f, 0.f); }
                });
                static_for<NP>([&](auto mm) {
                    constexpr int m2 = decltype(mm)::value, ma = 2 * m2, mb = ma + 1;
                    constexpr int sa = RK::sector_of(ma), sb = RK::sector_of(mb);
                    const float4 v = vpair(IC<m2>{});
                    const float4 c4 = reinterpret_cast<const float4*>(Cw)[m2];
                    const float2 ca = make_float2(c4.x, c4.y), cb = make_float2(c4.z, c4.w);
                    if constexpr (RK::first_of_sector(ma)) { X[sa] = __fmul2_rn(bcast2(v.x), ca); X[8 + sa] = __fmul2_rn(bcast2(v.y), ca); }
                    else { X[sa] = __ffma2_rn(bcast2(v.x), ca, X[sa]); X[8 + sa] = __ffma2_rn(bcast2(v.y), ca, X[8 + sa]); }
                    if constexpr (RK::first_of_sector(mb)) { X[sb] = __fmul2_rn(bcast2(v.z), cb); X[8 + sb] = __fmul2_rn(bcast2(v.w), cb); }
                    else { X[sb] = __ffma2_rn(bcast2(v.z), cb, X[sb]); X[8 + sb] = __ffma2_rn(bcast2(v.w), cb, X[8 + sb]); }
                });
                wht8(X, 0);
                wht8(X, 8);
                {
                    float ks[kWR];
                    wload16(C0, lane, ks);
#pragma unroll
                    for (int r = 0; r < kWR; ++r) ks[r] = fmaf(cv[r], X[r].x, fmaf(-sv[r], X[r].y, ks[r]));     // K / (8 N) is folded into lambda
                    wstore16(K + kslot(s) * 512, lane, ks);
                }
                ++n_rhs;
            }
            have_f0 = true;
            float d1[kWR];
            {
                float zero[kWR];
#pragma unroll
                for (int r = 0; r < kWR; ++r) zero[r] = 0.f;
                wstage_argument(6, K, lane, dt, zero, d1);    // y1 - y0, bit-identical to the last stage's increment
            }

            // ---- embedded error estimate and step-size controller (diffrax PIDController, I-only) ----
            float sqr = 0.f;
            {
                float e[kWR];
                {
                    constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                    constexpr double cf[] = {35.0 / 384 - 1951.0 / 21600, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                             -2187.0 / 6784 + 12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
                    float zero[kWR];
#pragma unroll
                    for (int r = 0; r < kWR; ++r) zero[r] = 0.f;
                    wlincomb<6>(K, lane, sl, cf, dt, zero, e);
                }
#pragma unroll
                for (int r = 0; r < kWR; ++r) {
                    const float yu0 = y0[r] + two_pi_r * (float)WD[r * 32 + lane];
                    const float yu1 = yu0 + d1[r];
                    const float scale = atol + fmaxf(fabsf(yu0), fabsf(yu1)) * rtol;
                    const float qv = __fdividef(e[r], scale);
                    sqr = fmaf(qv, qv, sqr);
                }
            }
            sqr = warp_sum(sqr);
            const float errf = sqrtf(sqr * inv_n);
            if (!(errf == errf)) { status |= STATUS_NAN; break; }
            const bool keep = errf < 1.0f;
            double factor;
            if (errf == 0.0f) factor = p.fmax;
            else factor = fmin(fmax((double)(safety_f * exp2f(-0.2f * __log2f(errf))), keep ? 1.0 : p.fmin), p.fmax);
            const double dt_next = dt_d * factor;

            double t_new0;
            if (keep) {
                ++n_acc;
                // ---- dense output (4th-order interpolant, increment form) + LFP samples ----
                if (save_idx < n_ts && ts[save_idx] <= tnext) {
                    float f0[kWR], pa[kWR], pb[kWR], pc[kWR];
                    {
                        float kk0[kWR], k6[kWR], dm[kWR];
                        wload16(K, lane, kk0);
                        wload16(K + kslot(6) * 512, lane, k6);
                        {
                            constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                            constexpr double cf[] = {0.5 * (6025192743.0 / 30085553152.0), 0.5 * (51252292925.0 / 65400821598.0),
                                                     0.5 * (-2691868925.0 / 45128329728.0), 0.5 * (187940372067.0 / 1594534317056.0),
                                                     0.5 * (-1776094331.0 / 19743644256.0), 0.5 * (11237099.0 / 235043384.0)};
                            float zero[kWR];
#pragma unroll
                            for (int r = 0; r < kWR; ++r) zero[r] = 0.f;
                            wlincomb<6>(K, lane, sl, cf, dt, zero, dm);
                        }
#pragma unroll
                        for (int r = 0; r < kWR; ++r) {
                            const float f0r = kk0[r] * dt, f1r = k6[r] * dt, dmr = dm[r], d = d1[r];
                            f0[r] = f0r;
                            pa[r] = 2.f * (f1r - f0r) - 8.f * d + 16.f * dmr;
                            pb[r] = 5.f * f0r - 3.f * f1r + 14.f * d - 32.f * dmr;
                            pc[r] = f1r - 4.f * f0r - 5.f * d + 16.f * dmr;
                        }
                    }
                    float rc[kWR];                       // recording conductance (env.py:404-412), L1 / L2 resident
                    if (p.weighted_rec) oct_load<float, float2>(reinterpret_cast<const float*>(p.rec) + base, OL, rc);
                    else {
#pragma unroll
                        for (int r = 0; r < kWR; ++r) rc[r] = 0.f;
                    }
                    // Samples in batches of up to 8.  Every lane leaves its two partial sums (plain, conductance weighted) of a
                    // sample in the partials buffer (free between RHS evaluations), S[slot][lane]; then lane l adds one quarter
                    // row of slot l & 7, two shuffles join the quarters, and lanes 0-7 store the 8 results: one short chain per
                    // batch instead of a shuffle tree per sample.
                    const float inv_h = (tnext == t) ? 0.f : 1.0f / (float)(tnext - t);
                    constexpr int SS = 68;                                   // words per slot row: 32 float2 + 4 (conflict-free both ways)
                    while (save_idx < n_ts && ts[save_idx] <= tnext) {
                        int first_idx = save_idx, nb = 0;
#pragma unroll 1
                        for (; nb < 8 && save_idx < n_ts && ts[save_idx] <= tnext; ++nb, ++save_idx) {
                            const double tsv = ts[save_idx];
                            float ysmp[kWR];
                            if (tsv == tnext) {                              // the end point of the sub-step is y1 itself
#pragma unroll
                                for (int r = 0; r < kWR; ++r) ysmp[r] = y0[r] + d1[r];
                            } else {
                                const float tau = (float)(tsv - t) * inv_h;
#pragma unroll
                                for (int r = 0; r < kWR; ++r)
                                    ysmp[r] = fmaf(fmaf(fmaf(fmaf(pa[r], tau, pb[r]), tau, pc[r]), tau, f0[r]), tau, y0[r]);
                            }
                            float st0 = 0.f, st1 = 0.f, sr0 = 0.f, sr1 = 0.f;
#pragma unroll
                            for (int r = 0; r < kWR; r += 2) {
                                const float ca = wcos(ysmp[r]), cb = wcos(ysmp[r + 1]);
                                st0 += ca; st1 += cb;
                                sr0 = fmaf(ca, rc[r], sr0); sr1 = fmaf(cb, rc[r + 1], sr1);
                            }
                            *reinterpret_cast<float2*>(Pw + nb * SS + 2 * lane) = make_float2(st0 + st1, sr0 + sr1);
                        }
                        __syncwarp();
                        {
                            const int slot = lane & 7, quarter = lane >> 3;
                            const float4* r4 = reinterpret_cast<const float4*>(Pw + slot * SS + quarter * 16);
                            float2 v[8];
#pragma unroll
                            for (int i = 0; i < 4; ++i) { const float4 x = r4[i]; v[2 * i] = make_float2(x.x, x.y); v[2 * i + 1] = make_float2(x.z, x.w); }
#pragma unroll
                            for (int w = 4; w > 0; w >>= 1) {
#pragma unroll
                                for (int i = 0; i < w; ++i) v[i] = __fadd2_rn(v[i], v[i + w]);
                            }
                            float a = v[0].x, b = v[0].y;
                            a += __shfl_xor_sync(FULL, a, 8);  b += __shfl_xor_sync(FULL, b, 8);
                            a += __shfl_xor_sync(FULL, a, 16); b += __shfl_xor_sync(FULL, b, 16);
                            const int idx = first_idx + lane;
                            if (lane < nb && idx >= rec_from && idx < n_rec) {
                                const double a_t = (double)(a * inv_n);
                                const double a_r = p.weighted_rec ? (double)(b * inv_n) : a_t;
                                if (p.mode == MODE_STEP) {
                                    p.lfp_true[(size_t)env * p.smax + out_base + idx] = a_t;
                                    p.lfp_rec[(size_t)env * p.smax + out_base + idx] = a_r;
                                    LF[(out_base + idx) & 31] = (float)a_r;
                                    if (DBSGYM_WARP_MIRROR_STORES && tail && p.mirror) {       // zero-copy store into the pinned host log (both copies) as soon as the
                                        float* mr = p.mirror + (size_t)env * 2 * p.mir_len;      // sample exists: PCIe drains under the integration
                                        int c = t_pos[33] + out_base + idx;
                                        if (c >= p.mir_len) c -= p.mir_len;
                                        mr[c] = (float)a_r;
                                        mr[c + p.mir_len] = (float)a_r;
                                    }
                                } else {
                                    reinterpret_cast<float*>(p.ring)[(size_t)env * p.W + (idx - rec_from)] = (float)a_r;
                                }
                            }
                        }
                        __syncwarp();                                        // the buffer is rewritten by the next batch / RHS evaluation
                    }
                }
                // ---- accept: y0 <- y1, FSAL k1 <- k7 ----
                {
                    float k6[kWR];
                    wload16(K + kslot(6) * 512, lane, k6);
                    wstore16(K, lane, k6);
                }
#pragma unroll
                for (int r = 0; r < kWR; ++r) {
                    float y1 = y0[r] + d1[r];
                    const float nwrap = floorf(y1 * 0.15915494309189535f);     // keep the phase wrapped: y = phase + 2 pi wind
                    if (nwrap != 0.0f) {
                        float yw = fmaf(-nwrap, 6.2831854820251465f, y1);
                        yw = fmaf(-nwrap, -1.7484555314695172e-07f, yw);
                        y1 = yw;
                        WD[r * 32 + lane] += (int)nwrap;
                    }
                    y0[r] = y1;
                }
                t_new0 = tnext;
            } else {
                ++n_rej;
                t_new0 = t;
            }
            const double new_t1 = t_new0 + dt_next;
            t = fmin(t_new0, T_end);
            tnext = (new_t1 > T_end - p.tol_end) ? (keep ? T_end : t + 0.5 * (T_end - t)) : new_t1;
        }
        if (status & (STATUS_MAX_STEPS | STATUS_NAN)) break;
        k0_valid = p.fsal_on != 0; amp_k0 = amp;
    }

    // ---- write back ----
    oct_store<float, float2>(reinterpret_cast<float*>(p.phase) + base, OL, y0);
    {
        int wd[kWR];
#pragma unroll
        for (int r = 0; r < kWR; ++r) wd[r] = WD[r * 32 + lane];
        oct_store<int, int2>(p.wind + base, OL, wd);
    }
    __syncwarp();                                                    // lane 0's LFP sample stores are visible to the tail's lanes
    if (tail) {
        if (DBSGYM_WARP_TAIL2) wobs_tail(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos, LF, u_step);
        else { __threadfence_block(); obs_tail<float>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos); }
    }
    if (p.fsal_on) {
        const bool keep_row = k0_valid && amp_k0 == 0.f;
        if (keep_row) {
            float k[kWR];
            wload16(K, lane, k);
            wstore16(reinterpret_cast<float*>(p.k_fsal) + base, lane, k);
        }
        if (lane == 0) p.fsal_valid[env] = keep_row ? 1 : 0;
    }
    if (lane == 0) {
        if (p.mode == MODE_TRANSIENT) p.head[env] = 0;
        atomicAdd(p.counters + 0, (unsigned long long)n_acc);
        atomicAdd(p.counters + 1, (unsigned long long)n_rej);
        atomicAdd(p.counters + 2, (unsigned long long)n_rhs);
        atomicAdd(p.counters + 3, (unsigned long long)n_reuse);
        if (status) atomicOr(p.status, status);
    }
    __syncwarp();
    }
}

}  // namespace dbsgym
