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
#include "step_kernel.cuh"

namespace dbsgym {

constexpr int kWR = 16;            // oscillators per lane
template <int V> struct IC { static constexpr int value = V; };

template <int... Rs> struct RankSet {
    static_assert(sizeof...(Rs) == 8, "one rank per parity sector");
    __host__ __device__ static constexpr int get(int s) { const int r[8] = {Rs...}; return r[s]; }
    __host__ __device__ static constexpr int off(int s) {             // first mode of sector s in the lane's eigenvector registers (off(8) = all modes)
        const int r[8] = {Rs...};
        int o = 0;
        for (int i = 0; i < s; ++i) o += r[i];
        return o;
    }
    __host__ __device__ static constexpr int coff(int s) {            // the same in the coefficient row, where sectors are padded to even counts
        const int r[8] = {Rs...};
        int o = 0;
        for (int i = 0; i < s; ++i) o += (r[i] + 1) & ~1;
        return o;
    }
};

template <class RK> struct WarpLayout {
    static constexpr int NM = RK::off(8);
    static constexpr int NC = RK::coff(8);
    static constexpr int RS = 36;                               // words per half row: 16 float2 partials + 4 (conflict-free both ways)
    static constexpr int HROWS = 2 * NM;                        // half rows: (mode, lanes 0-15 / 16-31)
    static constexpr int ROUNDS = (HROWS + 31) / 32;
    static constexpr int p_floats = HROWS * RS;
    static constexpr int c_floats = (2 * NC + 3) & ~3;
    // bytes of one warp's shared memory: K slots, partials + coefficients, winding counts, w0 + pulse, tail scratch
    static constexpr size_t bytes = (size_t)(kSlots * 512 + p_floats + c_floats + 512 + 512) * 4 + 32 * 8 + 36 * 4 + 8;
    static constexpr size_t bytes_aligned = (bytes + 15) & ~(size_t)15;
};

template <int S, class RK, class F> __device__ __forceinline__ void with_sector(F&& f) {
    f(IC<S>{}, IC<RK::get(S)>{}, IC<RK::off(S)>{}, IC<RK::coff(S)>{});
}
template <class RK, class F> __device__ __forceinline__ void for_each_sector(F&& f) {
    with_sector<0, RK>(f); with_sector<1, RK>(f); with_sector<2, RK>(f); with_sector<3, RK>(f);
    with_sector<4, RK>(f); with_sector<5, RK>(f); with_sector<6, RK>(f); with_sector<7, RK>(f);
}

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

template <class RK>
__global__ void __maxnreg__(DBSGYM_WARP_MAXNREG) warp_step_kernel(const StepParams p) {
    using L = WarpLayout<RK>;
    constexpr int NM = L::NM;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = (int)(threadIdx.x & 31), wid = (int)(threadIdx.x >> 5), nwarp = (int)(blockDim.x >> 5);
    unsigned char* wsm = smem_raw + (size_t)wid * L::bytes_aligned;
    float* K = reinterpret_cast<float*>(wsm);                 // [kSlots][512], thread-private interleaved rows
    float* Pw = K + kSlots * 512;                             // [HROWS][RS] projection partials
    float* Cw = Pw + L::p_floats;                             // [NC] float2 mode coefficients x lambda
    int* WD = reinterpret_cast<int*>(Cw + L::c_floats);       // [16][32] winding counts
    float* C0 = reinterpret_cast<float*>(WD + 512);           // [16][32] w0 + pulse of the segment (interleaved like K)
    double* t_delta = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(C0 + 512) + 7) & ~uintptr_t(7));
    int* t_pos = reinterpret_cast<int*>(t_delta + 32);

    const OctLane OL(lane);

    // eigenvector entries of this lane's two points (registers for the whole launch) and the eigenvalues of the half
    // rows it sums (already multiplied by K / (8 N))
    float V0[NM], V1[NM], lam_r[L::ROUNDS];
    int cslot_r[L::ROUNDS];                                   // where the mode of that half row goes in the coefficient row
    {
        const float2* v = reinterpret_cast<const float2*>(p.spec_v) + (size_t)lane * NM;
#pragma unroll
        for (int m = 0; m < NM; ++m) { const float2 t = __ldg(v + m); V0[m] = t.x; V1[m] = t.y; }
#pragma unroll
        for (int rd = 0; rd < L::ROUNDS; ++rd) {
            const int hg = lane + 32 * rd;
            lam_r[rd] = hg < L::HROWS ? __ldg(p.spec_lam + (hg >> 1)) : 0.f;
            const int m = hg >> 1;
            int cm = m;
            for_each_sector<RK>([&](auto, auto, auto oo, auto co) {
                constexpr int off = decltype(oo)::value, coff = decltype(co)::value;
                if (m >= off) cm = m - off + coff;
            });
            cslot_r[rd] = cm;
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

    float y0[kWR];
    oct_load<float, float2>(reinterpret_cast<const float*>(p.phase) + base, OL, y0);
    {
        int wd[kWR];
        oct_load<int, int2>(p.wind + base, OL, wd);
#pragma unroll
        for (int r = 0; r < kWR; ++r) WD[r * 32 + lane] = wd[r];
    }
    unsigned int n_acc = 0, n_rej = 0, n_rhs = 0, n_reuse = 0;
    bool k0_valid = false;                                // K slot 0 holds f(y0) for pulse amplitude amp_k0
    float amp_k0 = 0.f;
    if (p.fsal_on && p.mode == MODE_STEP && p.fsal_valid[env]) {
        float k[kWR];
        wload16(reinterpret_cast<const float*>(p.k_fsal) + base, lane, k);      // (kept in this kernel's private layout)
        wstore16(K, lane, k);
        k0_valid = true;
    }
    int status = 0;

    // ---- segment programme (env.py:415-441 / :605-612) ----
    int nseg;
    const double* seg_ts[2];
    int seg_nts[2], seg_nrec[2], seg_from[2], seg_out[2];
    float seg_amp[2];
    if (p.mode == MODE_STEP) {
        int k = p.step_idx[env];
        if (k < 0 || k >= p.n_sched) { status |= STATUS_SCHEDULE; k = k < 0 ? 0 : p.n_sched - 1; }
        const int nI = p.sched_nI[k], nII = p.sched_nII[k];
        const double a = (double)p.actions[env];                            // env.py:389-393 rescale_action, env.py:419
        const double u = p.act_lo + ((p.act_hi - p.act_lo) * (a - (-1.0))) / (1.0 - (-1.0));
        if (lane == 0) { p.u_out[env] = u; p.n_samples[env] = nI + nII - 1; }
        nseg = 2;
        seg_ts[0] = p.sched_offI + (size_t)k * p.maxI;   seg_nts[0] = nI;  seg_nrec[0] = nI;
        seg_ts[1] = p.sched_offII + (size_t)k * p.maxII; seg_nts[1] = nII; seg_nrec[1] = nII - 1;
        seg_from[0] = seg_from[1] = 0;
        seg_out[0] = 0; seg_out[1] = nI;
        seg_amp[0] = (float)u; seg_amp[1] = 0.f;
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
            float c0[kWR], stim[kWR];
            oct_load<float, float2>(reinterpret_cast<const float*>(p.w0) + base, OL, c0);
            oct_load<float, float2>(reinterpret_cast<const float*>(p.stim) + base, OL, stim);
#pragma unroll
            for (int r = 0; r < kWR; ++r) c0[r] = c0[r] + amp * stim[r];
            wstore16(C0, lane, c0);
            if (k0_valid) {                      // k1 of this segment from the carried k7: only the pulse term changes
                float k[kWR];
                wload16(K, lane, k);
                const float da = amp - amp_k0;
#pragma unroll
                for (int r = 0; r < kWR; ++r) k[r] = fmaf(da, stim[r], k[r]);
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
                    for (int r = 0; r < kWR; ++r) sincos_r(ys[r], &sv[r], &cv[r]);
                }
                float2 X[kWR];
#pragma unroll
                for (int r = 0; r < kWR; ++r) X[r] = make_float2(sv[r], cv[r]);
                wht8(X, 0);
                wht8(X, 8);
                // ---- projection: this lane's contribution to every mode sum, P[mode][lane] ----
                {
                    float* prow = Pw + (lane >> 4) * L::RS + 2 * (lane & 15);
                    for_each_sector<RK>([&](auto ss, auto rr, auto oo, auto) {
                        constexpr int sec = decltype(ss)::value, Rc = decltype(rr)::value, off = decltype(oo)::value;
#pragma unroll
                        for (int m = 0; m < Rc; ++m) {
                            float2 a = __fmul2_rn(bcast2(V0[off + m]), X[sec]);
                            a = __ffma2_rn(bcast2(V1[off + m]), X[8 + sec], a);
                            *reinterpret_cast<float2*>(prow + (off + m) * 2 * L::RS) = a;
                        }
                    });
                }
                __syncwarp();
                // ---- one lane per half row: sum of 16 lane partials; the two halves of a mode meet by one shuffle ----
#pragma unroll
                for (int rd = 0; rd < L::ROUNDS; ++rd) {
                    const int hg = lane + 32 * rd;
                    const bool live = (L::HROWS % 32 == 0) || hg < L::HROWS;
                    float2 tot = make_float2(0.f, 0.f);
                    if (live) {
                        const float4* r4 = reinterpret_cast<const float4*>(Pw + hg * L::RS);
                        float2 v[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float4 x = r4[i]; v[2 * i] = make_float2(x.x, x.y); v[2 * i + 1] = make_float2(x.z, x.w); }
#pragma unroll
                        for (int w = 8; w > 0; w >>= 1) {
#pragma unroll
                            for (int i = 0; i < w; ++i) v[i] = __fadd2_rn(v[i], v[i + w]);
                        }
                        tot = v[0];
                    }
                    const float2 oth = make_float2(__shfl_xor_sync(FULL, tot.x, 1), __shfl_xor_sync(FULL, tot.y, 1));
                    tot = __fmul2_rn(bcast2(lam_r[rd]), __fadd2_rn(tot, oth));
                    if (live && !(lane & 1)) reinterpret_cast<float2*>(Cw)[cslot_r[rd]] = tot;
                }
                __syncwarp();
                // ---- expansion back to the lane's sector coordinates, then sectors -> images ----
                for_each_sector<RK>([&](auto ss, auto rr, auto oo, auto co) {
                    constexpr int sec = decltype(ss)::value, Rc = decltype(rr)::value, off = decltype(oo)::value, coff = decltype(co)::value;
                    if constexpr (Rc == 0) {
                        X[sec] = make_float2(0.f, 0.f); X[8 + sec] = make_float2(0.f, 0.f);
                    } else {
                        constexpr int R2 = (Rc + 1) & ~1;
                        float2 cm[R2];
                        const float4* c4 = reinterpret_cast<const float4*>(Cw + 2 * coff);
#pragma unroll
                        for (int i = 0; i < R2 / 2; ++i) { const float4 x = c4[i]; cm[2 * i] = make_float2(x.x, x.y); cm[2 * i + 1] = make_float2(x.z, x.w); }
                        float2 a0 = __fmul2_rn(bcast2(V0[off]), cm[0]), a1 = __fmul2_rn(bcast2(V1[off]), cm[0]);
#pragma unroll
                        for (int m = 1; m < Rc; ++m) {
                            a0 = __ffma2_rn(bcast2(V0[off + m]), cm[m], a0);
                            a1 = __ffma2_rn(bcast2(V1[off + m]), cm[m], a1);
                        }
                        X[sec] = a0; X[8 + sec] = a1;
                    }
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
                    // up to 4 samples per pass: their lane sums are reduced together (independent shuffle chains)
                    while (save_idx < n_ts && ts[save_idx] <= tnext) {
                        float acc[4];
                        int sidx[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            acc[b] = 0.f; sidx[b] = -1;
                            if (save_idx < n_ts && ts[save_idx] <= tnext) {
                                const int idx = save_idx++;
                                if (idx >= rec_from && idx < n_rec) {
                                    sidx[b] = idx;
                                    const double tsv = ts[idx];
                                    const bool at_end = (tsv == tnext);
                                    const float tau = (tnext == t) ? 0.f : (float)(tsv - t) / (float)(tnext - t);
                                    float st = 0.f, sr = 0.f;
#pragma unroll
                                    for (int r = 0; r < kWR; ++r) {
                                        float inc = (((pa[r] * tau + pb[r]) * tau + pc[r]) * tau + f0[r]) * tau;
                                        if (at_end) inc = d1[r];
                                        const float c = cos_r(y0[r] + inc);
                                        st += c; sr = fmaf(c, rc[r], sr);
                                    }
                                    // lanes 0-15 go on with the plain sum, lanes 16-31 with the weighted one
                                    const float send = (lane & 16) ? st : sr, mine = (lane & 16) ? sr : st;
                                    acc[b] = mine + __shfl_xor_sync(FULL, send, 16);
                                }
                            }
                        }
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                            for (int b = 0; b < 4; ++b) acc[b] += __shfl_xor_sync(FULL, acc[b], o);
                        }
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const float wr = __shfl_sync(FULL, acc[b], 16);
                            if (lane == 0 && sidx[b] >= 0) {
                                const double a_t = (double)acc[b] / (double)p.N;
                                const double a_r = p.weighted_rec ? (double)wr / (double)p.N : a_t;
                                if (p.mode == MODE_STEP) {
                                    p.lfp_true[(size_t)env * p.smax + out_base + sidx[b]] = a_t;
                                    p.lfp_rec[(size_t)env * p.smax + out_base + sidx[b]] = a_r;
                                } else {
                                    reinterpret_cast<float*>(p.ring)[(size_t)env * p.W + (sidx[b] - rec_from)] = (float)a_r;
                                }
                            }
                        }
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
        __threadfence_block();
        obs_tail<float>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos);
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
