// Spectral step kernel for the 8 x 8 x 4 half grid (N = 256, the smallest point of BASELINE configs[4]): ONE WARP PER
// ENVIRONMENT, one octant point per lane (sm_100a, float32).
//
// The algorithm and the code structure of warp_kernel.cuh (adaptive Dopri5 + I-controller + dense output as the reference
// calls diffrax, environment/env.py:247-271; coupling sum of env.py:252-256 through the generalised mean-field identity over
// the eigenmodes of alpha; LFP samples env.py:396-412; fused observation tail env.py:447-454, :638-650, :669-688) with the
// geometry of the half grid: the fundamental octant has 2 x 4 x 4 = 32 points, so a lane owns ONE point and its 8 mirror
// images (8 oscillators, 128 registers), the mode sums run over 32 single-point partials per mode (two passes through a
// half-size buffer), and 16 environments share an SM.
//
// Lane l: zq = l >> 4, xq = (l >> 2) & 3, yq = l & 3 (octant point a = l).  Register r = image g = 4 my + 2 mz + mx (bit set =
// mirrored coordinate, z in {zq, 3 - zq}); after the butterfly index g is the sector s = 4 [odd y] + 2 [odd z] + [odd x].
#pragma once
#include "warp_kernel.cuh"
#ifndef DBSGYM_WARP1_ENVS
#define DBSGYM_WARP1_ENVS 16
#endif
// 1: the four coefficient vectors of the dense-output polynomial wait in the rows of k3 .. k6 (dead once they are formed)
// instead of 32 registers
#ifndef DBSGYM_WARP1_DENSE_SMEM
#define DBSGYM_WARP1_DENSE_SMEM 0
#endif
#ifndef DBSGYM_WARP1_PASSES
#define DBSGYM_WARP1_PASSES 2          // mode sums in two passes (half the partials buffer: 16 warps fit an SM)
#endif

namespace dbsgym {

constexpr int kW1R = 8;            // oscillators per lane
constexpr int kW1N = 256;          // oscillators per environment

template <class RK> struct Warp1Layout {
    static constexpr int NM = RK::off(8);                       // modes
    static constexpr int PASSES = DBSGYM_WARP1_PASSES;
    static constexpr int MP = (NM + PASSES - 1) / PASSES;       // modes per pass
    static constexpr int RS = 36;                               // words per half row: 16 float2 partials + 4 (conflict-free both ways)
    static constexpr int HR = 2 * MP;                           // half rows of a pass: (mode, lanes 0-15 / 16-31)
    static constexpr int ROUNDS = (HR + 31) / 32;
    static constexpr int p_floats = HR * RS > 8 * 68 ? HR * RS : 8 * 68;        // (also holds the lane sums of 8 LFP samples)
    static constexpr int c_floats = (2 * NM + 3) & ~3;
    static_assert(NM % 2 == 0, "modes come in pairs");
    // bytes of one warp's shared memory: K slots, partials + coefficients, winding counts, w0 + pulse, samples, tail scratch
    static constexpr size_t bytes = (size_t)(kSlots * kW1N + p_floats + c_floats + kW1N + kW1N + 32) * 4 + (32 + 2 * kWarpTs) * 8 + 36 * 4 + 8;
    static constexpr size_t bytes_aligned = (bytes + 15) & ~(size_t)15;
};

// thread-private rows of 8 floats in shared memory: piece q (4 floats) of lane l at float4 index q * 32 + l
__device__ __forceinline__ void w1load8(const float* __restrict__ row, int lane, float (&o)[kW1R]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) unpack(reinterpret_cast<const float4*>(row)[q * 32 + lane], o + 4 * q);
}
__device__ __forceinline__ void w1store8(float* __restrict__ row, int lane, const float (&o)[kW1R]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) reinterpret_cast<float4*>(row)[q * 32 + lane] = pack4(o + 4 * q);
}

// the lane's 8 entries of a natural-order [256] vector (index (z * 8 + x) * 8 + y, z < 4)
struct HalfGridPoint {
    int idx[kW1R];
    __device__ __forceinline__ explicit HalfGridPoint(int lane) {
        const int zq = lane >> 4, xq = (lane >> 2) & 3, yq = lane & 3;
#pragma unroll
        for (int g = 0; g < kW1R; ++g) {
            const int z = (g & 2) ? 3 - zq : zq, x = (g & 1) ? 7 - xq : xq, y = (g & 4) ? 7 - yq : yq;
            idx[g] = (z * 8 + x) * 8 + y;
        }
    }
};
template <typename T>
__device__ __forceinline__ void hg_load(const T* __restrict__ g, const HalfGridPoint& P, T (&o)[kW1R]) {
#pragma unroll
    for (int r = 0; r < kW1R; ++r) o[r] = g[P.idx[r]];
}
template <typename T>
__device__ __forceinline__ void hg_store(T* __restrict__ g, const HalfGridPoint& P, const T (&o)[kW1R]) {
#pragma unroll
    for (int r = 0; r < kW1R; ++r) g[P.idx[r]] = o[r];
}

template <int NJ>
__device__ __forceinline__ void w1lincomb(const float* __restrict__ Kb, int lane, const int (&slot)[NJ], const double (&coef)[NJ],
                                          float dt, const float (&start)[kW1R], float (&out)[kW1R]) {
#pragma unroll
    for (int r = 0; r < kW1R; ++r) out[r] = start[r];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float kj[kW1R];
        w1load8(Kb + slot[j] * kW1N, lane, kj);
        const float a = dt * float(coef[j]);
#pragma unroll
        for (int r = 0; r < kW1R; ++r) out[r] = fmaf(a, kj[r], out[r]);
    }
}

__device__ __forceinline__ void w1stage_argument(int s, const float* __restrict__ Kb, int lane, float dt, const float (&y0)[kW1R],
                                                 float (&y)[kW1R]) {
    switch (s) {
        case 1: { constexpr int sl[] = {0}; constexpr double cf[] = {1.0 / 5}; w1lincomb<1>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 2: { constexpr int sl[] = {0, 1}; constexpr double cf[] = {3.0 / 40, 9.0 / 40}; w1lincomb<2>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 3: { constexpr int sl[] = {0, 1, 2}; constexpr double cf[] = {44.0 / 45, -56.0 / 15, 32.0 / 9};
                  w1lincomb<3>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 4: { constexpr int sl[] = {0, 1, 2, 3};
                  constexpr double cf[] = {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729};
                  w1lincomb<4>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 5: { constexpr int sl[] = {0, 1, 2, 3, 4};
                  constexpr double cf[] = {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656};
                  w1lincomb<5>(Kb, lane, sl, cf, dt, y0, y); break; }
        case 6: { constexpr int sl[] = {0, 2, 3, 4, 5};
                  constexpr double cf[] = {35.0 / 384, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
                  w1lincomb<5>(Kb, lane, sl, cf, dt, y0, y); break; }
        default: {
#pragma unroll
            for (int r = 0; r < kW1R; ++r) y[r] = y0[r];
        }
    }
}

// 8-point Walsh-Hadamard butterfly on (sin, cos) pairs, in place
__device__ __forceinline__ void wht8p(float2 (&x)[kW1R]) {
    const float2 m1 = make_float2(-1.f, -1.f);
#pragma unroll
    for (int h = 1; h < 8; h <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((i & h) == 0) {
                const float2 a = x[i], b = x[i + h];
                x[i] = __fadd2_rn(a, b);
                x[i + h] = __ffma2_rn(b, m1, a);
            }
        }
    }
}

template <class RK>
__global__ void __launch_bounds__(DBSGYM_WARP1_ENVS * 32, 1) warp1_step_kernel(const StepParams p) {
    using L = Warp1Layout<RK>;
    constexpr int NM = L::NM, MP = L::MP;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = (int)(threadIdx.x & 31), wid = (int)(threadIdx.x >> 5), nwarp = (int)(blockDim.x >> 5);
    unsigned char* wsm = smem_raw + (size_t)wid * L::bytes_aligned;
    float* K = reinterpret_cast<float*>(wsm);                 // [kSlots][256], thread-private interleaved rows
    float* Pw = K + kSlots * kW1N;                            // [HR][RS] projection partials of one pass
    float* Cw = Pw + L::p_floats;                             // [NM] float2 mode coefficients x lambda
    int* WD = reinterpret_cast<int*>(Cw + L::c_floats);       // [8][32] winding counts
    float* C0 = reinterpret_cast<float*>(WD + kW1N);          // [2][32] float4: w0 + pulse of the segment
    float* LF = C0 + kW1N;                                    // [32] recorded LFP samples of the step
    double* t_delta = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(LF + 32) + 7) & ~uintptr_t(7));
    double* TS = t_delta + 32;                                // [2][kWarpTs] save times of the two segments of a step
    int* t_pos = reinterpret_cast<int*>(TS + 2 * kWarpTs);

    const HalfGridPoint OP(lane);

    // eigenvector entry of this lane's point for every mode (registers for the whole launch) and the eigenvalues of the
    // half rows it sums (already multiplied by K / (8 N))
    float V[NM], lam_r[L::PASSES][L::ROUNDS];
#pragma unroll
    for (int m = 0; m < NM; ++m) V[m] = __ldg(p.spec_v + m * 32 + lane);
#pragma unroll
    for (int q = 0; q < L::PASSES; ++q) {
#pragma unroll
        for (int rd = 0; rd < L::ROUNDS; ++rd) {
            const int hr = lane + 32 * rd, m = q * MP + (hr >> 1);
            lam_r[q][rd] = (hr < L::HR && m < NM) ? __ldg(p.spec_lam + m) : 0.f;
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
    const bool step_mode = p.mode == MODE_STEP;
    const int k_idx = step_mode ? p.step_idx[env] : 0;
    const float act = step_mode ? p.actions[env] : 0.f;
    const bool fsal_in = p.fsal_on && step_mode && p.fsal_valid[env] != 0;

    float y0[kW1R];
    hg_load<float>(reinterpret_cast<const float*>(p.phase) + base, OP, y0);
    {
        int wd[kW1R];
        hg_load<int>(p.wind + base, OP, wd);
#pragma unroll
        for (int r = 0; r < kW1R; ++r) WD[r * 32 + lane] = wd[r];
    }
    unsigned int n_acc = 0, n_rej = 0, n_rhs = 0, n_reuse = 0;
    bool k0_valid = false;                                // K slot 0 holds f(y0) for pulse amplitude amp_k0
    float amp_k0 = 0.f;
    if (fsal_in) {
        float k[kW1R];
        w1load8(reinterpret_cast<const float*>(p.k_fsal) + base, lane, k);      // (kept in this kernel's private layout)
        w1store8(K, lane, k);
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
    const bool tail = p.tail_on && step_mode;
    if (tail) obs_tail_prefetch<float>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos);
    __syncwarp();

#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
        const double* __restrict__ ts = seg_ts[sg];
        const int n_ts = seg_nts[sg], n_rec = seg_nrec[sg], rec_from = seg_from[sg], out_base = seg_out[sg];
        const float amp = seg_amp[sg];
        {                                     // w0 + pulse, constant over the segment (env.py:254-255, :421-424)
            float c0v[kW1R], stimv[kW1R];
            hg_load<float>(reinterpret_cast<const float*>(p.w0) + base, OP, c0v);
            hg_load<float>(reinterpret_cast<const float*>(p.stim) + base, OP, stimv);
#pragma unroll
            for (int r = 0; r < kW1R; ++r) c0v[r] = c0v[r] + amp * stimv[r];
            w1store8(C0, lane, c0v);
            if (k0_valid) {                      // k1 of this segment from the carried k7: only the pulse term changes
                float k[kW1R];
                w1load8(K, lane, k);
                const float da = amp - amp_k0;
#pragma unroll
                for (int r = 0; r < kW1R; ++r) k[r] = fmaf(da, stimv[r], k[r]);
                w1store8(K, lane, k);
            }
        }
        const double T_end = ts[n_ts - 1];
        double tt = 0.0;
        double tnext = fmin(p.dt0, T_end);
        int save_idx = 0;
        int attempts = 0;
        bool have_f0 = k0_valid;              // (still one logical RHS evaluation of the reference)
        if (have_f0) { ++n_rhs; ++n_reuse; }
        k0_valid = false;

        while (tt < T_end) {
            if (++attempts > p.max_steps) { status |= STATUS_MAX_STEPS; break; }
            const double dt_d = tnext - tt;
            const float dt = (float)dt_d;

#pragma unroll 1
            for (int s = have_f0 ? 1 : 0; s < 7; ++s) {
                float sv[kW1R], cv[kW1R];
                {
                    float ys[kW1R];
                    if (s == 6) {                 // y1 = y0 + d1 with d1 summed on its own: k7 = f(y1) exactly (FSAL)
                        float zero[kW1R];
#pragma unroll
                        for (int r = 0; r < kW1R; ++r) zero[r] = 0.f;
                        w1stage_argument(6, K, lane, dt, zero, ys);
#pragma unroll
                        for (int r = 0; r < kW1R; ++r) ys[r] += y0[r];
                    } else w1stage_argument(s, K, lane, dt, y0, ys);
#pragma unroll
                    for (int r = 0; r < kW1R; ++r) wsincos(ys[r], &sv[r], &cv[r]);
                }
                float2 X[kW1R];
#pragma unroll
                for (int r = 0; r < kW1R; ++r) X[r] = make_float2(sv[r], cv[r]);
                wht8p(X);
                // ---- mode sums, two passes of MP modes: the lane's partial of every mode, P[mode][lane], then one lane per half
                //      row adds 16 of them and the two halves of a mode meet by one shuffle ----
                static_for<L::PASSES>([&](auto qq) {
                    constexpr int q = decltype(qq)::value;
                    float* prow = Pw + (lane >> 4) * L::RS + 2 * (lane & 15);
                    static_for<MP>([&](auto jj) {
                        constexpr int j = decltype(jj)::value, m = q * MP + j;
                        if constexpr (m < NM) {
                            constexpr int sec = RK::sector_of(m);
                            *reinterpret_cast<float2*>(prow + j * 2 * L::RS) = __fmul2_rn(bcast2(V[m]), X[sec]);
                        }
                    });
                    __syncwarp();
                    constexpr int live_rows = 2 * ((NM - q * MP) < MP ? (NM - q * MP) : MP);      // half rows written in this pass
#pragma unroll
                    for (int rd = 0; rd < L::ROUNDS; ++rd) {
                        const int hr = lane + 32 * rd;
                        const bool live = (live_rows >= 32 * (rd + 1)) || hr < live_rows;
                        const float4* r4 = reinterpret_cast<const float4*>(Pw + (live ? hr : 0) * L::RS);
                        float2 v[16];
#pragma unroll
                        for (int i = 0; i < 8; ++i) { const float4 x = r4[i]; v[2 * i] = make_float2(x.x, x.y); v[2 * i + 1] = make_float2(x.z, x.w); }
#pragma unroll
                        for (int w = 8; w > 0; w >>= 1) {
#pragma unroll
                            for (int i = 0; i < w; ++i) v[i] = __fadd2_rn(v[i], v[i + w]);
                        }
                        float2 tot = v[0];
                        tot = __fadd2_rn(tot, make_float2(__shfl_xor_sync(FULL, tot.x, 1), __shfl_xor_sync(FULL, tot.y, 1)));
                        tot = __fmul2_rn(bcast2(lam_r[q][rd]), tot);
                        if (live && !(lane & 1)) reinterpret_cast<float2*>(Cw)[q * MP + (hr >> 1)] = tot;
                    }
                    __syncwarp();
                });
                // ---- expansion back to the lane's sector coordinates, then sectors -> images ----
                static_for<8>([&](auto ss) {
                    constexpr int sec = decltype(ss)::value;
                    if constexpr (RK::get(sec) == 0) X[sec] = make_float2(0.f, 0.f);
                });
                static_for<NM / 2>([&](auto mm) {
                    constexpr int ma = 2 * decltype(mm)::value, mb = ma + 1;
                    constexpr int sa = RK::sector_of(ma), sb = RK::sector_of(mb);
                    const float4 c4 = reinterpret_cast<const float4*>(Cw)[ma >> 1];
                    const float2 ca = make_float2(c4.x, c4.y), cb = make_float2(c4.z, c4.w);
                    if constexpr (RK::first_of_sector(ma)) X[sa] = __fmul2_rn(bcast2(V[ma]), ca);
                    else X[sa] = __ffma2_rn(bcast2(V[ma]), ca, X[sa]);
                    if constexpr (RK::first_of_sector(mb)) X[sb] = __fmul2_rn(bcast2(V[mb]), cb);
                    else X[sb] = __ffma2_rn(bcast2(V[mb]), cb, X[sb]);
                });
                wht8p(X);
                {
                    float ks[kW1R];
                    w1load8(C0, lane, ks);
#pragma unroll
                    for (int r = 0; r < kW1R; ++r) ks[r] = fmaf(cv[r], X[r].x, fmaf(-sv[r], X[r].y, ks[r]));     // K / (8 N) is folded into lambda
                    w1store8(K + kslot(s) * kW1N, lane, ks);
                }
                ++n_rhs;
            }
            have_f0 = true;
            float d1[kW1R];
            {
                float zero[kW1R];
#pragma unroll
                for (int r = 0; r < kW1R; ++r) zero[r] = 0.f;
                w1stage_argument(6, K, lane, dt, zero, d1);  // y1 - y0, bit-identical to the last stage's increment
            }

            // ---- embedded error estimate and step-size controller (diffrax PIDController, I-only) ----
            float sqr = 0.f;
            {
                float e[kW1R];
                {
                    constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                    constexpr double cf[] = {35.0 / 384 - 1951.0 / 21600, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                             -2187.0 / 6784 + 12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
                    float zero[kW1R];
#pragma unroll
                    for (int r = 0; r < kW1R; ++r) zero[r] = 0.f;
                    w1lincomb<6>(K, lane, sl, cf, dt, zero, e);
                }
#pragma unroll
                for (int r = 0; r < kW1R; ++r) {
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
                    float f0[kW1R], pa[kW1R], pb[kW1R], pc[kW1R];
                    {
                        float kk0[kW1R], k6[kW1R], dm[kW1R];
                        w1load8(K, lane, kk0);
                        w1load8(K + kslot(6) * kW1N, lane, k6);
                        {
                            constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                            constexpr double cf[] = {0.5 * (6025192743.0 / 30085553152.0), 0.5 * (51252292925.0 / 65400821598.0),
                                                     0.5 * (-2691868925.0 / 45128329728.0), 0.5 * (187940372067.0 / 1594534317056.0),
                                                     0.5 * (-1776094331.0 / 19743644256.0), 0.5 * (11237099.0 / 235043384.0)};
                            float zero[kW1R];
#pragma unroll
                            for (int r = 0; r < kW1R; ++r) zero[r] = 0.f;
                            w1lincomb<6>(K, lane, sl, cf, dt, zero, dm);
                        }
#pragma unroll
                        for (int r = 0; r < kW1R; ++r) {
                            const float f0r = kk0[r] * dt, f1r = k6[r] * dt, dmr = dm[r], d = d1[r];
                            f0[r] = f0r;
                            pa[r] = 2.f * (f1r - f0r) - 8.f * d + 16.f * dmr;
                            pb[r] = 5.f * f0r - 3.f * f1r + 14.f * d - 32.f * dmr;
                            pc[r] = f1r - 4.f * f0r - 5.f * d + 16.f * dmr;
                        }
                        if (DBSGYM_WARP1_DENSE_SMEM) {
                            w1store8(K + 2 * kW1N, lane, pa); w1store8(K + 3 * kW1N, lane, pb);
                            w1store8(K + 4 * kW1N, lane, pc); w1store8(K + 5 * kW1N, lane, f0);
                        }
                    }
                    float rc[kW1R];                      // recording conductance (env.py:404-412), L1 / L2 resident
                    if (p.weighted_rec) hg_load<float>(reinterpret_cast<const float*>(p.rec) + base, OP, rc);
                    else {
#pragma unroll
                        for (int r = 0; r < kW1R; ++r) rc[r] = 0.f;
                    }
                    // Samples in batches of up to 8 (as in warp_kernel.cuh): lane sums through the idle partials buffer
                    const float inv_h = (tnext == tt) ? 0.f : 1.0f / (float)(tnext - tt);
                    constexpr int SS = 68;                                   // words per slot row: 32 float2 + 4 (conflict-free both ways)
                    while (save_idx < n_ts && ts[save_idx] <= tnext) {
                        int first_idx = save_idx, nb = 0;
#pragma unroll 1
                        for (; nb < 8 && save_idx < n_ts && ts[save_idx] <= tnext; ++nb, ++save_idx) {
                            const double tsv = ts[save_idx];
                            float ysmp[kW1R];
                            if (tsv == tnext) {                              // the end point of the sub-step is y1 itself
#pragma unroll
                                for (int r = 0; r < kW1R; ++r) ysmp[r] = y0[r] + d1[r];
                            } else {
                                const float tau = (float)(tsv - tt) * inv_h;
                                if (DBSGYM_WARP1_DENSE_SMEM) {
                                    float c[kW1R];
                                    w1load8(K + 2 * kW1N, lane, ysmp);
                                    w1load8(K + 3 * kW1N, lane, c);
#pragma unroll
                                    for (int r = 0; r < kW1R; ++r) ysmp[r] = fmaf(ysmp[r], tau, c[r]);
                                    w1load8(K + 4 * kW1N, lane, c);
#pragma unroll
                                    for (int r = 0; r < kW1R; ++r) ysmp[r] = fmaf(ysmp[r], tau, c[r]);
                                    w1load8(K + 5 * kW1N, lane, c);
#pragma unroll
                                    for (int r = 0; r < kW1R; ++r) ysmp[r] = fmaf(fmaf(ysmp[r], tau, c[r]), tau, y0[r]);
                                } else {
#pragma unroll
                                for (int r = 0; r < kW1R; ++r)
                                    ysmp[r] = fmaf(fmaf(fmaf(fmaf(pa[r], tau, pb[r]), tau, pc[r]), tau, f0[r]), tau, y0[r]);
                                }
                            }
                            float st0 = 0.f, st1 = 0.f, sr0 = 0.f, sr1 = 0.f;
#pragma unroll
                            for (int r = 0; r < kW1R; r += 2) {
                                const float ca = wcos(ysmp[r]), cb = wcos(ysmp[r + 1]);
                                st0 += ca; st1 += cb;
                                sr0 = fmaf(ca, rc[r], sr0); sr1 = fmaf(cb, rc[r + 1], sr1);
                            }
                            *reinterpret_cast<float2*>(Pw + nb * SS + 2 * lane) = make_float2(st0 + st1, sr0 + sr1);
                        }
                        __syncwarp();
                        {
                            const int sl_ = lane & 7, quarter = lane >> 3;
                            const float4* r4 = reinterpret_cast<const float4*>(Pw + sl_ * SS + quarter * 16);
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
                                if (step_mode) {
                                    p.lfp_true[(size_t)env * p.smax + out_base + idx] = a_t;
                                    p.lfp_rec[(size_t)env * p.smax + out_base + idx] = a_r;
                                    LF[(out_base + idx) & 31] = (float)a_r;
                                    if (tail && p.mirror) {       // zero-copy store into the pinned host log (both copies)
                                        float* mr = p.mirror + (size_t)env * 2 * p.mir_len;
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
                    float k6[kW1R];
                    w1load8(K + kslot(6) * kW1N, lane, k6);
                    w1store8(K, lane, k6);
                }
#pragma unroll
                for (int r = 0; r < kW1R; ++r) {
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
                t_new0 = tt;
            }
            const double new_t1 = t_new0 + dt_next;
            tt = fmin(t_new0, T_end);
            tnext = (new_t1 > T_end - p.tol_end) ? (keep ? T_end : tt + 0.5 * (T_end - tt)) : new_t1;
        }
        if (status & (STATUS_MAX_STEPS | STATUS_NAN)) break;
        k0_valid = p.fsal_on != 0; amp_k0 = amp;
    }

    // ---- write back ----
    hg_store<float>(reinterpret_cast<float*>(p.phase) + base, OP, y0);
    {
        int wd[kW1R];
#pragma unroll
        for (int r = 0; r < kW1R; ++r) wd[r] = WD[r * 32 + lane];
        hg_store<int>(p.wind + base, OP, wd);
    }
    __syncwarp();                                                    // the step's samples (LF) are visible to the tail's lanes
    if (tail) wobs_tail(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos, LF, u_step);
    if (p.fsal_on) {
        const bool keep_row = k0_valid && amp_k0 == 0.f;
        if (keep_row) {
            float k[kW1R];
            w1load8(K, lane, k);
            w1store8(reinterpret_cast<float*>(p.k_fsal) + base, lane, k);
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
