// Spectral step kernel for regular grids of 1024 ... 4096 oscillators: ONE CTA PER ENVIRONMENT, one octant point per thread,
// eigenvectors in registers (sm_100a, float32); 8192 oscillators as a thread-block cluster of two such CTAs (NC = 2, see below).
//
// The algorithm of warp_kernel.cuh / warp1_kernel.cuh (adaptive Dopri5 + I-controller + dense output as the reference calls
// diffrax, environment/env.py:247-271; coupling sum of env.py:252-256 through the generalised mean-field identity over the
// eigenmodes of alpha; LFP samples env.py:396-412; fused observation tail env.py:447-454, :638-650, :669-688) for grids whose
// fundamental octant has more than 32 points: the environment is spread over NW = N / 256 warps of one CTA.
//
//   * the handle keeps the oscillators in OCTANT ORDER (dbsgym_set_coupling_lowrank_sectors: position 8 a + g = mirror image
//     g = 4 my + 2 mz + mx of octant point a), so thread a owns 8 consecutive floats of every state vector: two 16-byte
//     accesses, and the parity-sector transform is the in-register 8-point Walsh-Hadamard butterfly;
//   * the thread keeps its entry of every eigenvector in registers (compile-time rank list: 38 ... 62 modes) -- the block
//     kernel of step_kernel.cuh (CPL_LOWRANK, sector form) streams them from L2 twice per evaluation, which is what bounds it;
//   * mode sums: inside a warp as in warp1_kernel.cuh (partials through shared memory, one lane per half row), then ONE
//     __syncthreads per evaluation: every warp leaves its totals in a double-buffered table and, after the barrier, adds the
//     NW rows itself (lane = mode; the same order in every warp, so all warps hold bit-identical coefficients);
//   * error norm and LFP samples cross the warps the same way (double-buffered per-warp partials, one barrier each); the
//     controller runs redundantly in every thread on identical numbers, so the control flow stays uniform over the CTA.
#pragma once
#include "warp1_kernel.cuh"

// 1 / 2 (A/B switch, measured: within 1 % either way): sin / cos of the stage argument are formed again after the mode sums instead of being kept in 16 registers across
// them (1: the argument stays in 8 registers, 2: it waits in shared memory)
#ifndef DBSGYM_OCT_RESINCOS
#define DBSGYM_OCT_RESINCOS 0
#endif

// 1: the four coefficient vectors of the dense-output polynomial wait in the rows of k3 .. k6 (dead once they are formed)
// instead of 32 registers
#ifndef DBSGYM_OCT_DENSE_SMEM
#define DBSGYM_OCT_DENSE_SMEM 1          // (measured on B200: 0.698 -> 0.640 ms per 512-env step at N = 4096; no spills at 38 modes)
#endif

namespace dbsgym {

constexpr int kOctMP = 13;         // modes per pass of the in-warp reduction: 26 half rows, one round of 32 lanes

template <class RK, int NW, int NC = 1> struct OctLayout {
    static constexpr int NM = RK::off(8);                       // modes (even: the expansion takes them in pairs)
    static constexpr int NT = 32 * NW, N = 8 * NT;          // threads and oscillators of ONE CTA (the environment has NC of them)
    static constexpr int MP = kOctMP;
    static constexpr int PASSES = (NM + MP - 1) / MP;
    static constexpr int RS = 36;                               // words per half row: 16 float2 partials + 4
    static constexpr int p_floats = 2 * MP * RS;                // per warp (>= 8 * 68: also holds the lane sums of 8 LFP samples)
    static constexpr int NMP = (NM + 3) & ~3;                   // padded mode count of the coefficient tables
    static_assert(NM % 2 == 0, "modes come in pairs");
    static_assert(p_floats >= 8 * 68, "the partials buffer also holds the lane sums of 8 LFP samples");
    // floats: K slots, winding counts, w0 + pulse, per-warp partials, per-warp totals (2 buffers), per-warp coefficients,
    // error-norm partials (2 buffers), sample partials (2 buffers of [NW][8] float2), samples of the step
    static constexpr int CPB = NC > 1 ? 1 : 2;                  // buffers of the warps' totals (a cluster has a second barrier per evaluation)
    static constexpr size_t floats = (size_t)kSlots * N + N + N + (size_t)NW * p_floats + CPB * (size_t)NW * 2 * NMP + (size_t)NW * 2 * NMP +
                                     2 * NW + 2 * NW * 16 + 32 +
                                     (NC > 1 ? 2 * NC * 2 * NMP + 2 * NC + 2 * NC * 16 : 0);   // cluster exchange tables XC, RX, SX
    static constexpr size_t bytes = floats * 4 + (32 + 2 * kWarpTs) * 8 + 36 * 4 + 16;
    static_assert(bytes <= (size_t)227 * 1024, "one environment (or its part) must fit the shared memory of an SM");
};

template <int NT> __device__ __forceinline__ void oload8(const float* __restrict__ row, int tid, float (&o)[8]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) unpack(reinterpret_cast<const float4*>(row)[q * NT + tid], o + 4 * q);
}
template <int NT> __device__ __forceinline__ void ostore8(float* __restrict__ row, int tid, const float (&o)[8]) {
#pragma unroll
    for (int q = 0; q < 2; ++q) reinterpret_cast<float4*>(row)[q * NT + tid] = pack4(o + 4 * q);
}
// the thread's 8 consecutive entries of an octant-ordered global vector
__device__ __forceinline__ void gload8(const float* __restrict__ g, float (&o)[8]) {
    unpack(reinterpret_cast<const float4*>(g)[0], o);
    unpack(reinterpret_cast<const float4*>(g)[1], o + 4);
}
__device__ __forceinline__ void gstore8(float* __restrict__ g, const float (&o)[8]) {
    reinterpret_cast<float4*>(g)[0] = pack4(o);
    reinterpret_cast<float4*>(g)[1] = pack4(o + 4);
}

template <int NT, int NJ>
__device__ __forceinline__ void olincomb(const float* __restrict__ Kb, int tid, const int (&slot)[NJ], const double (&coef)[NJ],
                                         float dt, const float (&start)[8], float (&out)[8]) {
#pragma unroll
    for (int r = 0; r < 8; ++r) out[r] = start[r];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float kj[8];
        oload8<NT>(Kb + slot[j] * (8 * NT), tid, kj);
        const float a = dt * float(coef[j]);
#pragma unroll
        for (int r = 0; r < 8; ++r) out[r] = fmaf(a, kj[r], out[r]);
    }
}

template <int NT>
__device__ __forceinline__ void ostage_argument(int s, const float* __restrict__ Kb, int tid, float dt, const float (&y0)[8], float (&y)[8]) {
    switch (s) {
        case 1: { constexpr int sl[] = {0}; constexpr double cf[] = {1.0 / 5}; olincomb<NT, 1>(Kb, tid, sl, cf, dt, y0, y); break; }
        case 2: { constexpr int sl[] = {0, 1}; constexpr double cf[] = {3.0 / 40, 9.0 / 40}; olincomb<NT, 2>(Kb, tid, sl, cf, dt, y0, y); break; }
        case 3: { constexpr int sl[] = {0, 1, 2}; constexpr double cf[] = {44.0 / 45, -56.0 / 15, 32.0 / 9};
                  olincomb<NT, 3>(Kb, tid, sl, cf, dt, y0, y); break; }
        case 4: { constexpr int sl[] = {0, 1, 2, 3};
                  constexpr double cf[] = {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729};
                  olincomb<NT, 4>(Kb, tid, sl, cf, dt, y0, y); break; }
        case 5: { constexpr int sl[] = {0, 1, 2, 3, 4};
                  constexpr double cf[] = {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656};
                  olincomb<NT, 5>(Kb, tid, sl, cf, dt, y0, y); break; }
        case 6: { constexpr int sl[] = {0, 2, 3, 4, 5};
                  constexpr double cf[] = {35.0 / 384, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
                  olincomb<NT, 5>(Kb, tid, sl, cf, dt, y0, y); break; }
        default: {
#pragma unroll
            for (int r = 0; r < 8; ++r) y[r] = y0[r];
        }
    }
}

// stores into the shared memory of CTA `rank` of the cluster, at the address `local_ptr` has in this CTA
__device__ __forceinline__ void st_cluster(const void* local_ptr, int rank, float2 v) {
    const uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_ptr);
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_cluster(const void* local_ptr, int rank, float v) {
    const uint32_t laddr = (uint32_t)__cvta_generic_to_shared(local_ptr);
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}

// NC > 1 (8192 oscillators: NC = 2): the environment is a thread-block cluster of NC such CTAs.  Every sum over the
// environment then takes a second level: the CTA's total is PUSHED into the exchange tables of all CTAs of the cluster
// (st.shared::cluster, double buffered), one cluster barrier, and every CTA adds the NC entries in the same order.
template <class RK, int NW, int NC>
__global__ void __launch_bounds__(32 * NW, NC > 1 ? 1 : 16 / NW) oct_step_kernel(const StepParams p) {
    using L = OctLayout<RK, NW, NC>;
    constexpr int NM = L::NM, MP = L::MP, NT = L::NT, N = L::N, NMP = L::NMP;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = (int)threadIdx.x, lane = tid & 31, wid = tid >> 5;
    float* K = reinterpret_cast<float*>(smem_raw);            // [kSlots][N], thread-private interleaved rows
    int* WD = reinterpret_cast<int*>(K + kSlots * N);         // [8][NT] winding counts
    float* C0 = reinterpret_cast<float*>(WD + N);             // [2][NT] float4: w0 + pulse of the segment
    float* Pw = C0 + N + wid * L::p_floats;                   // this warp's projection partials of one pass
    float* CP = C0 + N + NW * L::p_floats;                    // [2][NW][NMP] float2: the warps' mode totals, double buffered
    float* Cw = CP + L::CPB * NW * 2 * NMP + wid * 2 * NMP;   // this warp's copy of the mode coefficients
    float* RD = CP + L::CPB * NW * 2 * NMP + NW * 2 * NMP;    // [2][NW] error-norm partials
    float* SL = RD + 2 * NW;                                  // [2][NW][8] float2 sample partials
    float* LF = SL + 2 * NW * 16;                             // [32] recorded LFP samples of the step
    float* XC = LF + 32;                                      // cluster: [2][NC][NMP] float2 mode totals of the CTAs
    float* RX = XC + (NC > 1 ? 2 * NC * 2 * NMP : 0);         // cluster: [2][NC] error-norm totals of the CTAs
    float* SX = RX + (NC > 1 ? 2 * NC : 0);                   // cluster: [2][NC][8] float2 sample totals of the CTAs (read by CTA 0)
    float* after = SX + (NC > 1 ? 2 * NC * 16 : 0);
    double* t_delta = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(after) + 7) & ~uintptr_t(7));
    int crank = 0;
    if constexpr (NC > 1) asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int pt = crank * NT + tid;                          // octant point of this thread
    auto sync_env = [&]() { if constexpr (NC > 1) cluster_barrier(); else __syncthreads(); };
    double* TS = t_delta + 32;                                // [2][kWarpTs] save times of the two segments of a step
    int* t_pos = reinterpret_cast<int*>(TS + 2 * kWarpTs);

    // eigenvector entry of this thread's octant point for every mode, and the eigenvalue of the half row the lane sums in each
    // pass (already multiplied by K / (8 N))
    float V[NM], lam_r[L::PASSES];
#pragma unroll
    for (int m = 0; m < NM; ++m) V[m] = __ldg(p.spec_v + (size_t)m * (NC * NT) + pt);
#pragma unroll
    for (int q = 0; q < L::PASSES; ++q) {
        const int m = q * MP + (lane >> 1);
        lam_r[q] = (lane < 2 * MP && m < NM) ? __ldg(p.spec_lam + m) : 0.f;
    }

    const float rtol = (float)p.rtol, atol = (float)p.atol;
    const float two_pi_r = (float)kTwoPi;
    const float safety_f = (float)p.safety;
    const float inv_n = 1.0f / (float)p.N;
    int par = 0, rpar = 0, spar = 0;                          // buffer parities (uniform over the environment)

#pragma unroll 1
    for (int slot = (int)blockIdx.x / NC; slot < p.n_launch; slot += (int)gridDim.x / NC) {
    const int env = p.env_ids ? p.env_ids[slot] : slot;
    const size_t base = (size_t)env * p.Np + (size_t)pt * 8;
    const size_t krow = (size_t)env * p.Np + (size_t)crank * N;       // this CTA's part of the carried-derivative row (private layout)
    const bool step_mode = p.mode == MODE_STEP;
    const int k_idx = step_mode ? p.step_idx[env] : 0;
    const float act = step_mode ? p.actions[env] : 0.f;
    const bool fsal_in = p.fsal_on && step_mode && p.fsal_valid[env] != 0;

    float y0[8];
    gload8(reinterpret_cast<const float*>(p.phase) + base, y0);
    {
        const int4 a = reinterpret_cast<const int4*>(p.wind + base)[0], b = reinterpret_cast<const int4*>(p.wind + base)[1];
        const int wd[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int r = 0; r < 8; ++r) WD[r * NT + tid] = wd[r];
    }
    unsigned int n_acc = 0, n_rej = 0, n_rhs = 0, n_reuse = 0;
    bool k0_valid = false;                                // K slot 0 holds f(y0) for pulse amplitude amp_k0
    float amp_k0 = 0.f;
    if (fsal_in) {
        float k[8];
        oload8<NT>(reinterpret_cast<const float*>(p.k_fsal) + krow, tid, k);
        ostore8<NT>(K, tid, k);
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
        if (pt == 0) { p.u_out[env] = u; p.n_samples[env] = nI + nII - 1; }
        nseg = 2;
        seg_ts[0] = p.sched_offI + (size_t)k * p.maxI;   seg_nts[0] = nI;  seg_nrec[0] = nI;
        seg_ts[1] = p.sched_offII + (size_t)k * p.maxII; seg_nts[1] = nII; seg_nrec[1] = nII - 1;
        seg_from[0] = seg_from[1] = 0;
        seg_out[0] = 0; seg_out[1] = nI;
        seg_amp[0] = (float)u; seg_amp[1] = 0.f;
        if (nI <= kWarpTs && nII <= kWarpTs) {               // the save times are consulted all the time: keep them on chip
            if (tid < nI) TS[tid] = seg_ts[0][tid];
            if (tid < nII) TS[kWarpTs + tid] = seg_ts[1][tid];
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
    if (tail && wid == 0 && crank == 0) obs_tail_prefetch<float>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos);
    sync_env();

#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
        const double* __restrict__ ts = seg_ts[sg];
        const int n_ts = seg_nts[sg], n_rec = seg_nrec[sg], rec_from = seg_from[sg], out_base = seg_out[sg];
        const float amp = seg_amp[sg];
        {                                     // w0 + pulse, constant over the segment (env.py:254-255, :421-424)
            float c0v[8], stimv[8];
            gload8(reinterpret_cast<const float*>(p.w0) + base, c0v);
            gload8(reinterpret_cast<const float*>(p.stim) + base, stimv);
#pragma unroll
            for (int r = 0; r < 8; ++r) c0v[r] = c0v[r] + amp * stimv[r];
            ostore8<NT>(C0, tid, c0v);
            if (k0_valid) {                      // k1 of this segment from the carried k7: only the pulse term changes
                float k[8];
                oload8<NT>(K, tid, k);
                const float da = amp - amp_k0;
#pragma unroll
                for (int r = 0; r < 8; ++r) k[r] = fmaf(da, stimv[r], k[r]);
                ostore8<NT>(K, tid, k);
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
                float sv[8], cv[8], ys[8];
                {
                    if (s == 6) {                 // y1 = y0 + d1 with d1 summed on its own: k7 = f(y1) exactly (FSAL)
                        float zero[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) zero[r] = 0.f;
                        ostage_argument<NT>(6, K, tid, dt, zero, ys);
#pragma unroll
                        for (int r = 0; r < 8; ++r) ys[r] += y0[r];
                    } else ostage_argument<NT>(s, K, tid, dt, y0, ys);
#pragma unroll
                    for (int r = 0; r < 8; ++r) wsincos(ys[r], &sv[r], &cv[r]);
                    // (2: the stage argument waits in the row of the derivative this stage produces -- not an input of the stage)
                    if (DBSGYM_OCT_RESINCOS == 2) ostore8<NT>(K + kslot(s) * N, tid, ys);
                }
                float2 X[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) X[r] = make_float2(sv[r], cv[r]);
                wht8p(X);
                // ---- mode sums inside the warp, PASSES passes of MP modes: the lane's partial of every mode, P[mode][lane], then
                //      one lane per half row adds 16 of them and the two halves of a mode meet by one shuffle ----
                float2* cp_mine = reinterpret_cast<float2*>(CP) + ((NC > 1 ? 0 : par) * NW + wid) * NMP;
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
                    {
                        const bool live = lane < live_rows;
                        const float4* r4 = reinterpret_cast<const float4*>(Pw + (live ? lane : 0) * L::RS);
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
                        tot = __fmul2_rn(bcast2(lam_r[q]), tot);
                        if (live && !(lane & 1)) cp_mine[q * MP + (lane >> 1)] = tot;
                    }
                    __syncwarp();
                });
                __syncthreads();
                // ---- across the warps: lane = mode, the NW rows in a fixed order ----
                if (NC == 1 || wid == 0) {
                    const float2* cp = reinterpret_cast<const float2*>(CP) + (NC > 1 ? 0 : par) * NW * NMP;
#pragma unroll
                    for (int mb = 0; mb < NM; mb += 32) {
                        const int m = mb + lane;
                        if (m < NM) {
                            float2 acc = cp[m];
#pragma unroll
                            for (int w = 1; w < NW; ++w) acc = __fadd2_rn(acc, cp[w * NMP + m]);
                            if constexpr (NC == 1) reinterpret_cast<float2*>(Cw)[m] = acc;
                            else {
                                float2* mine = reinterpret_cast<float2*>(XC) + (par * NC + crank) * NMP + m;
#pragma unroll
                                for (int c = 0; c < NC; ++c) st_cluster(mine, c, acc);
                            }
                        }
                    }
                }
                if constexpr (NC > 1) {               // ---- across the CTAs of the cluster ----
                    cluster_barrier();
                    const float2* xc = reinterpret_cast<const float2*>(XC) + par * NC * NMP;
#pragma unroll
                    for (int mb = 0; mb < NM; mb += 32) {
                        const int m = mb + lane;
                        if (m < NM) {
                            float2 acc = xc[m];
#pragma unroll
                            for (int c = 1; c < NC; ++c) acc = __fadd2_rn(acc, xc[c * NMP + m]);
                            reinterpret_cast<float2*>(Cw)[m] = acc;
                        }
                    }
                }
                par ^= 1;
                __syncwarp();
                // ---- expansion back to the thread's sector coordinates, then sectors -> images ----
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
                    float ks[8];
                    oload8<NT>(C0, tid, ks);
                    if (DBSGYM_OCT_RESINCOS) {
                        if (DBSGYM_OCT_RESINCOS == 2) oload8<NT>(K + kslot(s) * N, tid, ys);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            float s2, c2;
                            asm volatile("" : "+f"(ys[r]));          // (keep the second evaluation from being merged with the first)
                            wsincos(ys[r], &s2, &c2);
                            ks[r] = fmaf(c2, X[r].x, fmaf(-s2, X[r].y, ks[r]));
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 8; ++r) ks[r] = fmaf(cv[r], X[r].x, fmaf(-sv[r], X[r].y, ks[r]));     // K / (8 N) is folded into lambda
                    }
                    ostore8<NT>(K + kslot(s) * N, tid, ks);
                }
                ++n_rhs;
            }
            have_f0 = true;
            float d1[8];
            {
                float zero[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) zero[r] = 0.f;
                ostage_argument<NT>(6, K, tid, dt, zero, d1);  // y1 - y0, bit-identical to the last stage's increment
            }

            // ---- embedded error estimate and step-size controller (diffrax PIDController, I-only) ----
            float sqr = 0.f;
            {
                float e[8];
                {
                    constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                    constexpr double cf[] = {35.0 / 384 - 1951.0 / 21600, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                             -2187.0 / 6784 + 12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
                    float zero[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) zero[r] = 0.f;
                    olincomb<NT, 6>(K, tid, sl, cf, dt, zero, e);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const float yu0 = y0[r] + two_pi_r * (float)WD[r * NT + tid];
                    const float yu1 = yu0 + d1[r];
                    const float scale = atol + fmaxf(fabsf(yu0), fabsf(yu1)) * rtol;
                    const float qv = __fdividef(e[r], scale);
                    sqr = fmaf(qv, qv, sqr);
                }
            }
            sqr = warp_sum(sqr);
            if (lane == 0) RD[rpar * NW + wid] = sqr;
            __syncthreads();
            sqr = RD[rpar * NW];
#pragma unroll
            for (int w = 1; w < NW; ++w) sqr += RD[rpar * NW + w];
            if constexpr (NC > 1) {
                if (tid < NC) st_cluster(RX + rpar * NC + crank, tid, sqr);
                cluster_barrier();
                sqr = RX[rpar * NC];
#pragma unroll
                for (int c = 1; c < NC; ++c) sqr += RX[rpar * NC + c];
            }
            rpar ^= 1;
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
                    float f0[8], pa[8], pb[8], pc[8];
                    {
                        float kk0[8], k6[8], dm[8];
                        oload8<NT>(K, tid, kk0);
                        oload8<NT>(K + kslot(6) * N, tid, k6);
                        {
                            constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                            constexpr double cf[] = {0.5 * (6025192743.0 / 30085553152.0), 0.5 * (51252292925.0 / 65400821598.0),
                                                     0.5 * (-2691868925.0 / 45128329728.0), 0.5 * (187940372067.0 / 1594534317056.0),
                                                     0.5 * (-1776094331.0 / 19743644256.0), 0.5 * (11237099.0 / 235043384.0)};
                            float zero[8];
#pragma unroll
                            for (int r = 0; r < 8; ++r) zero[r] = 0.f;
                            olincomb<NT, 6>(K, tid, sl, cf, dt, zero, dm);
                        }
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            const float f0r = kk0[r] * dt, f1r = k6[r] * dt, dmr = dm[r], d = d1[r];
                            f0[r] = f0r;
                            pa[r] = 2.f * (f1r - f0r) - 8.f * d + 16.f * dmr;
                            pb[r] = 5.f * f0r - 3.f * f1r + 14.f * d - 32.f * dmr;
                            pc[r] = f1r - 4.f * f0r - 5.f * d + 16.f * dmr;
                        }
                        if (DBSGYM_OCT_DENSE_SMEM) {
                            ostore8<NT>(K + 2 * N, tid, pa); ostore8<NT>(K + 3 * N, tid, pb);
                            ostore8<NT>(K + 4 * N, tid, pc); ostore8<NT>(K + 5 * N, tid, f0);
                        }
                    }
                    float rc[8];                         // recording conductance (env.py:404-412), L1 / L2 resident
                    if (p.weighted_rec) gload8(reinterpret_cast<const float*>(p.rec) + base, rc);
                    else {
#pragma unroll
                        for (int r = 0; r < 8; ++r) rc[r] = 0.f;
                    }
                    // Samples in batches of up to 8: lane sums through the warp's idle partials buffer (as in warp_kernel.cuh), the
                    // warps' totals through the double-buffered table SL, one barrier per batch, warp 0 stores the results
                    const float inv_h = (tnext == tt) ? 0.f : 1.0f / (float)(tnext - tt);
                    constexpr int SS = 68;                                   // words per slot row: 32 float2 + 4 (conflict-free both ways)
                    while (save_idx < n_ts && ts[save_idx] <= tnext) {
                        int first_idx = save_idx, nb = 0;
#pragma unroll 1
                        for (; nb < 8 && save_idx < n_ts && ts[save_idx] <= tnext; ++nb, ++save_idx) {
                            const double tsv = ts[save_idx];
                            float ysmp[8];
                            if (tsv == tnext) {                              // the end point of the sub-step is y1 itself
#pragma unroll
                                for (int r = 0; r < 8; ++r) ysmp[r] = y0[r] + d1[r];
                            } else {
                                const float tau = (float)(tsv - tt) * inv_h;
                                if (DBSGYM_OCT_DENSE_SMEM) {
                                    float c[8];
                                    oload8<NT>(K + 2 * N, tid, ysmp);
                                    oload8<NT>(K + 3 * N, tid, c);
#pragma unroll
                                    for (int r = 0; r < 8; ++r) ysmp[r] = fmaf(ysmp[r], tau, c[r]);
                                    oload8<NT>(K + 4 * N, tid, c);
#pragma unroll
                                    for (int r = 0; r < 8; ++r) ysmp[r] = fmaf(ysmp[r], tau, c[r]);
                                    oload8<NT>(K + 5 * N, tid, c);
#pragma unroll
                                    for (int r = 0; r < 8; ++r) ysmp[r] = fmaf(fmaf(ysmp[r], tau, c[r]), tau, y0[r]);
                                } else {
#pragma unroll
                                for (int r = 0; r < 8; ++r)
                                    ysmp[r] = fmaf(fmaf(fmaf(fmaf(pa[r], tau, pb[r]), tau, pc[r]), tau, f0[r]), tau, y0[r]);
                                }
                            }
                            float st0 = 0.f, st1 = 0.f, sr0 = 0.f, sr1 = 0.f;
#pragma unroll
                            for (int r = 0; r < 8; r += 2) {
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
                            if (lane < 8) reinterpret_cast<float2*>(SL)[(spar * NW + wid) * 8 + lane] = make_float2(a, b);
                        }
                        __syncthreads();
                        if constexpr (NC > 1) {            // the CTAs' totals meet in CTA 0
                            if (wid == 0 && lane < 8) {
                                const float2* sl2 = reinterpret_cast<const float2*>(SL) + spar * NW * 8 + lane;
                                float2 acc = sl2[0];
#pragma unroll
                                for (int w = 1; w < NW; ++w) { acc.x += sl2[w * 8].x; acc.y += sl2[w * 8].y; }
                                st_cluster(reinterpret_cast<float2*>(SX) + (spar * NC + crank) * 8 + lane, 0, acc);
                            }
                            cluster_barrier();
                        }
                        if (wid == 0 && crank == 0) {
                            const int idx = first_idx + lane;
                            if (lane < nb && idx >= rec_from && idx < n_rec) {
                                const float2* sl2 = NC > 1 ? reinterpret_cast<const float2*>(SX) + spar * NC * 8 + lane
                                                           : reinterpret_cast<const float2*>(SL) + spar * NW * 8 + lane;
                                constexpr int NR = NC > 1 ? NC : NW;
                                float a = sl2[0].x, b = sl2[0].y;
#pragma unroll
                                for (int w = 1; w < NR; ++w) { a += sl2[w * 8].x; b += sl2[w * 8].y; }
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
                        spar ^= 1;
                    }
                }
                // ---- accept: y0 <- y1, FSAL k1 <- k7 ----
                {
                    float k6[8];
                    oload8<NT>(K + kslot(6) * N, tid, k6);
                    ostore8<NT>(K, tid, k6);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float y1 = y0[r] + d1[r];
                    const float nwrap = floorf(y1 * 0.15915494309189535f);     // keep the phase wrapped: y = phase + 2 pi wind
                    if (nwrap != 0.0f) {
                        float yw = fmaf(-nwrap, 6.2831854820251465f, y1);
                        yw = fmaf(-nwrap, -1.7484555314695172e-07f, yw);
                        y1 = yw;
                        WD[r * NT + tid] += (int)nwrap;
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
    gstore8(reinterpret_cast<float*>(p.phase) + base, y0);
    {
        int wd[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) wd[r] = WD[r * NT + tid];
        reinterpret_cast<int4*>(p.wind + base)[0] = make_int4(wd[0], wd[1], wd[2], wd[3]);
        reinterpret_cast<int4*>(p.wind + base)[1] = make_int4(wd[4], wd[5], wd[6], wd[7]);
    }
    if (tail && wid == 0 && crank == 0) {
        __syncwarp();                                                // the step's samples (LF) were stored by lanes of this warp
        wobs_tail(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos, LF, u_step);
    }
    if (p.fsal_on) {
        const bool keep_row = k0_valid && amp_k0 == 0.f;
        if (keep_row) {
            float k[8];
            oload8<NT>(K, tid, k);
            ostore8<NT>(reinterpret_cast<float*>(p.k_fsal) + krow, tid, k);
        }
        if (pt == 0) p.fsal_valid[env] = keep_row ? 1 : 0;
    }
    if (pt == 0) {
        if (p.mode == MODE_TRANSIENT) p.head[env] = 0;
        atomicAdd(p.counters + 0, (unsigned long long)n_acc);
        atomicAdd(p.counters + 1, (unsigned long long)n_rej);
        atomicAdd(p.counters + 2, (unsigned long long)n_rhs);
        atomicAdd(p.counters + 3, (unsigned long long)n_reuse);
        if (status) atomicOr(p.status, status);
    }
    sync_env();                                                      // the shared tables (save times, samples) belong to the next environment now
    }
}

}  // namespace dbsgym
