// Fused Kuramoto environment-step kernel (sm_100a).
//
// One WORKER (64 threads at N = 512; a whole CTA, or one of the 8 workers of a multi-worker CTA, or a thread-block
// cluster for N > 4096) integrates ONE environment through every Runge-Kutta sub-step of one SpatialKuramoto.step()
// call (reference environment/env.py:415-454) -- or through the reset transient (env.py:605-612) -- with the
// oscillator state resident on chip, and finishes with the observation tail (window append, beta-power reward,
// episode bookkeeping: env.py:447-454, :638-650, :669-688):
//
//   registers : phases y0 (fp32: wrapped, winding counts in shared memory), w0 + pulse        (8 oscillators / thread)
//   shared    : six slots for the seven Dopri5 stage derivatives (k7 re-uses the slot of the dead k2), the [sin,cos]
//               operand of the coupling contraction (double buffered), the coupling table / sector coefficients
//
// The ODE right-hand side (env.py:252-256) is evaluated through
//     sum_j a_ij sin(th_j - th_i) = cos(th_i) (A sin th)_i - sin(th_i) (A cos th)_i,
// i.e. one [N x N] x [N x 2] contraction per evaluation.  In GRID mode the coupling
// a_ij = f(|dz|,|dx|,|dy|) of a regular neuron grid (utils.py:478-497, env.py:219-229) is a
// 3-level block-Toeplitz operator held as a 2 KB table; it commutes with the three reflections of the grid, so the
// contraction runs block-diagonally in the basis of the 8 parity sectors (GRID_SYM: 1/8 of the multiply-adds, the
// same sum reassociated).  DENSE mode streams an arbitrary symmetric alpha from global memory (generic fallback).
//
// The integrator follows diffrax 0.7.0's Dopri5 + PIDController(I-only) + SaveAt(ts) as
// the reference calls it (env.py:247-249, :260-271); see oracle/diffrax_restated.py for the
// statement of those semantics this kernel is tested against.  DESIGN.md section 3 has the measurements.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dbsgym {

constexpr int kRows = 8;           // oscillators per thread
constexpr int kSlots = 6;          // K slots: stage derivative k_{j+1} lives in slot j, except k7 -> slot 1
                                   // (a[6][1] = b_err[1] = c_mid[1] = 0: k2 is dead once stage 7 starts)
__device__ __forceinline__ int kslot(int j) { return j == 6 ? 1 : j; }
constexpr int kSampleBatch = 16;   // dense-output samples reduced per block barrier

// y-parity: besides z and x the coupling is also even in dy, and a grid line (8 oscillators along y) lives in ONE
// thread, so the y reflection needs no shuffles: with e[j] = v[j] + v[7-j], o[j] = v[j] - v[7-j] (j < 4) the 8x8
// Toeplitz block t[|i-j|] splits into two 4x4 blocks  E_i = sum_j (t[|i-j|] + t[7-i-j]) e[j],
// O_i = sum_j (t[|i-j|] - t[7-i-j]) o[j]:  32 FFMA2 + 32 FADD instead of 64 FFMA2 per (zj,xj) block (96 instead
// of 128 FMA-pipe cycles).  Applied in the fp32 fixed-extent contraction (GEO = 1, 2 and cluster mode).
#ifndef DBSGYM_Y_PARITY
#define DBSGYM_Y_PARITY 1
#endif
constexpr bool kYParity = DBSGYM_Y_PARITY != 0;

enum { MODE_STEP = 0, MODE_TRANSIENT = 1 };
enum { STATUS_MAX_STEPS = 1, STATUS_NAN = 2, STATUS_SCHEDULE = 4 };

struct StepParams {
    int N, Np, B;
    int GX, GZ;                    // GRID coupling: lines along y, GZ*GX lines
    int GY;                        // oscillators per line: 8, or 16 (GEO = 3: two threads per line)
    int weighted_rec;
    int max_steps;
    double k_over_n;
    double rtol, atol, dt0, safety, fmin, fmax, tol_end;
    double act_lo, act_hi;
    const void* table;             // GRID: split-layout table (real)
    const void* alpha;             // DENSE: alpha^T [Np][Np] (real)
    void* phase; int32_t* wind;    // [B][Np]
    const void* w0; const void* stim; const void* rec;
    int mode;
    const int32_t* env_ids; int n_launch;
    // MODE_STEP
    const float* actions; const int32_t* step_idx;
    const int32_t* sched_nI; const int32_t* sched_nII;
    const double* sched_offI; const double* sched_offII;
    int maxI, maxII, n_sched;
    double* lfp_true; double* lfp_rec; int32_t* n_samples; int smax; double* u_out;
    // MODE_TRANSIENT
    const double* ts; int n_ts; void* ring; int W; int32_t* head;
    // bookkeeping
    unsigned long long* counters;  // accepted, rejected, rhs evals
    int32_t* status;
    // fused observation tail (MODE_STEP, beta-power rewards): see obs_tail() below
    int tail_on, tail_kind, tail_nbins;
    double* spec;                  // [B][2 * kTailBins] running DFT bins of the ring (storage order), float64
    const double* tw_full;         // [W][nbins][2] cos, sin of 2*pi*k*m/W
    float* samples_f; float* mirror; float* reward_f; double* reward; uint8_t* done_out; uint8_t* done_dev;
    // host mirror (append-only log, see dbsgym.h: dbsgym_host_mirror): rows of 2 * mir_len floats, every sample stored at
    // column c and c + mir_len, c = per-environment write position mpos advanced modulo mir_len = W + guard
    int mir_len; int32_t* mpos;
    int32_t* step_idx_rw; const int32_t* episode_len;
    int32_t* nsamp_out; int32_t* head_out;      // optional copies of n_samples / the new ring head (mapped host memory)
    // FSAL across segments and launches (fp32 mode): the first stage of a segment is f(y0) with the NEW pulse; the
    // coupling part of it was already evaluated as k7 of the previous accepted sub-step (same y0), so
    // k1 = k7 + (amp_new - amp_old) * stim replaces one RHS evaluation per segment (2 of 32 per step).  The row is
    // carried across launches in k_fsal (valid flag cleared whenever the host rewrites an environment's vectors).
    int fsal_on; void* k_fsal; int32_t* fsal_valid;
    // spectral coupling (CPL_SPECTRAL): eigenvectors per worker thread [64][4 * (RE + RO)] and the eigenvalues, already
    // multiplied by K / (8 N), one per reduction slot [4 * (RE + RO)] (see spectral_contract below)
    const float* spec_v; const float* spec_lam;
    // low-rank coupling (CPL_LOWRANK): eigenvectors [lr_rank][Np] (mode-major, rows padded to a multiple of 4 with zeros)
    // and eigenvalues [lr_rank] of a DENSE alpha (see couple_lowrank_* below)
    const float* lr_v; const float* lr_lam; int lr_rank;
    // sector form (lr_sectors != 0): the oscillators are stored in OCTANT order (position 8 a + g = image g of octant point
    // a, g = 4 my + 2 mz + mx), lr_v holds the eigenvectors of the 8 parity-sector blocks over the octant points only,
    // [lr_rank][Np / 8], modes sorted by sector (counts padded to multiples of 4), lr_soff[9] = first mode of every sector
    int lr_sectors; const int32_t* lr_soff;
    double* trace; int32_t* trace_len; int trace_cap;   // optional recording of the TRUE LFP of every step (evaluation)
    double power_scale, action_cost, threshold, threshold_penalty;
    // cluster mode (one environment = a thread-block cluster of `cluster` CTAs, N > 4096)
    int cluster;                   // CTAs per environment (1 = plain)
    void* cl_operand;              // [B][2][2*Np + kScPad] contraction operand in global memory (L2)
    double* cl_scratch;            // [B][2][cluster][kClSlots] cross-CTA reduction scratch
};

constexpr int kTailBins = 32;      // at most one rfft bin per lane of the tail warp
constexpr int kClTileFloats = 8192;   // cluster mode: 32 KB operand tile staged in shared memory per CTA
constexpr int kClSlots = 2 * 16;   // doubles per CTA and parity in cl_scratch (= 2 * kSampleBatch)

__device__ __forceinline__ void cluster_barrier() {
    // release / acquire at cluster scope: global writes made before the barrier by any CTA of the cluster are
    // visible to every thread of the cluster after it (reads below use ld.global.cg: L1 is not coherent)
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// Dormand-Prince 5(4) coefficients (same values as oracle/diffrax_restated.py): compile-time lists in
// stage_increment() (tableau rows a[s][.]) and at the error-estimate / dense-output sites (b_err, c_mid) below.

constexpr double kTwoPi = 6.283185307179586476925286766559;

// ---- small typed helpers ---------------------------------------------------------------
template <typename real> struct Vec;
template <> struct Vec<float>  { using T = float4;  static constexpr int n = 4; };
template <> struct Vec<double> { using T = double2; static constexpr int n = 2; };

__device__ __forceinline__ void unpack(const float4& v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
__device__ __forceinline__ void unpack(const double2& v, double* o) { o[0] = v.x; o[1] = v.y; }
__device__ __forceinline__ float4 pack4(const float* o) { return make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ double2 pack4(const double* o) { return make_double2(o[0], o[1]); }

// load / store CNT contiguous reals (16-byte aligned) with 128-bit accesses
template <int CNT, typename real>
__device__ __forceinline__ void loadv(const real* __restrict__ p, real* o) {
    using V = typename Vec<real>::T;
    constexpr int n = Vec<real>::n;
#pragma unroll
    for (int q = 0; q < CNT / n; ++q) unpack(reinterpret_cast<const V*>(p)[q], o + q * n);
}
template <int CNT, typename real>
__device__ __forceinline__ void storev(real* __restrict__ p, const real* o) {
    using V = typename Vec<real>::T;
    constexpr int n = Vec<real>::n;
#pragma unroll
    for (int q = 0; q < CNT / n; ++q) reinterpret_cast<V*>(p)[q] = pack4(o + q * n);
}

// Thread-private rows of 8 reals in shared memory (stage derivatives K, winding counts).  Stored INTERLEAVED: the
// q-th 16-byte piece of thread t sits at vector index q * nt + t, so that a warp's 128-bit access covers 512
// contiguous bytes (4 wavefronts).  The natural layout (8 contiguous reals per thread, lane stride 32 B) costs 8
// wavefronts per LDS.128 / STS.128 and 8 per scalar LDS: ncu showed 32 M bank-conflict wavefronts per launch.
template <typename real>
__device__ __forceinline__ void load_row(const real* __restrict__ row, int tid, int nt, real* o) {
    using V = typename Vec<real>::T;
    constexpr int n = Vec<real>::n;
#pragma unroll
    for (int q = 0; q < kRows / n; ++q) unpack(reinterpret_cast<const V*>(row)[q * nt + tid], o + q * n);
}
template <typename real>
__device__ __forceinline__ void store_row(real* __restrict__ row, int tid, int nt, const real* o) {
    using V = typename Vec<real>::T;
    constexpr int n = Vec<real>::n;
#pragma unroll
    for (int q = 0; q < kRows / n; ++q) reinterpret_cast<V*>(row)[q * nt + tid] = pack4(o + q * n);
}

__device__ __forceinline__ float  fma_r(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_r(double a, double b, double c) { return fma(a, b, c); }

// sum_j coef[j] * K[slot[j]] over the thread's 8 oscillators with compile-time coefficient lists: all rows are loaded
// first (independent LDS), then accumulated in ascending j -- the same FMA sequence as a rolled loop over the
// tableau row, without its per-j constant-bank load, zero test, branch and exposed LDS latency.
template <typename real, int NJ>
__device__ __forceinline__ void lincomb(const real* __restrict__ Kb, int Nl, int tid, int nt, const int (&slot)[NJ],
                                        const double (&coef)[NJ], real (&out)[kRows]) {
    real kj[NJ][kRows];
#pragma unroll
    for (int j = 0; j < NJ; ++j) load_row<real>(Kb + slot[j] * Nl, tid, nt, kj[j]);
#pragma unroll
    for (int r = 0; r < kRows; ++r) out[r] = real(0);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const real a = real(coef[j]);
#pragma unroll
        for (int r = 0; r < kRows; ++r) out[r] = fma_r(a, kj[j][r], out[r]);
    }
}

// stage increment sum_j a[s][j] k_{j+1} (k_{j+1} lives in slot j; row 6 skips the zero a[6][1])
template <typename real>
__device__ __forceinline__ void stage_increment(int s, const real* __restrict__ Kb, int Nl, int tid, int nt, real (&inc)[kRows]) {
    switch (s) {
        case 1: { constexpr int sl[] = {0}; constexpr double cf[] = {1.0 / 5}; lincomb<real, 1>(Kb, Nl, tid, nt, sl, cf, inc); break; }
        case 2: { constexpr int sl[] = {0, 1}; constexpr double cf[] = {3.0 / 40, 9.0 / 40}; lincomb<real, 2>(Kb, Nl, tid, nt, sl, cf, inc); break; }
        case 3: { constexpr int sl[] = {0, 1, 2}; constexpr double cf[] = {44.0 / 45, -56.0 / 15, 32.0 / 9};
                  lincomb<real, 3>(Kb, Nl, tid, nt, sl, cf, inc); break; }
        case 4: { constexpr int sl[] = {0, 1, 2, 3};
                  constexpr double cf[] = {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729};
                  lincomb<real, 4>(Kb, Nl, tid, nt, sl, cf, inc); break; }
        case 5: { constexpr int sl[] = {0, 1, 2, 3, 4};
                  constexpr double cf[] = {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656};
                  lincomb<real, 5>(Kb, Nl, tid, nt, sl, cf, inc); break; }
        case 6: { constexpr int sl[] = {0, 2, 3, 4, 5};
                  constexpr double cf[] = {35.0 / 384, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
                  lincomb<real, 5>(Kb, Nl, tid, nt, sl, cf, inc); break; }
        default: {
#pragma unroll
            for (int r = 0; r < kRows; ++r) inc[r] = real(0);
        }
    }
}

// sin / cos of a phase.  fp64 follows the reference literally (theta = fmod(y, 2*pi),
// env.py:253); fp32 phases are kept wrapped, so the library range reduction is exact enough.
#ifndef DBSGYM_PRECISE_SINCOS
// range reduction to [-pi, pi] (two-constant Cody-Waite) + MUFU.SIN / MUFU.COS: abs error ~5e-7
__device__ __forceinline__ void sincos_r(float x, float* s, float* c) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    r = fmaf(-k, -1.7484555314695172e-07f, r);
    *s = __sinf(r); *c = __cosf(r);
}
__device__ __forceinline__ float cos_r(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    r = fmaf(-k, -1.7484555314695172e-07f, r);
    return __cosf(r);
}
#else
__device__ __forceinline__ void sincos_r(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ float  cos_r(float x) { return cosf(x); }
#endif
__device__ __forceinline__ void sincos_r(double x, double* s, double* c) { sincos(fmod(x, kTwoPi), s, c); }
__device__ __forceinline__ double cos_r(double x) { return cos(x); }

// ---- coupling contraction, GRID mode --------------------------------------------------
// Shared table layout: value for block c = dz*GX+dx and offset dy sits at
//   T[((dy / n) * NC + c) * n + dy % n],   n = elements per 16 bytes,
// so that a quarter warp (8 lanes, 8 different dx) reads 8 different 16-byte granules.
template <typename real>
__device__ __forceinline__ void couple_grid(const real* __restrict__ sc, const real* __restrict__ T,
                                            int GZ, int GX, int zi, int xi,
                                            real (&as)[kRows], real (&ac)[kRows]) {
    using V = typename Vec<real>::T;
    constexpr int n = Vec<real>::n;
    const int NC = GZ * GX;
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = real(0); ac[r] = real(0); }
    for (int zj = 0; zj < GZ; ++zj) {
        const int dz = zi > zj ? zi - zj : zj - zi;
        const real* tz = T + dz * GX * n;
        const real* bz = sc + zj * GX * (2 * kRows);
#pragma unroll 2
        for (int xj = 0; xj < GX; ++xj) {
            const int dx = xi > xj ? xi - xj : xj - xi;
            real t[kRows];
#pragma unroll
            for (int q = 0; q < kRows / n; ++q)
                unpack(*reinterpret_cast<const V*>(tz + (q * NC + dx) * n), t + q * n);
            real b[2 * kRows];
            loadv<2 * kRows>(bz + xj * (2 * kRows), b);
#pragma unroll
            for (int yj = 0; yj < kRows; ++yj) {
                const real s = b[2 * yj], c = b[2 * yj + 1];
#pragma unroll
                for (int yi = 0; yi < kRows; ++yi) {
                    const real a = t[yi > yj ? yi - yj : yj - yi];
                    as[yi] = fma_r(a, s, as[yi]);
                    ac[yi] = fma_r(a, c, ac[yi]);
                }
            }
        }
    }
}

// fp32 specialisation: the (sin, cos) accumulators of one oscillator form a packed pair, so one
// Blackwell FFMA2 (fma.rn.f32x2, scalar-broadcast multiplier) does both FMAs in ONE issue slot.
// The scalar loop above is issue-bound (84 % issue utilisation, 70 % FMA pipe in ncu); halving
// the FMA instruction count leaves issue slots for the LDS / address instructions.
template <>
__device__ __forceinline__ void couple_grid<float>(const float* __restrict__ sc, const float* __restrict__ T,
                                                   int GZ, int GX, int zi, int xi,
                                                   float (&as)[kRows], float (&ac)[kRows]) {
    const int NC = GZ * GX;
    float2 acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = make_float2(0.f, 0.f);
    for (int zj = 0; zj < GZ; ++zj) {
        const int dz = zi > zj ? zi - zj : zj - zi;
        const float* tz = T + dz * GX * 4;
        const float* bz = sc + zj * GX * (2 * kRows);
#pragma unroll 2
        for (int xj = 0; xj < GX; ++xj) {
            const int dx = xi > xj ? xi - xj : xj - xi;
            float t[kRows];
            unpack(*reinterpret_cast<const float4*>(tz + dx * 4), t);
            unpack(*reinterpret_cast<const float4*>(tz + (NC + dx) * 4), t + 4);
            float b[2 * kRows];
            loadv<2 * kRows>(bz + xj * (2 * kRows), b);
#pragma unroll
            for (int yj = 0; yj < kRows; ++yj) {
                const float2 scj = make_float2(b[2 * yj], b[2 * yj + 1]);
#pragma unroll
                for (int yi = 0; yi < kRows; ++yi) {
                    const float a = t[yi > yj ? yi - yj : yj - yi];
                    acc[yi] = __ffma2_rn(make_float2(a, a), scj, acc[yi]);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
}

// ---- coupling contraction, GRID_SYM mode: reflection-symmetry reduced -------------------------
// alpha_ij = f(|dz|,|dx|,|dy|) commutes with the reflections z -> GZ-1-z and x -> GX-1-x of the grid.
// In the basis of the four (pz,px) parity sectors  s^p[j'] = sum_g chi_p(g) s[g j']  (j' in the
// fundamental quarter GZ/2 x GX/2, g in {1,Rz,Rx,RzRx}) the operator is block diagonal:
//     y^p[i'] = sum_j' U^p[i'][j'] s^p[j'],   U^p[i'][j'] = sum_k chi_p(k) alpha(i', k j'),
// so one RHS needs 4 x (N/4)^2 instead of N^2 multiply-adds -- the same numbers, reassociated.
// The four lanes of a quad own the four mirror-image grid lines; the sector transform is a 2-stage
// warp-shuffle butterfly (no extra shared memory, no extra barrier).  U^p is not stored: its 8 entries
// per (zj,xj) block are combined on the fly from four rows of the 2 KB Toeplitz table.
template <typename real, int MX = 1, int MZ = 2>   // lane bits that select the x- and the z-mirror image
__device__ __forceinline__ void quad_butterfly(real (&v)[kRows], real sx, real sz, unsigned mask) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) { const real o = __shfl_xor_sync(mask, v[r], MX); v[r] = fma_r(sx, v[r], o); }
#pragma unroll
    for (int r = 0; r < kRows; ++r) { const real o = __shfl_xor_sync(mask, v[r], MZ); v[r] = fma_r(sz, v[r], o); }
}

// the same butterfly on two arrays at once (sin and cos, or the two contraction results): for float the pair shares
// one FFMA2 per stage and row -- the same two fmas, half the FMA-pipe issue slots
template <typename real, int MX, int MZ>
__device__ __forceinline__ void quad_butterfly2(real (&a)[kRows], real (&b)[kRows], real sx, real sz, unsigned mask) {
    quad_butterfly<real, MX, MZ>(a, sx, sz, mask);
    quad_butterfly<real, MX, MZ>(b, sx, sz, mask);
}
template <int MX, int MZ>
__device__ __forceinline__ void quad_butterfly2f(float (&a)[kRows], float (&b)[kRows], float sx, float sz, unsigned mask) {
    const float2 sx2 = make_float2(sx, sx), sz2 = make_float2(sz, sz);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const float2 o = make_float2(__shfl_xor_sync(mask, a[r], MX), __shfl_xor_sync(mask, b[r], MX));
        const float2 v = __ffma2_rn(sx2, make_float2(a[r], b[r]), o);
        a[r] = v.x; b[r] = v.y;
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const float2 o = make_float2(__shfl_xor_sync(mask, a[r], MZ), __shfl_xor_sync(mask, b[r], MZ));
        const float2 v = __ffma2_rn(sz2, make_float2(a[r], b[r]), o);
        a[r] = v.x; b[r] = v.y;
    }
}

template <typename real>
__device__ __forceinline__ void couple_grid_sym(const real* __restrict__ bp, const real* __restrict__ T,
                                                int GZ, int GX, int zq, int xq, real pz, real px,
                                                real (&as)[kRows], real (&ac)[kRows]) {
    using V = typename Vec<real>::T;
    constexpr int n = Vec<real>::n;
    const int NC = GZ * GX, HZ = GZ >> 1, HX = GX >> 1;
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = real(0); ac[r] = real(0); }
    for (int zj = 0; zj < HZ; ++zj) {
        const int dz0 = zq > zj ? zq - zj : zj - zq, dz1 = GZ - 1 - zq - zj;
        const real* t0 = T + dz0 * GX * n;
        const real* t1 = T + dz1 * GX * n;
#pragma unroll 1
        for (int xj = 0; xj < HX; ++xj) {
            const int dx0 = xq > xj ? xq - xj : xj - xq, dx1 = GX - 1 - xq - xj;
            real u[kRows];
#pragma unroll
            for (int q = 0; q < kRows / n; ++q) {
                real a00[n], a01[n], a10[n], a11[n];
                unpack(*reinterpret_cast<const V*>(t0 + (q * NC + dx0) * n), a00);
                unpack(*reinterpret_cast<const V*>(t0 + (q * NC + dx1) * n), a01);
                unpack(*reinterpret_cast<const V*>(t1 + (q * NC + dx0) * n), a10);
                unpack(*reinterpret_cast<const V*>(t1 + (q * NC + dx1) * n), a11);
#pragma unroll
                for (int e = 0; e < n; ++e)
                    u[q * n + e] = fma_r(pz, fma_r(px, a11[e], a10[e]), fma_r(px, a01[e], a00[e]));
            }
            real b[2 * kRows];
            loadv<2 * kRows>(bp + (zj * HX + xj) * (2 * kRows), b);
#pragma unroll
            for (int yj = 0; yj < kRows; ++yj) {
                const real s = b[2 * yj], c = b[2 * yj + 1];
#pragma unroll
                for (int yi = 0; yi < kRows; ++yi) {
                    const real a = u[yi > yj ? yi - yj : yj - yi];
                    as[yi] = fma_r(a, s, as[yi]);
                    ac[yi] = fma_r(a, c, ac[yi]);
                }
            }
        }
    }
}

// fp32: FFMA2 everywhere, and the loop is software pipelined -- the four table rows and the operand line
// of block b+1 are loaded while the 64 FFMA2 of block b execute (ncu on the un-pipelined version: 22 % of
// the stall samples were short-scoreboard waits on the LDS issued right before their first use).
struct SymRows { float4 a00l, a00h, a01l, a01h, a10l, a10h, a11l, a11h; };

__device__ __forceinline__ void sym_load_rows(SymRows& r, const float* __restrict__ T, int GZ, int GX, int NC,
                                              int zq, int xq, int zj, int xj) {
    const int dz0 = zq > zj ? zq - zj : zj - zq, dz1 = GZ - 1 - zq - zj;
    const int dx0 = xq > xj ? xq - xj : xj - xq, dx1 = GX - 1 - xq - xj;
    const float* t00 = T + (dz0 * GX + dx0) * 4;
    const float* t01 = T + (dz0 * GX + dx1) * 4;
    const float* t10 = T + (dz1 * GX + dx0) * 4;
    const float* t11 = T + (dz1 * GX + dx1) * 4;
    r.a00l = *reinterpret_cast<const float4*>(t00); r.a00h = *reinterpret_cast<const float4*>(t00 + NC * 4);
    r.a01l = *reinterpret_cast<const float4*>(t01); r.a01h = *reinterpret_cast<const float4*>(t01 + NC * 4);
    r.a10l = *reinterpret_cast<const float4*>(t10); r.a10h = *reinterpret_cast<const float4*>(t10 + NC * 4);
    r.a11l = *reinterpret_cast<const float4*>(t11); r.a11h = *reinterpret_cast<const float4*>(t11 + NC * 4);
}

__device__ __forceinline__ void sym_combine(const SymRows& r, float2 px2, float2 pz2, float (&u)[kRows]) {
    auto mix = [&](float a00x, float a00y, float a01x, float a01y, float a10x, float a10y, float a11x, float a11y) {
        return __ffma2_rn(pz2, __ffma2_rn(px2, make_float2(a11x, a11y), make_float2(a10x, a10y)),
                          __ffma2_rn(px2, make_float2(a01x, a01y), make_float2(a00x, a00y)));
    };
    const float2 u01 = mix(r.a00l.x, r.a00l.y, r.a01l.x, r.a01l.y, r.a10l.x, r.a10l.y, r.a11l.x, r.a11l.y);
    const float2 u23 = mix(r.a00l.z, r.a00l.w, r.a01l.z, r.a01l.w, r.a10l.z, r.a10l.w, r.a11l.z, r.a11l.w);
    const float2 u45 = mix(r.a00h.x, r.a00h.y, r.a01h.x, r.a01h.y, r.a10h.x, r.a10h.y, r.a11h.x, r.a11h.y);
    const float2 u67 = mix(r.a00h.z, r.a00h.w, r.a01h.z, r.a01h.w, r.a10h.z, r.a10h.w, r.a11h.z, r.a11h.w);
    u[0] = u01.x; u[1] = u01.y; u[2] = u23.x; u[3] = u23.y; u[4] = u45.x; u[5] = u45.y; u[6] = u67.x; u[7] = u67.y;
}

template <>
__device__ __forceinline__ void couple_grid_sym<float>(const float* __restrict__ bp, const float* __restrict__ T,
                                                       int GZ, int GX, int zq, int xq, float pz, float px,
                                                       float (&as)[kRows], float (&ac)[kRows]) {
    const int NC = GZ * GX, HZ = GZ >> 1, HX = GX >> 1, nblk = HZ * HX;
    float2 acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = make_float2(0.f, 0.f);
    const float2 px2 = make_float2(px, px), pz2 = make_float2(pz, pz);
    SymRows rows;
    float bn[2 * kRows];
    sym_load_rows(rows, T, GZ, GX, NC, zq, xq, 0, 0);
    loadv<2 * kRows>(bp, bn);
    int zj = 0, xj = 0;
#pragma unroll 2
    for (int blk = 0; blk < nblk; ++blk) {
        float u[kRows], b[2 * kRows];
        sym_combine(rows, px2, pz2, u);
#pragma unroll
        for (int i = 0; i < 2 * kRows; ++i) b[i] = bn[i];
        if (++xj == HX) { xj = 0; ++zj; }
        if (blk + 1 < nblk) {                       // prefetch block b+1 behind the FMAs of block b
            sym_load_rows(rows, T, GZ, GX, NC, zq, xq, zj, xj);
            loadv<2 * kRows>(bp + (blk + 1) * (2 * kRows), bn);
        }
#pragma unroll
        for (int yj = 0; yj < kRows; ++yj) {
            const float2 scj = make_float2(b[2 * yj], b[2 * yj + 1]);
#pragma unroll
            for (int yi = 0; yi < kRows; ++yi) {
                const float a = u[yi > yj ? yi - yj : yj - yi];
                acc[yi] = __ffma2_rn(make_float2(a, a), scj, acc[yi]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
}

// Compile-time extents (GEO = 1: 8 x 8 x 8 grid, HZ = HX = 4): the 16 (zj,xj) blocks are fully unrolled, so every
// |zq-zj| / |xq-xj| offset is computed once, the mirrored rows are reached through immediates, and the
// scheduler can hoist the loads of block b+1 above the FFMA2 of block b (ncu on the rolled loop: 76 FFMA2 but
// ~108 other instructions per block, mostly address arithmetic and operand copies).
__device__ __forceinline__ float4 ld_table(const float* p, bool gl) {
    return gl ? __ldg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
}

template <int HZ, int HX, bool GL = false, bool GLO = GL>   // HZ == 0: z half-planes given at run time (hz_rt); the x loop is
                                             // always unrolled.  GL: the table lives in global memory (cluster mode); GLO: the operand too
// accumulates the source planes zj in [z0, z1) (HZ > 0: all of them) into acc
__device__ __forceinline__ void couple_grid_sym_fixed_acc(const float* __restrict__ bp, const float* __restrict__ T,
                                                          int zq, int xq, float pz, float px, int hz_rt, int z0, int z1,
                                                          float2 (&acc)[kRows]) {
    const int hz = HZ > 0 ? HZ : hz_rt;
    const int GZ = 2 * hz;
    constexpr int GX = 2 * HX;
    const int NC = GZ * GX;
    const float2 px2 = make_float2(px, px), pz2 = make_float2(pz, pz);
    const float* tz1 = T + ((GZ - 1 - zq) * GX + (GX - 1 - xq)) * 4;      // row (dz1, dx1) of block (0,0); moves by immediates
    auto zplane = [&](int zj) {
        const int dz0 = zq > zj ? zq - zj : zj - zq;
        const float* tz0 = T + dz0 * GX * 4;
#pragma unroll
        for (int xj = 0; xj < HX; ++xj) {
            const int dx0 = xq > xj ? xq - xj : xj - xq;
            SymRows r;
            const float* t00 = tz0 + dx0 * 4;
            const float* t01 = tz0 + (GX - 1 - xq) * 4 - xj * 4;
            const float* t10 = tz1 - zj * GX * 4 + (dx0 - (GX - 1 - xq)) * 4;
            const float* t11 = tz1 - (zj * GX + xj) * 4;
            r.a00l = ld_table(t00, GL); r.a00h = ld_table(t00 + NC * 4, GL);
            r.a01l = ld_table(t01, GL); r.a01h = ld_table(t01 + NC * 4, GL);
            r.a10l = ld_table(t10, GL); r.a10h = ld_table(t10 + NC * 4, GL);
            r.a11l = ld_table(t11, GL); r.a11h = ld_table(t11 + NC * 4, GL);
            float u[kRows], b[2 * kRows];
            sym_combine(r, px2, pz2, u);
            if (GLO) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(bp + (zj * HX + xj) * (2 * kRows)) + q4);
                    b[4 * q4] = v.x; b[4 * q4 + 1] = v.y; b[4 * q4 + 2] = v.z; b[4 * q4 + 3] = v.w;
                }
            } else loadv<2 * kRows>(bp + (zj * HX + xj) * (2 * kRows), b);
            if (kYParity) {
                // operand line layout: pairs (sin, cos) of e[0..3] then o[0..3]; accumulators: E_0..3 then O_0..3
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 be = make_float2(b[2 * j], b[2 * j + 1]);
                    const float2 bo = make_float2(b[8 + 2 * j], b[8 + 2 * j + 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float t = u[i > j ? i - j : j - i], h = u[7 - i - j];
                        const float ce = t + h, co = t - h;
                        acc[i] = __ffma2_rn(make_float2(ce, ce), be, acc[i]);
                        acc[4 + i] = __ffma2_rn(make_float2(co, co), bo, acc[4 + i]);
                    }
                }
            } else {
#pragma unroll
                for (int yj = 0; yj < kRows; ++yj) {
                    const float2 scj = make_float2(b[2 * yj], b[2 * yj + 1]);
#pragma unroll
                    for (int yi = 0; yi < kRows; ++yi) {
                        const float a = u[yi > yj ? yi - yj : yj - yi];
                        acc[yi] = __ffma2_rn(make_float2(a, a), scj, acc[yi]);
                    }
                }
            }
        }
    };
    if (HZ > 0) {
#pragma unroll
        for (int zj = 0; zj < HZ; ++zj) zplane(zj);
    } else {
#pragma unroll 1
        for (int zj = z0; zj < z1; ++zj) zplane(zj);
    }
}

template <int HZ, int HX, bool GL = false>
__device__ __forceinline__ void couple_grid_sym_fixed(const float* __restrict__ bp, const float* __restrict__ T,
                                                      int zq, int xq, float pz, float px, int hz_rt,
                                                      float (&as)[kRows], float (&ac)[kRows]) {
    float2 acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = make_float2(0.f, 0.f);
    couple_grid_sym_fixed_acc<HZ, HX, GL, GL>(bp, T, zq, xq, pz, px, hz_rt, 0, HZ > 0 ? HZ : hz_rt, acc);
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
}

// ---- GRID_SYM with lines of 8*C oscillators (GEO = 3, cubic 16^3 grids): C threads per line ------------------
// Thread (quad q, mirror image, chunk a) owns oscillators y = 8a .. 8a+7 of its line.  Per source block (zj,xj) the
// whole sector row u[0 .. 8C-1] is combined once from the four table rows; the block between target chunk a and
// source chunk a' is Toeplitz with coefficients u[|8(a-a') + i - j|].  The chunk index is uniform per warp
// (tid = (a * quads + q) * 4 + image), so the switch over a does not diverge and every index is a compile-time constant.
template <int C, int A, bool GL>
__device__ __forceinline__ void chunk_row_fma(const float (&u)[8 * C], const float* __restrict__ bline, float2 (&acc)[kRows]) {
#pragma unroll
    for (int ap = 0; ap < C; ++ap) {
        float b[2 * kRows];
        if (GL) {                                     // cluster mode: the operand lives in global memory (L2)
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(bline + ap * (2 * kRows)) + q4);
                b[4 * q4] = v.x; b[4 * q4 + 1] = v.y; b[4 * q4 + 2] = v.z; b[4 * q4 + 3] = v.w;
            }
        } else loadv<2 * kRows>(bline + ap * (2 * kRows), b);
#pragma unroll
        for (int yj = 0; yj < kRows; ++yj) {
            const float2 scj = make_float2(b[2 * yj], b[2 * yj + 1]);
#pragma unroll
            for (int yi = 0; yi < kRows; ++yi) {
                const int d = 8 * (A - ap) + yi - yj;
                const float cf = u[d < 0 ? -d : d];
                acc[yi] = __ffma2_rn(make_float2(cf, cf), scj, acc[yi]);
            }
        }
    }
}

template <int C, bool GL, bool GLO>      // GL: table in global memory; GLO: operand in global memory
__device__ __forceinline__ void couple_grid_sym_chunks_acc(const float* __restrict__ bp, const float* __restrict__ T,
                                                           int GZ, int GX, int zq, int xq, int a, float pz, float px,
                                                           int z0, int z1, float2 (&acc)[kRows]) {
    const int NC = GZ * GX, HX = GX >> 1;
    const float* Tp = T;                                            // piece q of row c at float4 index q * NC + c
#pragma unroll 1
    for (int zj = z0; zj < z1; ++zj) {
        const int dz0 = zq > zj ? zq - zj : zj - zq, dz1 = GZ - 1 - zq - zj;
#pragma unroll 1
        for (int xj = 0; xj < HX; ++xj) {
            const int dx0 = xq > xj ? xq - xj : xj - xq, dx1 = GX - 1 - xq - xj;
            const int c00 = dz0 * GX + dx0, c01 = dz0 * GX + dx1, c10 = dz1 * GX + dx0, c11 = dz1 * GX + dx1;
            float u[8 * C];
#pragma unroll
            for (int q = 0; q < 2 * C; ++q) {
                const float4 a00 = ld_table(Tp + (q * NC + c00) * 4, GL), a01 = ld_table(Tp + (q * NC + c01) * 4, GL);
                const float4 a10 = ld_table(Tp + (q * NC + c10) * 4, GL), a11 = ld_table(Tp + (q * NC + c11) * 4, GL);
                u[4 * q + 0] = fmaf(pz, fmaf(px, a11.x, a10.x), fmaf(px, a01.x, a00.x));
                u[4 * q + 1] = fmaf(pz, fmaf(px, a11.y, a10.y), fmaf(px, a01.y, a00.y));
                u[4 * q + 2] = fmaf(pz, fmaf(px, a11.z, a10.z), fmaf(px, a01.z, a00.z));
                u[4 * q + 3] = fmaf(pz, fmaf(px, a11.w, a10.w), fmaf(px, a01.w, a00.w));
            }
            const float* bline = bp + (zj * HX + xj) * C * (2 * kRows);
            if constexpr (C == 1) chunk_row_fma<C, 0, GLO>(u, bline, acc);
            else if constexpr (C == 2) {
                if (a == 0) chunk_row_fma<C, 0, GLO>(u, bline, acc); else chunk_row_fma<C, 1, GLO>(u, bline, acc);
            } else {
                static_assert(C == 4, "lines of 8, 16 or 32");
                switch (a) {
                    case 0: chunk_row_fma<C, 0, GLO>(u, bline, acc); break;
                    case 1: chunk_row_fma<C, 1, GLO>(u, bline, acc); break;
                    case 2: chunk_row_fma<C, 2, GLO>(u, bline, acc); break;
                    default: chunk_row_fma<C, 3, GLO>(u, bline, acc); break;
                }
            }
        }
    }
}

template <int C, bool GL>
__device__ __forceinline__ void couple_grid_sym_chunks(const float* __restrict__ bp, const float* __restrict__ T,
                                                       int GZ, int GX, int zq, int xq, int a, float pz, float px,
                                                       float (&as)[kRows], float (&ac)[kRows]) {
    float2 acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = make_float2(0.f, 0.f);
    couple_grid_sym_chunks_acc<C, GL, GL>(bp, T, GZ, GX, zq, xq, a, pz, px, 0, GZ >> 1, acc);
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
}

// ---- multi-worker mode (MW): precomputed sector coefficients ------------------------------------------
// ncu on the kernel above (profiles/r01b_*): the contraction is bound by the shared-memory pipe, not by the FMA
// pipe -- per (zj,xj) block a warp issues 8 LDS.128 for the four table rows (2 cycles each: the lanes of a quad
// share the address) and 4 LDS.128 for the operand line (4 cycles each: lanes with equal sector share the
// address, but the LSU only merges ADJACENT lanes -- scripts/microbench/lds_patterns.cu), 32 cycles against 27
// cycles of FFMA2 / FADD.  The sector coefficients u^p[(zq,xq)][(zj,xj)][dy] do not depend on the environment:
// 4 x 16 x 16 x 8 floats = 32 KB.  One CTA therefore hosts kMwEnvs environments (one 64-thread WORKER each,
// named barriers, persistent loop over environments), builds the 32 KB table once and shares it: per block a
// thread now loads its 8 coefficients directly (2 LDS.128, distinct per lane: 4 cycles each) and -- with the
// lanes of a warp ordered sector-major (lane = 8 * sector + quad) -- the operand line with merged 2-cycle loads:
// 16 cycles of shared memory and 21 cycles of FFMA2 / FADD per block (no coefficient combination either).
#ifndef DBSGYM_MW_ENVS
#define DBSGYM_MW_ENVS 8
#endif
constexpr int kMwEnvs = DBSGYM_MW_ENVS;             // environments (workers) per CTA; more than 8 only fit with a
                                                    // single-buffered operand (two barriers per RHS) and <= 96 registers
constexpr int kMwThreads = 64;                      // threads per worker = grid lines of the 8 x 8 x 8 grid
constexpr int kMwUFloat4 = 16 * 2 * kMwThreads;     // [block][half][worker thread] float4 entries of the coefficient table

__device__ __forceinline__ void couple_sym_upre(const float* __restrict__ bp, const float4* __restrict__ U4, int tid_w,
                                                float (&as)[kRows], float (&ac)[kRows]) {
    float2 acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = make_float2(0.f, 0.f);
#ifndef DBSGYM_MW_UNROLL
#define DBSGYM_MW_UNROLL 16       // 16 = fully unrolled; 4 / 8 were tried for the instruction-cache footprint
#endif
    constexpr int kUnroll = DBSGYM_MW_UNROLL;
#pragma unroll kUnroll
    for (int blk = 0; blk < 16; ++blk) {
        float u[kRows], b[2 * kRows];
        unpack(U4[(2 * blk) * kMwThreads + tid_w], u);
        unpack(U4[(2 * blk + 1) * kMwThreads + tid_w], u + 4);
        loadv<2 * kRows>(bp + blk * (2 * kRows), b);
        // operand line layout: pairs (sin, cos) of e[0..3] then o[0..3]; accumulators: E_0..3 then O_0..3
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 be = make_float2(b[2 * j], b[2 * j + 1]);
            const float2 bo = make_float2(b[8 + 2 * j], b[8 + 2 * j + 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float t = u[i > j ? i - j : j - i], h = u[7 - i - j];
                const float ce = t + h, co = t - h;
                acc[i] = __ffma2_rn(make_float2(ce, ce), be, acc[i]);
                acc[4 + i] = __ffma2_rn(make_float2(co, co), bo, acc[4 + i]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
}

// coefficient table of the 8 x 8 x 8 grid, built once per CTA from the 2 KB Toeplitz table in global memory.
// Entry (blk = zj*4+xj, half, t): u^p[dy = 4*half .. 4*half+3] of worker thread t, combined in the same order
// (and therefore to the same bits) as sym_combine() does on the fly.
__host__ __device__ __forceinline__ void mw_decode(int t, int& zq, int& xq, int& sec) {
    const int lane = t & 31, q = (t >> 5) * 8 + (lane & 7);
    sec = lane >> 3; zq = q >> 2; xq = q & 3;
}
__device__ __forceinline__ void mw_build_table(float4* U4, const float* __restrict__ Tg, int tid, int nthreads) {
    constexpr int GX = 8, GZ = 8, NC = GX * GZ;
    for (int e = tid; e < kMwUFloat4; e += nthreads) {
        const int t = e & (kMwThreads - 1), half = (e >> 6) & 1, blk = e >> 7;
        int zq, xq, sec;
        mw_decode(t, zq, xq, sec);
        const int zj = blk >> 2, xj = blk & 3;
        const int dz0 = zq > zj ? zq - zj : zj - zq, dz1 = GZ - 1 - zq - zj;
        const int dx0 = xq > xj ? xq - xj : xj - xq, dx1 = GX - 1 - xq - xj;
        const float px = (sec & 1) ? -1.f : 1.f, pz = (sec & 2) ? -1.f : 1.f;
        const float4* T4 = reinterpret_cast<const float4*>(Tg) + half * NC;
        const float4 a00 = __ldg(T4 + dz0 * GX + dx0), a01 = __ldg(T4 + dz0 * GX + dx1);
        const float4 a10 = __ldg(T4 + dz1 * GX + dx0), a11 = __ldg(T4 + dz1 * GX + dx1);
        float4 u;
        u.x = fmaf(pz, fmaf(px, a11.x, a10.x), fmaf(px, a01.x, a00.x));
        u.y = fmaf(pz, fmaf(px, a11.y, a10.y), fmaf(px, a01.y, a00.y));
        u.z = fmaf(pz, fmaf(px, a11.z, a10.z), fmaf(px, a01.z, a00.z));
        u.w = fmaf(pz, fmaf(px, a11.w, a10.w), fmaf(px, a01.w, a00.w));
        U4[e] = u;
    }
}

// ---- coupling contraction, SPECTRAL mode: the generalised mean-field identity -------------------------------
// alpha = sum_m lambda_m v_m v_m^T with every eigenvector in one of the 8 parity sectors (geometry.py: spectral_factors;
// 34 modes above 1e-10 |lambda_max| for the shipped 8 x 8 x 8 cos kernel, at most RE = 9 in an even-y and RO = 4 in an
// odd-y sector).  Per RHS evaluation and sector:
//     c_m = sum_b v_m[b] X[b]   (projection: weighted order parameters of the sector, (sin, cos) pairs)
//     y[a] = sum_m lambda_m v_m[a] c_m   (expansion)
// Thread (sector sec, fundamental line q) owns X[(q, j)], j < 4, of the sectors (sec, even y) and (sec, odd y) and keeps
// its 4 x (RE + RO) eigenvector entries in REGISTERS for the whole launch (they depend on the thread only, not on the
// environment).  The sum over the 16 lines of a sector goes through shared memory: partials P[sec][m][q] (float2), one
// thread per (sec, m) adds the 16 entries of a row (FADD2 tree), scales by lambda and writes C[sec][m]; two worker
// barriers per evaluation, ~150 instructions per thread instead of ~930 for the sector-block contraction.
template <int RE, int RO> struct SpecLayout {
    static constexpr int R = RE + RO;
    static constexpr int RS = 36;                                            // words per partial row: 16 float2 + 4 (readers conflict-free)
    static constexpr int PS = R * RS + ((16 - (R * RS) % 32) + 32) % 32;     // sector stride == 16 (mod 32): STS.64 of a half warp conflict-free
    static constexpr int CW = (2 * R + 3) & ~3;                              // words of one sector's coefficient row
    static constexpr int CS = (CW % 16 == 0) ? CW + 4 : CW;                  // sector stride: the 4 broadcast reads hit 4 bank groups
    static constexpr int NSUM = 4 * R;                                       // rows to reduce per evaluation
    static constexpr int ROUNDS = (NSUM + 63) / 64;
    static constexpr int floats = 4 * PS + 4 * CS;
    static_assert(PS % 32 == 16 && CS % 4 == 0 && (CS % 32) != 0 && (2 * CS) % 32 != 0 && CS % 32 != (3 * CS) % 32, "bank layout");
};

template <int RE, int RO>
__device__ __forceinline__ void spectral_contract(const float (&scw)[2 * kRows], const float (&Ve)[4][RE], const float (&Vo)[4][RO],
                                                  const float (&lam)[SpecLayout<RE, RO>::ROUNDS], float* __restrict__ Pw,
                                                  float* __restrict__ Cw, int tid_w, int sec, int q, int wid,
                                                  float (&as)[kRows], float (&ac)[kRows]) {
    using L = SpecLayout<RE, RO>;
    // ---- projection partials over this thread's 4 + 4 sector coordinates
    {
        float2 be[4], bo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { be[j] = make_float2(scw[2 * j], scw[2 * j + 1]); bo[j] = make_float2(scw[8 + 2 * j], scw[8 + 2 * j + 1]); }
        float* row = Pw + sec * L::PS + 2 * q;
#pragma unroll
        for (int m = 0; m < RE; ++m) {
            float2 a = __fmul2_rn(make_float2(Ve[0][m], Ve[0][m]), be[0]);
#pragma unroll
            for (int j = 1; j < 4; ++j) a = __ffma2_rn(make_float2(Ve[j][m], Ve[j][m]), be[j], a);
            *reinterpret_cast<float2*>(row + m * L::RS) = a;
        }
#pragma unroll
        for (int m = 0; m < RO; ++m) {
            float2 a = __fmul2_rn(make_float2(Vo[0][m], Vo[0][m]), bo[0]);
#pragma unroll
            for (int j = 1; j < 4; ++j) a = __ffma2_rn(make_float2(Vo[j][m], Vo[j][m]), bo[j], a);
            *reinterpret_cast<float2*>(row + (RE + m) * L::RS) = a;
        }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(wid + 1), "n"(kMwThreads) : "memory");
    // ---- one thread per (sector, mode): sum of the 16 line partials, times lambda * K / (8 N)
#pragma unroll
    for (int rd = 0; rd < L::ROUNDS; ++rd) {
        const int g = tid_w + 64 * rd;
        if (g < L::NSUM) {
            const int sg = g / L::R, mg = g - sg * L::R;
            const float4* r4 = reinterpret_cast<const float4*>(Pw + sg * L::PS + mg * L::RS);
            float2 v[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float4 t = r4[i]; v[2 * i] = make_float2(t.x, t.y); v[2 * i + 1] = make_float2(t.z, t.w); }
#pragma unroll
            for (int w = 8; w > 0; w >>= 1) {
#pragma unroll
                for (int i = 0; i < w; ++i) v[i] = __fadd2_rn(v[i], v[i + w]);
            }
            *reinterpret_cast<float2*>(Cw + sg * L::CS + 2 * mg) = __fmul2_rn(make_float2(lam[rd], lam[rd]), v[0]);
        }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(wid + 1), "n"(kMwThreads) : "memory");
    // ---- expansion back to this thread's sector coordinates: E_0..3 (even y), O_0..3 (odd y)
    {
        float c[L::CW];
        const float4* c4 = reinterpret_cast<const float4*>(Cw + sec * L::CS);
#pragma unroll
        for (int i = 0; i < L::CW / 4; ++i) { const float4 t = c4[i]; c[4 * i] = t.x; c[4 * i + 1] = t.y; c[4 * i + 2] = t.z; c[4 * i + 3] = t.w; }
        float2 acc[kRows];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[i] = __fmul2_rn(make_float2(Ve[i][0], Ve[i][0]), make_float2(c[0], c[1]));
#pragma unroll
            for (int m = 1; m < RE; ++m) acc[i] = __ffma2_rn(make_float2(Ve[i][m], Ve[i][m]), make_float2(c[2 * m], c[2 * m + 1]), acc[i]);
            acc[4 + i] = __fmul2_rn(make_float2(Vo[i][0], Vo[i][0]), make_float2(c[2 * RE], c[2 * RE + 1]));
#pragma unroll
            for (int m = 1; m < RO; ++m)
                acc[4 + i] = __ffma2_rn(make_float2(Vo[i][m], Vo[i][m]), make_float2(c[2 * (RE + m)], c[2 * (RE + m) + 1]), acc[4 + i]);
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
    }
}

// ---- coupling contraction, DENSE mode (alpha^T streamed from global / L2) -----------------
template <typename real>
__device__ __forceinline__ void couple_dense(const real* __restrict__ sc, const real* __restrict__ alphaT,
                                             int Np, int i0, real (&as)[kRows], real (&ac)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = real(0); ac[r] = real(0); }
    const real* col = alphaT + i0;
#pragma unroll 4
    for (int j = 0; j < Np; ++j) {
        real a[kRows];
        loadv<kRows>(col + (size_t)j * Np, a);
        const real s = sc[2 * j], c = sc[2 * j + 1];
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            as[r] = fma_r(a[r], s, as[r]);
            ac[r] = fma_r(a[r], c, ac[r]);
        }
    }
}

// ---- coupling contraction, LOWRANK mode: the generalised mean-field identity for ANY symmetric alpha ---------------
// alpha ~ sum_m lam_m v_m v_m^T over the eigenpairs above a truncation threshold (geometry.lowrank_factors: a smooth
// kernel of a compact neuron cloud has a few dozen of them whatever the ordering of the neurons -- env.py:219-229,
// utils.py:483-497 with shuffle=True).  (alpha s)_i = sum_m v_m[i] C_m,  C_m = lam_m sum_j v_m[j] (s_j, c_j):
// O(N R) instead of O(N^2) per evaluation and R N floats of operator instead of N^2.
// Phase 1, warps over modes (four at a time): V rows stream from global / L2 (coalesced), the (sin, cos) operand comes
// from shared memory, one shuffle reduction per mode and warp.  Phase 2, threads over their 8 oscillators.
// (cluster mode: every CTA sums over its own slice of oscillators [off, off + n) and leaves unscaled partial sums in `out`)
// one group of four modes m0 .. m0 + 3 by one warp: rows of length ld, entries [off, off + n) against the operand x4
__device__ __forceinline__ void lowrank_project_group(const float4* __restrict__ x4, const float* __restrict__ V,
                                                      const float* __restrict__ lam, int m0, int ld, int off, int n,
                                                      float2* __restrict__ out, int lane) {
    const int n4 = n >> 2;
    float2 acc[4];
    const float4* v4[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc[q] = make_float2(0.f, 0.f); v4[q] = reinterpret_cast<const float4*>(V + (size_t)(m0 + q) * ld + off); }
#pragma unroll 4
    for (int j4 = lane; j4 < n4; j4 += 32) {
        const float4 xa = x4[2 * j4], xb = x4[2 * j4 + 1];          // (s, c) of entries 4 j4 .. 4 j4 + 3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(v4[q] + j4);
            acc[q] = __ffma2_rn(make_float2(v.x, v.x), make_float2(xa.x, xa.y), acc[q]);
            acc[q] = __ffma2_rn(make_float2(v.y, v.y), make_float2(xa.z, xa.w), acc[q]);
            acc[q] = __ffma2_rn(make_float2(v.z, v.z), make_float2(xb.x, xb.y), acc[q]);
            acc[q] = __ffma2_rn(make_float2(v.w, v.w), make_float2(xb.z, xb.w), acc[q]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
            acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
        }
    }
    if (lane < 4) {
        const float2 a = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
        const float l = lam ? __ldg(lam + m0 + lane) : 1.f;
        out[m0 + lane] = make_float2(l * a.x, l * a.y);
    }
}
__device__ __forceinline__ void couple_lowrank_project(const float* __restrict__ sc, const float* __restrict__ V,
                                                       const float* __restrict__ lam, int R4, int Np, int off, int n,
                                                       float2* __restrict__ out, int lane, int warp, int nwarps) {
    for (int m0 = warp * 4; m0 < R4; m0 += nwarps * 4)
        lowrank_project_group(reinterpret_cast<const float4*>(sc), V, lam, m0, Np, off, n, out, lane);
}
// sector form: the groups of ALL sectors are dealt out to the warps together (a sector has only 4 .. 32 modes), each against
// the operand of its own sector, XS[sector][point]
__device__ __forceinline__ void couple_lowrank_project_sectors(const float* __restrict__ xs, const float* __restrict__ Z,
                                                               const float* __restrict__ lam, const int32_t* __restrict__ soff,
                                                               int R4, int Np8, int off, int n, float2* __restrict__ out,
                                                               int lane, int warp, int nwarps) {
    for (int m0 = warp * 4; m0 < R4; m0 += nwarps * 4) {
        int sct = 0;
        while (sct < 7 && m0 >= soff[sct + 1]) ++sct;
        lowrank_project_group(reinterpret_cast<const float4*>(xs + 2 * (size_t)sct * n), Z, lam, m0, Np8, off, n, out, lane);
    }
}
__device__ __forceinline__ void couple_lowrank_expand(const float2* __restrict__ Cs, const float* __restrict__ V, int R4, int Np,
                                                      int i0, float (&as)[kRows], float (&ac)[kRows]) {
    float2 acc[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) acc[r] = make_float2(0.f, 0.f);
    const float* col = V + i0;
#pragma unroll 4
    for (int m = 0; m < R4; ++m) {
        float v[kRows];
        loadv<kRows>(col + (size_t)m * Np, v);
        const float2 c = Cs[m];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = __ffma2_rn(make_float2(v[r], v[r]), c, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) { as[r] = acc[r].x; ac[r] = acc[r].y; }
}

// Sector form of the low-rank contraction (regular grids with even extents): alpha commutes with the three reflections, so
// every eigenvector lives in one parity sector and is fixed by its values on the fundamental octant.  A thread owns the 8
// mirror images of one octant point (the oscillators are stored in that order), transforms its (sin, cos) values to the 8
// sector coordinates by an in-register Walsh-Hadamard butterfly, and the mode sums / expansions run over N / 8 octant points
// with sector-wise eigenvectors: 1/8 of the multiply-adds AND 1/8 of the eigenvector traffic of the plain form.
__device__ __forceinline__ void wht8_pairs(float (&a)[kRows], float (&b)[kRows]) {
#pragma unroll
    for (int h = 1; h < 8; h <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if ((i & h) == 0) {
                const float2 u = make_float2(a[i], b[i]), v = make_float2(a[i + h], b[i + h]);
                const float2 p_ = __fadd2_rn(u, v), m_ = __ffma2_rn(v, make_float2(-1.f, -1.f), u);
                a[i] = p_.x; b[i] = p_.y; a[i + h] = m_.x; b[i + h] = m_.y;
            }
        }
    }
}
// expansion in sector form: Y_s[point] = sum over the modes of sector s of Z[m][point] C_m (C already carries lambda / 8)
__device__ __forceinline__ void couple_lowrank_expand_sectors(const float2* __restrict__ Cs, const float* __restrict__ Z,
                                                              const int32_t* __restrict__ soff, int Np8, int point,
                                                              float (&as)[kRows], float (&ac)[kRows]) {
    const float* col = Z + point;
#pragma unroll
    for (int sct = 0; sct < 8; ++sct) {
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f), acc2 = make_float2(0.f, 0.f), acc3 = make_float2(0.f, 0.f);
        const int m1 = soff[sct + 1];
        // four modes per trip (sector counts are multiples of 4), all loads of a trip issued before its arithmetic.  The
        // mode-major table keeps a warp's loads coalesced (128 bytes per mode); a point-major copy -- one wide load per 4
        // modes and thread -- was measured and is slower from N = 4096 up (32 different lines per request).
#pragma unroll 2
        for (int m = soff[sct]; m < m1; m += 4) {
            const float z0 = __ldg(col + (size_t)m * Np8), z1 = __ldg(col + (size_t)(m + 1) * Np8);
            const float z2 = __ldg(col + (size_t)(m + 2) * Np8), z3 = __ldg(col + (size_t)(m + 3) * Np8);
            const float4 ca = *reinterpret_cast<const float4*>(Cs + m), cb = *reinterpret_cast<const float4*>(Cs + m + 2);
            acc0 = __ffma2_rn(make_float2(z0, z0), make_float2(ca.x, ca.y), acc0);
            acc1 = __ffma2_rn(make_float2(z1, z1), make_float2(ca.z, ca.w), acc1);
            acc2 = __ffma2_rn(make_float2(z2, z2), make_float2(cb.x, cb.y), acc2);
            acc3 = __ffma2_rn(make_float2(z3, z3), make_float2(cb.z, cb.w), acc3);
        }
        as[sct] = (acc0.x + acc1.x) + (acc2.x + acc3.x); ac[sct] = (acc0.y + acc1.y) + (acc2.y + acc3.y);
    }
    wht8_pairs(as, ac);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- fused observation tail ------------------------------------------------------------------
// What the separate observation kernel did after every step (env.py:447-454, :638-650, :669-688), done by
// the first warp of the environment's own CTA while the other CTAs of the SM keep the FMA pipe busy:
// append the step's S recorded samples to the ring (and to the host mirror / sample buffer), and update
// the beta-band reward INCREMENTALLY.  The rfft bins are kept in ring STORAGE order (|X_k| is invariant
// under the circular shift when the DFT length equals the ring length), so overwriting ring[pos] changes
//     X_k  by  (x_new - x_old) * e^{j 2 pi k pos / W}:
// S * nbins float64 FMAs per step instead of a W * nbins pass over the window, and no read of the window
// at all.  The bins are (re)initialised from the whole ring by spec_init_kernel after a reset transient.
template <typename real>
__device__ __forceinline__ void obs_tail_prefetch(const StepParams& p, int env, int lane, int S, double* t_delta, int* t_pos) {
    // ring positions of the step's samples and the values they overwrite, fetched at the START of the step so that
    // the tail itself has no dependent global-load chain (S <= 32, checked by the host)
    const int W = p.W;
    const int head = p.head[env];
    if (lane < S) {
        int pos = head + lane;
        if (pos >= W) pos -= W;
        t_pos[lane] = pos;
        t_delta[lane] = (double)(reinterpret_cast<const real*>(p.ring) + (size_t)env * W)[pos];
    }
    if (lane == 0) { t_pos[32] = head; t_pos[33] = p.mirror ? p.mpos[env] : 0; }
}

template <typename real>
__device__ __forceinline__ void obs_tail(const StepParams& p, int env, int lane, int S, double* t_delta, int* t_pos) {
    const int W = p.W, nb = p.tail_nbins;
    real* ring = reinterpret_cast<real*>(p.ring) + (size_t)env * W;
    const int head = t_pos[32];
    const double2* tw = reinterpret_cast<const double2*>(p.tw_full);
    double re = 0.0, im = 0.0;
    if (lane < S) {
        const int pos = t_pos[lane];
        const real v = real(p.lfp_rec[(size_t)env * p.smax + lane]);
        ring[pos] = v;
        if (p.samples_f) p.samples_f[(size_t)env * p.smax + lane] = (float)v;
        if (p.mirror) {                               // zero-copy store into the pinned host mirror (both copies)
            float* mr = p.mirror + (size_t)env * 2 * p.mir_len;
            int c = t_pos[33] + lane;
            if (c >= p.mir_len) c -= p.mir_len;
            mr[c] = (float)v;
            mr[c + p.mir_len] = (float)v;
        }
        t_delta[lane] = (double)v - t_delta[lane];
        if (p.trace) {                                // evaluation trace: theta_mean of the step (env.py:441)
            const int at = p.trace_len[env] + lane;
            if (at < p.trace_cap) p.trace[(size_t)env * p.trace_cap + at] = p.lfp_true[(size_t)env * p.smax + lane];
        }
    }
    __syncwarp();
    if (lane < nb) {
#pragma unroll 4
        for (int j = 0; j < S; ++j) {
            const double2 w = tw[(size_t)t_pos[j] * nb + lane];
            re = fma(t_delta[j], w.x, re);
            im = fma(t_delta[j], w.y, im);
        }
    }
    double pw = 0.0;
    if (lane < nb) {
        double2* sp = reinterpret_cast<double2*>(p.spec) + (size_t)env * kTailBins + lane;
        double2 X = *sp;
        X.x += re; X.y += im;
        *sp = X;
        const double a = X.x / (double)W, b = X.y / (double)W;
        pw = (a * a + b * b) * 2.0;                   // utils.py:21-27: one-sided power of bin k
    }
    pw = warp_sum(pw);
    if (lane == 0) {
        const double au = fabs(p.u_out[env]);
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

// =========================================================================================
// resident CTAs per SM the register allocation is tuned for (fp32: 128 regs/thread)
template <typename real, int MAXT> struct MinBlocks {
#ifndef DBSGYM_MINB64
#define DBSGYM_MINB64 8
#endif
    static constexpr int v = (sizeof(real) == 4 && MAXT <= 128) ? (DBSGYM_MINB64 * 64 / MAXT) : 1;
};

enum { CPL_GRID = 0, CPL_DENSE = 1, CPL_GRID_SYM = 2, CPL_SPECTRAL = 3, CPL_LOWRANK = 4 };
#ifndef DBSGYM_SC_BUFFERS
#define DBSGYM_SC_BUFFERS 2
#endif
constexpr int kScBuffers = DBSGYM_SC_BUFFERS;   // 2: one barrier per RHS evaluation; 1: two barriers, 4 KB less shared memory
constexpr int kMwScBuffers = kMwEnvs > 8 ? 1 : kScBuffers;
constexpr int kScPad = 16;     // reals of padding per operand buffer (GRID_SYM staggers its 4 sectors by 16 B)

// shared memory of one worker in multi-worker mode (fp32): K slots, double-buffered operand, winding counts,
// reduction scratch, observation-tail scratch
__host__ __device__ inline size_t step_smem_bytes_worker(int Np, int op_floats = -1) {
    const int nwarps = kMwThreads / 32;
    // op_floats: size of the contraction scratch (default: the double-buffered [sin, cos] operand; spectral: partials + coefficients)
    const int opf = op_floats >= 0 ? op_floats : kMwScBuffers * (2 * Np + kScPad);
    return (size_t)(kSlots * Np + opf) * sizeof(float) + (size_t)Np * sizeof(int) +
           (size_t)(nwarps * kSampleBatch * 2 + nwarps + 32) * sizeof(double) + 36 * sizeof(int);
}

// GEO = 1: the grid extents are the compile-time constants 8 x 8 x 8 (every shipped config), which
// turns the table / operand address arithmetic of the contraction into immediates.  GEO = 2: gx = 8 at compile
// time, gz at run time (the 8 x 8 x gz grids of the oscillator-count sweep): x loop unrolled, z loop rolled.
#ifdef DBSGYM_MAXNREG
#define DBSGYM_KERNEL_BOUNDS(real, MAXT) __maxnreg__(DBSGYM_MAXNREG)
#else
#define DBSGYM_KERNEL_BOUNDS(real, MAXT) __launch_bounds__(MAXT, MinBlocks<real, MAXT>::v)
#endif
// CL = 1: one environment is integrated by a thread-block CLUSTER of p.cluster CTAs (N > 4096: the stage
// derivatives no longer fit one SM).  Each CTA keeps its share of the RK state in its own shared memory; the
// contraction operand is exchanged through a double-buffered global (L2-resident) buffer, the coupling table is
// read through the read-only path, and the barrier per RHS evaluation as well as the error-norm / LFP reductions
// become cluster-scope (barrier.cluster release/acquire + a few doubles of global scratch).
template <typename real, int CPL, int MAXT, int GEO = 0, int CL = 0, int EPC = 1, int RE = 1, int RO = 1>
__global__ void DBSGYM_KERNEL_BOUNDS(real, MAXT) step_kernel(const StepParams p) {
    constexpr bool LR = CPL == CPL_LOWRANK;           // low-rank form of a DENSE operator: same thread layout and operand
    constexpr bool DENSE = CPL == CPL_DENSE || LR;
    static_assert(!LR || (sizeof(real) == 4 && EPC == 1 && GEO == 0), "low-rank coupling: fp32, one CTA or one cluster per environment");
    constexpr bool SPEC = CPL == CPL_SPECTRAL;      // spectral contraction: multi-worker hosting, parity-sector thread layout
    constexpr bool SYM = CPL == CPL_GRID_SYM || SPEC;
    constexpr bool MW = EPC > 1;       // multi-worker mode: EPC environments per CTA, one 64-thread worker each
    using SL = SpecLayout<RE, RO>;
    static_assert(!SPEC || (MW && GEO == 1 && sizeof(real) == 4 && CL == 0), "spectral mode: fp32, 8 x 8 x 8 grid, multi-worker");
    static_assert(CL == 0 || LR || (CPL == CPL_GRID_SYM && (GEO == 2 || GEO == 3 || GEO == 4) && sizeof(real) == 4),
                  "cluster mode: fp32 GRID_SYM with gx = 8 or with lines of 16 / 32");
    static_assert(!MW || (SYM && GEO == 1 && sizeof(real) == 4 && CL == 0 && kYParity && MAXT == EPC * kMwThreads),
                  "multi-worker mode: fp32 GRID_SYM on the 8 x 8 x 8 grid with y parity");
    const int GZ = GEO == 1 ? 8 : p.GZ, GX = (GEO == 1 || GEO == 2) ? 8 : p.GX;      // GEO == 2: gx = 8 fixed, gz at run time
    constexpr int CH = GEO == 3 ? 2 : GEO == 4 ? 4 : 1;                  // GEO == 3 / 4: lines of 16 / 32, 2 / 4 threads (chunks) per line
    static_assert((GEO != 3 && GEO != 4) || (CPL == CPL_GRID_SYM && sizeof(real) == 4 && EPC == 1), "GEO 3 / 4: fp32 GRID_SYM");
    const int GY = CH * kRows;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int wid = MW ? (int)(threadIdx.x / kMwThreads) : 0;         // worker of this thread
    const int tid = MW ? (int)(threadIdx.x % kMwThreads) : (int)threadIdx.x, nt = MW ? kMwThreads : (int)blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = (nt + 31) >> 5;
    const int Np = p.Np;
    const int NC_ = CL ? p.cluster : 1;                   // CTAs per environment
    const int crank = CL ? (int)(blockIdx.x % NC_) : 0;   // rank of this CTA in its cluster (1-D grid, cluster dims (C,1,1))
    const int Nl = CL ? nt * kRows : Np;                  // oscillators whose state lives in THIS CTA's (worker's) shared memory
    const int tab = LR ? 2 * p.lr_rank : (DENSE || CL || MW) ? 0 : GZ * GX * GY;     // (LR: the mode coefficients C live there)
    constexpr bool CLG = CL != 0 && !LR;                  // cluster mode with the operand of the whole environment in global memory
    const int scsz = 2 * ((CL && LR) ? Nl : Np) + kScPad;  // (low-rank cluster mode: every CTA keeps the operand of its own oscillators)

    // shared memory of this CTA (MW: the coefficient table, then one such block per worker)
    unsigned char* wsm = smem_raw + (SPEC ? (size_t)wid * step_smem_bytes_worker(Np, SL::floats)
                                          : MW ? kMwUFloat4 * sizeof(float4) + (size_t)wid * step_smem_bytes_worker(Np) : 0);
    real* K = reinterpret_cast<real*>(wsm);               // [kSlots][Nl] stage derivatives f(y_s), thread-private slots
    real* SCs = K + kSlots * Nl;                          // [kScBuffers][scsz] (sin, cos) contraction operand (not in cluster mode)
    constexpr int SCB = MW ? kMwScBuffers : kScBuffers;   // operand buffers
    real* Ts = SCs + (CLG ? 0 : SPEC ? SL::floats : SCB * scsz);   // [tab]  (SPEC: SCs holds the partials P and the coefficients C)
    real* RC = Ts + tab;                                  // [Nl] recording conductance (thread-private slots; MW: read from global)
    int* WD = reinterpret_cast<int*>(RC + (MW ? 0 : Nl)); // [Nl] fp32 mode: winding counts, y = phase + 2*pi*wind
    double* part = reinterpret_cast<double*>(WD + Nl);    // [nwarps][kSampleBatch][2]
    double* red = part + nwarps * kSampleBatch * 2;       // [nwarps]
    double* t_delta = red + nwarps;                       // [32] observation tail scratch
    int* t_pos = reinterpret_cast<int*>(t_delta + 32);    // [36]
    // cluster mode: operand tile [4 sectors][TZ planes] (kClTileFloats + 16 floats), 16-byte aligned
    float* cl_tile = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(t_pos + 36) + 15) & ~uintptr_t(15));
    const real* T = CLG ? reinterpret_cast<const real*>(p.table) : Ts;
    const float4* U4 = reinterpret_cast<const float4*>(smem_raw);       // MW: [16][2][64] sector coefficients

    // barrier over the threads that integrate one environment
    auto env_sync = [&]() {
        if (MW) asm volatile("bar.sync %0, %1;" ::"r"(wid + 1), "n"(kMwThreads) : "memory");
        else __syncthreads();
    };

    const int tid_g = crank * nt + tid;                   // thread index within the environment
    const int k0 = tid * kRows;                           // private slot in K / RC / WD (and in the plain operand)
    // grid line owned by this thread.  GRID: line = tid.  GRID_SYM: the four mirror images (z,x), (z,X-x), (Z-z,x),
    // (Z-z,X-x) of fundamental line q are owned by the four lanes of a quad (q = tid / 4, image = tid & 3), in
    // multi-worker mode by lanes 8 apart (lane = 8 * image + q % 8: lanes with equal image are adjacent).
    int zi = 0, xi = 0, zq = 0, xq = 0, img = 0, qline = 0, chunk = 0;
    real sgn_x = real(1), sgn_z = real(1);
    if (SYM) {
        const int HX = GX >> 1;
        if (MW) mw_decode(tid, zq, xq, img);
        else if (CH > 1) {                                // thread (chunk * quads + q) * 4 + image of the environment: the chunk is warp-uniform
            const int nq4 = (GZ >> 1) * HX * 4;
            chunk = tid_g / nq4;
            const int r4 = tid_g % nq4, q = r4 >> 2;
            zq = q / HX; xq = q % HX; img = r4 & 3;
        } else { const int q = tid_g >> 2; zq = q / HX; xq = q % HX; img = tid & 3; }
        qline = zq * HX + xq;
        zi = (img & 2) ? GZ - 1 - zq : zq;
        xi = (img & 1) ? GX - 1 - xq : xq;
        sgn_x = (img & 1) ? real(-1) : real(1);
        sgn_z = (img & 2) ? real(-1) : real(1);
    } else if (!DENSE) {
        zi = tid / GX; xi = tid % GX;
    }
    const int i0 = DENSE ? tid_g * kRows : (zi * GX + xi) * GY + chunk * kRows; // first oscillator index in the global arrays
    const unsigned wmask = __activemask();
    // operand slot written by this thread: plain = own line; GRID_SYM = sector img, line q
    const int sec_stride = (GZ >> 1) * (GX >> 1) * CH * 2 * kRows + (int)(16 / sizeof(real));
    const int sc_sector = SYM ? img * sec_stride : 0;
    const int sc_slot = SYM ? sc_sector + (qline * CH + chunk) * 2 * kRows : 2 * k0;
    constexpr int BMX = MW ? 8 : 1, BMZ = MW ? 16 : 2;     // lane bits of the x / z mirror image (quad_butterfly)

    // spectral mode: this thread's eigenvector entries and the eigenvalues of the reduction rows it sums -- registers,
    // loaded once per launch (they depend on the thread only)
    float Ve[4][RE], Vo[4][RO], lam_r[SL::ROUNDS];
    if constexpr (SPEC) {
        const float* v = p.spec_v + (size_t)tid * 4 * SL::R;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int m = 0; m < RE; ++m) Ve[j][m] = __ldg(v + j * SL::R + m);
#pragma unroll
            for (int m = 0; m < RO; ++m) Vo[j][m] = __ldg(v + j * SL::R + RE + m);
        }
#pragma unroll
        for (int rd = 0; rd < SL::ROUNDS; ++rd) lam_r[rd] = (tid + 64 * rd < SL::NSUM) ? __ldg(p.spec_lam + tid + 64 * rd) : 0.f;
    }
    if (SPEC) {
    } else if (MW) {                                      // the coefficient table, once per CTA
        mw_build_table(reinterpret_cast<float4*>(smem_raw), reinterpret_cast<const float*>(p.table), (int)threadIdx.x, (int)blockDim.x);
        __syncthreads();
    } else if (!DENSE && !CL) {
        using V = typename Vec<real>::T;
        constexpr int vn = Vec<real>::n;
        const V* tg = reinterpret_cast<const V*>(p.table);
        for (int i = tid; i < tab / vn; i += nt) reinterpret_cast<V*>(Ts)[i] = tg[i];
    }

    // ---- environments of this CTA (worker): exactly one, or in multi-worker mode a persistent loop --------
    int slot = MW ? wid * (int)gridDim.x + (int)blockIdx.x : (CL ? (int)(blockIdx.x / NC_) : (int)blockIdx.x);
    const int slot_stride = MW ? (int)gridDim.x * EPC : 0;
#pragma unroll 1
    for (; slot < p.n_launch; slot += slot_stride) {      // (in cluster mode whole clusters leave together)
    const int env = p.env_ids ? p.env_ids[slot] : slot;
    const size_t base = (size_t)env * Np;
    real* SC = CLG ? reinterpret_cast<real*>(p.cl_operand) + (size_t)env * 2 * scsz : SCs;
    double* cls = CL ? p.cl_scratch + (size_t)env * 2 * NC_ * kClSlots : nullptr;
    int cl_par = 0, lr_par = 0;

    // Register diet: only y0 and (per segment) c0 = w0 + amp * stim stay in registers across the
    // contraction; the recording conductance and the winding counts live in thread-private shared slots.
    real y0[kRows];
    loadv<kRows>(reinterpret_cast<const real*>(p.phase) + base + i0, y0);
    {
        real rc[kRows];
        if (p.weighted_rec) loadv<kRows>(reinterpret_cast<const real*>(p.rec) + base + i0, rc);
        else {
#pragma unroll
            for (int r = 0; r < kRows; ++r) rc[r] = real(0);
        }
        if (!MW) storev<kRows>(RC + k0, rc);
#pragma unroll
        for (int r = 0; r < kRows; ++r) WD[r * nt + tid] = sizeof(real) == 4 ? p.wind[base + i0 + r] : 0;   // [r][thread]: conflict-free
    }
    env_sync();

    constexpr bool YPAR = kYParity && SYM && (GEO == 1 || GEO == 2) && sizeof(real) == 4;     // the paths that call couple_grid_sym_fixed
    const real kn = real(SYM ? (YPAR ? 0.125 : 0.25) * p.k_over_n : p.k_over_n);
    const real rtol = real(p.rtol), atol = real(p.atol);
    const real two_pi_r = real(kTwoPi);
    unsigned int n_acc = 0, n_rej = 0, n_rhs = 0, n_reuse = 0;
    constexpr bool REUSE = sizeof(real) == 4;             // fp64 replays the reference's evaluation sequence exactly
    bool k0_valid = false;                                // K slot 0 holds f(y0) for pulse amplitude amp_k0
    real amp_k0 = real(0);
    if (REUSE && p.fsal_on && p.mode == MODE_STEP && p.fsal_valid[env]) {
        real k[kRows];
        load_row<real>(reinterpret_cast<const real*>(p.k_fsal) + base + (CL ? crank * Nl : 0), tid, nt, k);
        store_row<real>(K, tid, nt, k);
        k0_valid = true;
    }
    int status = 0;
    int pbuf = 0;

    // ---- segment programme -------------------------------------------------------------
    int nseg;
    const double* seg_ts[2];
    int seg_nts[2], seg_nrec[2], seg_from[2], seg_out[2];
    real seg_amp[2];
    if (p.mode == MODE_STEP) {
        int k = p.step_idx[env];
        if (k < 0 || k >= p.n_sched) { status |= STATUS_SCHEDULE; k = k < 0 ? 0 : p.n_sched - 1; }
        const int nI = p.sched_nI[k], nII = p.sched_nII[k];
        // env.py:389-393 rescale_action, env.py:419
        const double a = (double)p.actions[env];
        const double u = p.act_lo + ((p.act_hi - p.act_lo) * (a - (-1.0))) / (1.0 - (-1.0));
        if (tid == 0 && crank == 0) { p.u_out[env] = u; p.n_samples[env] = nI + nII - 1; }
        nseg = 2;
        seg_ts[0] = p.sched_offI + (size_t)k * p.maxI;   seg_nts[0] = nI;  seg_nrec[0] = nI;
        seg_ts[1] = p.sched_offII + (size_t)k * p.maxII; seg_nts[1] = nII; seg_nrec[1] = nII - 1;
        seg_from[0] = seg_from[1] = 0;
        seg_out[0] = 0; seg_out[1] = nI;
        seg_amp[0] = real(u); seg_amp[1] = real(0);
    } else {
        nseg = 1;
        seg_ts[0] = p.ts; seg_nts[0] = p.n_ts; seg_nrec[0] = p.n_ts - 1;
        seg_from[0] = seg_nrec[0] > p.W ? seg_nrec[0] - p.W : 0;
        seg_out[0] = 0; seg_amp[0] = real(0);
        seg_ts[1] = nullptr; seg_nts[1] = seg_nrec[1] = seg_from[1] = seg_out[1] = 0; seg_amp[1] = real(0);
    }

    const bool tail = p.tail_on && p.mode == MODE_STEP && crank == 0 && warp == 0;
    if (tail) obs_tail_prefetch<real>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos);

#pragma unroll 1
    for (int sg = 0; sg < nseg; ++sg) {
        const double* __restrict__ ts = seg_ts[sg];
        const int n_ts = seg_nts[sg], n_rec = seg_nrec[sg], rec_from = seg_from[sg], out_base = seg_out[sg];
        const real amp = seg_amp[sg];
        real c0[kRows];                       // w0 + pulse, constant over the segment (env.py:254-255, :421-424)
        {
            real w0[kRows], stim[kRows];
            loadv<kRows>(reinterpret_cast<const real*>(p.w0) + base + i0, w0);
            loadv<kRows>(reinterpret_cast<const real*>(p.stim) + base + i0, stim);
#pragma unroll
            for (int r = 0; r < kRows; ++r) c0[r] = w0[r] + amp * stim[r];
            if (REUSE && k0_valid) {             // k1 of this segment from the carried k7: only the pulse term changes
                real k[kRows];
                load_row<real>(K, tid, nt, k);
                const real da = amp - amp_k0;
#pragma unroll
                for (int r = 0; r < kRows; ++r) k[r] = fma_r(da, stim[r], k[r]);
                store_row<real>(K, tid, nt, k);
            }
        }
        const double T_end = ts[n_ts - 1];
        double t = 0.0;
        double tnext = fmin(p.dt0, T_end);
        int save_idx = 0;
        int attempts = 0;
        // each forward() starts without FSAL data (env.py:260): stage 0 is evaluated -- or taken from the carried row,
        // which still counts as one (logical) RHS evaluation of the reference
        bool have_f0 = REUSE && k0_valid;
        if (have_f0) { ++n_rhs; ++n_reuse; }
        k0_valid = false;

        while (t < T_end) {
            if (++attempts > p.max_steps) { status |= STATUS_MAX_STEPS; break; }
            const double dt_d = tnext - t;
            const real dt = real(dt_d);
            real d1[kRows];

            // ---- stages: s = 0 is f(y0) (only when no FSAL value is carried) -----------
#pragma unroll 1
            for (int s = have_f0 ? 1 : 0; s < 7; ++s) {
                real inc[kRows];
                stage_increment<real>(s, K, Nl, tid, nt, inc);
                real sv[kRows], cv[kRows];
                real scw[2 * kRows];
                {
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        inc[r] *= dt;
                        sincos_r(y0[r] + inc[r], &sv[r], &cv[r]);
                    }
                    if (SYM) {                                   // to the parity-sector basis (quad butterfly)
                        real ts_[kRows], tc_[kRows];
#pragma unroll
                        for (int r = 0; r < kRows; ++r) { ts_[r] = sv[r]; tc_[r] = cv[r]; }
                        if constexpr (sizeof(real) == 4)
                            quad_butterfly2f<BMX, BMZ>(reinterpret_cast<float(&)[kRows]>(ts_), reinterpret_cast<float(&)[kRows]>(tc_),
                                                       (float)sgn_x, (float)sgn_z, wmask);
                        else quad_butterfly2<real, BMX, BMZ>(ts_, tc_, sgn_x, sgn_z, wmask);
                        if (YPAR) {                              // y reflection: even / odd combinations of the line
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                scw[2 * r] = ts_[r] + ts_[7 - r];         scw[2 * r + 1] = tc_[r] + tc_[7 - r];
                                scw[8 + 2 * r] = ts_[r] - ts_[7 - r];     scw[8 + 2 * r + 1] = tc_[r] - tc_[7 - r];
                            }
                        } else {
#pragma unroll
                            for (int r = 0; r < kRows; ++r) { scw[2 * r] = ts_[r]; scw[2 * r + 1] = tc_[r]; }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < kRows; ++r) { scw[2 * r] = sv[r]; scw[2 * r + 1] = cv[r]; }
                    }
                    if constexpr (LR) {
                        if (p.lr_sectors) {                          // sector coordinates of this thread's octant point, XS[sector][point]
                            real xs_[kRows], xc_[kRows];
#pragma unroll
                            for (int r = 0; r < kRows; ++r) { xs_[r] = sv[r]; xc_[r] = cv[r]; }
                            wht8_pairs(reinterpret_cast<float(&)[kRows]>(xs_), reinterpret_cast<float(&)[kRows]>(xc_));
                            float2* xsm = reinterpret_cast<float2*>(SC + pbuf * scsz);
#pragma unroll
                            for (int r = 0; r < kRows; ++r) xsm[r * nt + tid] = make_float2((float)xs_[r], (float)xc_[r]);
                        } else storev<2 * kRows>(SC + pbuf * scsz + sc_slot, scw);
                    } else if (!SPEC) storev<2 * kRows>(SC + pbuf * scsz + sc_slot, scw);
                }
                if (CLG) cluster_barrier(); else if (!SPEC) env_sync();
                real as[kRows], ac[kRows];
                if constexpr (LR) {
                    float2* Cs = reinterpret_cast<float2*>(Ts);
                    const float* opnd = reinterpret_cast<const float*>(SC + pbuf * scsz);
                    // one pass of phase 1 over the oscillators (plain form) or over the octant points of a sector (sector form):
                    // all modes [m_lo, m_hi) against `n` operand entries at `x`, eigenvector rows of length `ld` starting at `off`
                    auto project = [&](const float* lamp, float2* out) {
                        if (p.lr_sectors)
                            couple_lowrank_project_sectors(opnd, p.lr_v, lamp, p.lr_soff, p.lr_rank, Np >> 3, crank * nt, nt, out,
                                                           lane, warp, nwarps);
                        else couple_lowrank_project(opnd, p.lr_v, lamp, p.lr_rank, Np, crank * Nl, Nl, out, lane, warp, nwarps);
                    };
                    if constexpr (CL != 0) {
                        // every CTA of the cluster sums over its own oscillators into its OWN shared memory; after one cluster
                        // barrier every CTA reads the partial sums of all ranks through distributed shared memory (mapa +
                        // ld.shared::cluster) and adds them in the same order, so all CTAs expand with identical coefficients.
                        // Two buffers by parity: a rank rewrites a buffer two barriers after the others finished reading it.
                        float2* part_s = reinterpret_cast<float2*>(cl_tile) + lr_par * p.lr_rank;
                        project(nullptr, part_s);
                        cluster_barrier();
                        for (int m = tid; m < p.lr_rank; m += nt) {
                            const uint32_t laddr = (uint32_t)__cvta_generic_to_shared(part_s + m);
                            float2 a = make_float2(0.f, 0.f);
#pragma unroll 4
                            for (int r = 0; r < NC_; ++r) {
                                uint32_t raddr;
                                float2 b;
                                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(laddr), "r"(r));
                                asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(b.x), "=f"(b.y) : "r"(raddr) : "memory");
                                a.x += b.x; a.y += b.y;
                            }
                            const float l = __ldg(p.lr_lam + m);
                            Cs[m] = make_float2(l * a.x, l * a.y);
                        }
                        lr_par ^= 1;
                    } else project(p.lr_lam, Cs);
                    env_sync();
                    if (p.lr_sectors)
                        couple_lowrank_expand_sectors(Cs, p.lr_v, p.lr_soff, Np >> 3, tid_g, reinterpret_cast<float(&)[kRows]>(as),
                                                      reinterpret_cast<float(&)[kRows]>(ac));
                    else
                        couple_lowrank_expand(Cs, p.lr_v, p.lr_rank, Np, i0, reinterpret_cast<float(&)[kRows]>(as),
                                              reinterpret_cast<float(&)[kRows]>(ac));
                } else if (DENSE) couple_dense<real>(SC + pbuf * scsz, reinterpret_cast<const real*>(p.alpha), Np, i0, as, ac);
                else if (SYM) {
                    if constexpr (SPEC) {
                        spectral_contract<RE, RO>(reinterpret_cast<const float(&)[2 * kRows]>(scw), Ve, Vo, lam_r,
                                                  reinterpret_cast<float*>(SCs), reinterpret_cast<float*>(SCs) + 4 * SL::PS, tid, img,
                                                  qline, wid, reinterpret_cast<float(&)[kRows]>(as), reinterpret_cast<float(&)[kRows]>(ac));
                    } else if constexpr (CLG) {
                        // Cluster mode: the operand of the whole environment sits in global memory (L2).  Every CTA
                        // stages it through its own shared memory in tiles of whole source z-planes (one cooperative,
                        // coalesced copy per tile and CTA instead of every warp fetching every line from L2), the
                        // contraction then reads shared memory; only the coupling table still comes through L1.
                        const int HZc = GZ >> 1;
                        const int plane_f = (GX >> 1) * CH * 2 * kRows;             // floats per source z-plane and sector
                        int TZ = kClTileFloats / (4 * plane_f);
                        if (TZ < 1) TZ = 1;
                        const int tsec = TZ * plane_f + 4;                          // sector stride in the tile (staggered banks)
                        const float* opg = reinterpret_cast<const float*>(SC + pbuf * scsz);
                        float2 acc2[kRows];
#pragma unroll
                        for (int r = 0; r < kRows; ++r) acc2[r] = make_float2(0.f, 0.f);
#pragma unroll 1
                        for (int z0 = 0; z0 < HZc; z0 += TZ) {
                            const int tz = HZc - z0 < TZ ? HZc - z0 : TZ;
                            const int n4 = tz * plane_f / 4;                         // float4 per sector in this tile
                            __syncthreads();                                         // the previous tile has been consumed
                            for (int i = tid; i < 4 * n4; i += nt) {
                                const int sp = i / n4, r4 = i - sp * n4;
                                reinterpret_cast<float4*>(cl_tile + sp * tsec)[r4] =
                                    __ldcg(reinterpret_cast<const float4*>(opg + sp * sec_stride + z0 * plane_f) + r4);
                            }
                            __syncthreads();
                            const float* bpt = cl_tile + img * tsec - z0 * plane_f;  // so that plane zj is found at zj * plane_f
                            if constexpr (GEO == 3 || GEO == 4)
                                couple_grid_sym_chunks_acc<CH, true, false>(bpt, reinterpret_cast<const float*>(T), GZ, GX, zq, xq, chunk,
                                                                            (float)sgn_z, (float)sgn_x, z0, z0 + tz, acc2);
                            else
                                couple_grid_sym_fixed_acc<0, 4, true, false>(bpt, reinterpret_cast<const float*>(T), zq, xq, (float)sgn_z,
                                                                             (float)sgn_x, HZc, z0, z0 + tz, acc2);
                        }
#pragma unroll
                        for (int r = 0; r < kRows; ++r) { as[r] = real(acc2[r].x); ac[r] = real(acc2[r].y); }
                    } else if (GEO == 3 || GEO == 4)
                        couple_grid_sym_chunks<CH, CL != 0>(reinterpret_cast<const float*>(SC + pbuf * scsz + sc_sector),
                                                   reinterpret_cast<const float*>(T), GZ, GX, zq, xq, chunk, (float)sgn_z, (float)sgn_x,
                                                   reinterpret_cast<float(&)[kRows]>(as), reinterpret_cast<float(&)[kRows]>(ac));
                    else if (MW)
                        couple_sym_upre(reinterpret_cast<const float*>(SC + pbuf * scsz + sc_sector), U4, tid,
                                        reinterpret_cast<float(&)[kRows]>(as), reinterpret_cast<float(&)[kRows]>(ac));
                    else if (GEO == 1 && sizeof(real) == 4)
                        couple_grid_sym_fixed<4, 4>(reinterpret_cast<const float*>(SC + pbuf * scsz + sc_sector),
                                                    reinterpret_cast<const float*>(T), zq, xq, (float)sgn_z, (float)sgn_x, 4,
                                                    reinterpret_cast<float(&)[kRows]>(as), reinterpret_cast<float(&)[kRows]>(ac));
                    else if (GEO == 2 && sizeof(real) == 4)
                        couple_grid_sym_fixed<0, 4, CL != 0>(reinterpret_cast<const float*>(SC + pbuf * scsz + sc_sector),
                                                    reinterpret_cast<const float*>(T), zq, xq, (float)sgn_z, (float)sgn_x, GZ >> 1,
                                                    reinterpret_cast<float(&)[kRows]>(as), reinterpret_cast<float(&)[kRows]>(ac));
                    else
                        couple_grid_sym<real>(SC + pbuf * scsz + sc_sector, T, GZ, GX, zq, xq, sgn_z, sgn_x, as, ac);
                    if (YPAR) {                                  // (E, O) -> rows i and 7-i (x 1/2 folded into kn)
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            const real es = as[r], os = as[4 + r], ec = ac[r], oc = ac[4 + r];
                            as[r] = es + os; ac[r] = ec + oc;
                            as[4 + r] = es - os; ac[4 + r] = ec - oc;       // temporarily row 7-r at index 4+r
                        }
#pragma unroll
                        for (int r = 0; r < 2; ++r) {                    // index 4+r holds row 7-r: reverse the upper half
                            real t_ = as[4 + r]; as[4 + r] = as[7 - r]; as[7 - r] = t_;
                            t_ = ac[4 + r]; ac[4 + r] = ac[7 - r]; ac[7 - r] = t_;
                        }
                    }
                    // back to the grid lines (x 1/4 folded into kn)
                    if constexpr (sizeof(real) == 4)
                        quad_butterfly2f<BMX, BMZ>(reinterpret_cast<float(&)[kRows]>(as), reinterpret_cast<float(&)[kRows]>(ac),
                                                   (float)sgn_x, (float)sgn_z, wmask);
                    else quad_butterfly2<real, BMX, BMZ>(as, ac, sgn_x, sgn_z, wmask);
                } else couple_grid<real>(SC + pbuf * scsz, T, GZ, GX, zi, xi, as, ac);
                real ks[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if constexpr (SPEC) ks[r] = fma_r(cv[r], as[r], fma_r(-sv[r], ac[r], c0[r]));    // K / (8 N) is folded into lambda
                    else ks[r] = c0[r] + kn * (cv[r] * as[r] - sv[r] * ac[r]);
                }
                store_row<real>(K + kslot(s) * Nl, tid, nt, ks);
                if (SPEC) {}                      // (partials / coefficients are fenced by the two barriers of spectral_contract)
                else if (SCB == 2 || CL) pbuf ^= 1;    // (the global operand of cluster mode is always double buffered)
                else env_sync();                  // operand buffer is about to be overwritten by the next stage
                ++n_rhs;
            }
            have_f0 = true;
            // d1 = y1 - y0 = dt * sum_j b_j k_j, recomputed here (bit-identical to the last stage's increment) so that
            // the stage loop body stays uniform and the compiler keeps ONE copy of the unrolled contraction
            {
                stage_increment<real>(6, K, Nl, tid, nt, d1);
#pragma unroll
                for (int r = 0; r < kRows; ++r) d1[r] *= dt;
            }

            // ---- embedded error estimate and step-size controller ----------------------
            double sq;
            {
                real sqr = real(0);               // 8 terms in the compute precision, then float64 across the CTA
                real e[kRows];
                {                                     // b_err[1] = 0; k7 lives in slot 1
                    constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                    constexpr double cf[] = {35.0 / 384 - 1951.0 / 21600, 500.0 / 1113 - 22642.0 / 50085, 125.0 / 192 - 451.0 / 720,
                                             -2187.0 / 6784 + 12231.0 / 42400, 11.0 / 84 - 649.0 / 6300, -1.0 / 60};
                    lincomb<real, 6>(K, Nl, tid, nt, sl, cf, e);
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if (i0 + r < p.N) {
                        const real yu0 = y0[r] + two_pi_r * real(WD[r * nt + tid]);
                        const real yu1 = yu0 + d1[r];
                        const real scale = atol + fmax(fabs(yu0), fabs(yu1)) * rtol;
                        const real q = (e[r] * dt) / scale;
                        sqr = fma_r(q, q, sqr);
                    }
                }
                sq = (double)sqr;
            }
            sq = warp_sum(sq);
            if (lane == 0) red[warp] = sq;
            env_sync();
            double tot = 0.0;
            for (int w = 0; w < nwarps; ++w) tot += red[w];
            if (CL) {                              // sum the per-CTA partials of the cluster (same order in every CTA)
                if (tid == 0) cls[(cl_par * NC_ + crank) * kClSlots] = tot;
                cluster_barrier();
                tot = 0.0;
                for (int r = 0; r < NC_; ++r) tot += __ldcg(cls + (cl_par * NC_ + r) * kClSlots);
                cl_par ^= 1;
            }
            const double err = sqrt(tot / (double)p.N);
            if (!(err == err)) { status |= STATUS_NAN; break; }
            const bool keep = err < 1.0;
            double factor;
            if (err == 0.0) factor = p.fmax;
            else factor = fmin(fmax(p.safety * pow(1.0 / err, 0.2), keep ? 1.0 : p.fmin), p.fmax);
            const double dt_next = dt_d * factor;

            double t_new0;
            if (keep) {
                ++n_acc;
                // ---- dense output (4th-order Dopri5 interpolant, increment form) -------
                if (save_idx < n_ts && ts[save_idx] <= tnext) {
                    real f0[kRows], pa[kRows], pb[kRows], pc[kRows];
                    {
                        real kk0[kRows], k6[kRows], dm[kRows];
                        load_row<real>(K, tid, nt, kk0);
                        load_row<real>(K + kslot(6) * Nl, tid, nt, k6);
                        {                             // c_mid[1] = 0; k7 lives in slot 1
                            constexpr int sl[] = {0, 2, 3, 4, 5, 1};
                            constexpr double cf[] = {0.5 * (6025192743.0 / 30085553152.0), 0.5 * (51252292925.0 / 65400821598.0),
                                                     0.5 * (-2691868925.0 / 45128329728.0), 0.5 * (187940372067.0 / 1594534317056.0),
                                                     0.5 * (-1776094331.0 / 19743644256.0), 0.5 * (11237099.0 / 235043384.0)};
                            lincomb<real, 6>(K, Nl, tid, nt, sl, cf, dm);
                        }
#pragma unroll
                        for (int r = 0; r < kRows; ++r) {
                            const real f0r = kk0[r] * dt, f1r = k6[r] * dt, dmr = dm[r] * dt, d = d1[r];
                            f0[r] = f0r;
                            pa[r] = real(2) * (f1r - f0r) - real(8) * d + real(16) * dmr;
                            pb[r] = real(5) * f0r - real(3) * f1r + real(14) * d - real(32) * dmr;
                            pc[r] = f1r - real(4) * f0r - real(5) * d + real(16) * dmr;
                        }
                    }
                    while (save_idx < n_ts && ts[save_idx] <= tnext) {
                        int nb = 0;
                        while (nb < kSampleBatch && save_idx + nb < n_ts && ts[save_idx + nb] <= tnext) {
                            const int idx = save_idx + nb;
                            if (idx >= rec_from && idx < n_rec) {
                                const double tsv = ts[idx];
                                const bool at_end = (tsv == tnext);
                                // fp32: the quotient of the (exact, float64) differences in single precision -- a float64
                                // division per sample and thread costs ~40 instructions for one ulp of tau
                                const real tau = (tnext == t) ? real(0)
                                               : (sizeof(real) == 4 ? real((float)(tsv - t) / (float)(tnext - t)) : real((tsv - t) / (tnext - t)));
                                real st = real(0), sr = real(0);
                                real rc[kRows];
                                if (!MW) loadv<kRows>(RC + k0, rc);
                                else if (p.weighted_rec) loadv<kRows>(reinterpret_cast<const real*>(p.rec) + base + i0, rc);
                                else {
#pragma unroll
                                    for (int r = 0; r < kRows; ++r) rc[r] = real(0);
                                }
#pragma unroll
                                for (int r = 0; r < kRows; ++r) {
                                    real inc = (((pa[r] * tau + pb[r]) * tau + pc[r]) * tau + f0[r]) * tau;
                                    if (at_end) inc = d1[r];
                                    const real c = cos_r(y0[r] + inc);
                                    if (i0 + r < p.N) { st += c; sr = fma_r(c, rc[r], sr); }
                                }
                                // 256 terms of magnitude <= 1 per warp: the compute precision is ample here
                                const real wt = warp_sum(st), wr = warp_sum(sr);
                                if (lane == 0) {
                                    part[(warp * kSampleBatch + nb) * 2] = (double)wt;
                                    part[(warp * kSampleBatch + nb) * 2 + 1] = (double)wr;
                                }
                            }
                            ++nb;
                        }
                        env_sync();
                        if (CL) {                  // CTA partials -> global scratch -> summed by the rank-0 CTA below
                            for (int q = tid; q < nb; q += nt) {
                                double a_t = 0.0, a_r = 0.0;
                                for (int w = 0; w < nwarps; ++w) {
                                    a_t += part[(w * kSampleBatch + q) * 2];
                                    a_r += part[(w * kSampleBatch + q) * 2 + 1];
                                }
                                cls[(cl_par * NC_ + crank) * kClSlots + 2 * q] = a_t;
                                cls[(cl_par * NC_ + crank) * kClSlots + 2 * q + 1] = a_r;
                            }
                            cluster_barrier();
                        }
                        for (int q = tid; q < nb && crank == 0; q += nt) {
                            const int idx = save_idx + q;
                            if (idx >= rec_from && idx < n_rec) {
                                double a_t = 0.0, a_r = 0.0;
                                if (CL) {
                                    for (int r = 0; r < NC_; ++r) {
                                        a_t += __ldcg(cls + (cl_par * NC_ + r) * kClSlots + 2 * q);
                                        a_r += __ldcg(cls + (cl_par * NC_ + r) * kClSlots + 2 * q + 1);
                                    }
                                } else
                                for (int w = 0; w < nwarps; ++w) {
                                    a_t += part[(w * kSampleBatch + q) * 2];
                                    a_r += part[(w * kSampleBatch + q) * 2 + 1];
                                }
                                a_t /= (double)p.N; a_r /= (double)p.N;
                                if (!p.weighted_rec) a_r = a_t;
                                if (p.mode == MODE_STEP) {
                                    p.lfp_true[(size_t)env * p.smax + out_base + idx] = a_t;
                                    p.lfp_rec[(size_t)env * p.smax + out_base + idx] = a_r;
                                } else {
                                    reinterpret_cast<real*>(p.ring)[(size_t)env * p.W + (idx - rec_from)] = real(a_r);
                                }
                            }
                        }
                        if (CL) cl_par ^= 1;
                        env_sync();
                        save_idx += nb;
                    }
                }
                // ---- accept: y0 <- y1, FSAL k1 <- k7 ------------------------------------
                {
                    real k6[kRows];
                    load_row<real>(K + kslot(6) * Nl, tid, nt, k6);
                    store_row<real>(K, tid, nt, k6);
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    real y1 = y0[r] + d1[r];
                    if (sizeof(real) == 4) {
                        // keep the fp32 phase wrapped: y = phase + 2*pi*wind (two-constant reduction)
                        const float n = floorf((float)y1 * 0.15915494309189535f);
                        if (n != 0.0f) {
                            float yw = fmaf(-n, 6.2831854820251465f, (float)y1);
                            yw = fmaf(-n, -1.7484555314695172e-07f, yw);
                            y1 = real(yw);
                            WD[r * nt + tid] += (int)n;
                        }
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
        k0_valid = p.fsal_on != 0; amp_k0 = amp;          // slot 0 = k7 of the last accepted sub-step = f(y0) with this pulse
    }

    // ---- write back ------------------------------------------------------------------------
    storev<kRows>(reinterpret_cast<real*>(p.phase) + base + i0, y0);
    if (sizeof(real) == 4) {
#pragma unroll
        for (int r = 0; r < kRows; ++r) p.wind[base + i0 + r] = WD[r * nt + tid];
    }
    if (tail)                                                              // (the samples were written before the
        obs_tail<real>(p, env, lane, seg_nrec[0] + seg_nrec[1], t_delta, t_pos);   //  last barrier of the solve loop)
    if (REUSE && p.fsal_on) {
        const bool keep_row = k0_valid && amp_k0 == real(0);     // (segment II and the transient run without a pulse)
        if (keep_row) {
            real k[kRows];
            load_row<real>(K, tid, nt, k);
            store_row<real>(reinterpret_cast<real*>(p.k_fsal) + base + (CL ? crank * Nl : 0), tid, nt, k);
        }
        if (tid == 0 && crank == 0) p.fsal_valid[env] = keep_row ? 1 : 0;
    }
    if (tid == 0 && crank == 0) {
        if (p.mode == MODE_TRANSIENT) p.head[env] = 0;
        atomicAdd(p.counters + 0, (unsigned long long)n_acc);
        atomicAdd(p.counters + 1, (unsigned long long)n_rej);
        atomicAdd(p.counters + 2, (unsigned long long)n_rhs);
        atomicAdd(p.counters + 3, (unsigned long long)n_reuse);
        if (status) atomicOr(p.status, status);
    }
    if (!MW) break;
    env_sync();                                           // the worker's shared memory is reused by its next environment
    }
}

inline size_t step_smem_bytes_mw(int Np) { return kMwUFloat4 * sizeof(float4) + kMwEnvs * step_smem_bytes_worker(Np); }
template <int RE, int RO>
inline size_t step_smem_bytes_spectral(int Np, int workers) { return (size_t)workers * step_smem_bytes_worker(Np, SpecLayout<RE, RO>::floats); }

inline size_t step_smem_bytes_cluster(int nthreads, size_t real_bytes) {
    const int nwarps = (nthreads + 31) / 32, Nl = nthreads * kRows;
    return (size_t)((kSlots + 1) * Nl) * real_bytes + (size_t)Nl * sizeof(int) +
           (size_t)(nwarps * kSampleBatch * 2 + nwarps + 32) * sizeof(double) + 36 * sizeof(int) +
           (size_t)(kClTileFloats + 16) * sizeof(float) + 16;
}

// low-rank cluster mode: K slots, the double-buffered operand of the CTA's own oscillators, mode coefficients, recording
// conductance, winding counts, reduction scratch, the two parity buffers of partial mode sums
inline size_t step_smem_bytes_cluster_lr(int nthreads, int rank4) {
    const int nwarps = (nthreads + 31) / 32, Nl = nthreads * kRows;
    return (size_t)((kSlots + 1) * Nl + kScBuffers * (2 * Nl + kScPad) + 2 * rank4) * sizeof(float) + (size_t)Nl * sizeof(int) +
           (size_t)(nwarps * kSampleBatch * 2 + nwarps + 32) * sizeof(double) + 36 * sizeof(int) + 64 +
           (size_t)2 * rank4 * sizeof(float2);          // the CTA's partial mode sums, read by the other ranks through DSMEM
}

inline size_t step_smem_bytes(int Np, int tab, int nthreads, size_t real_bytes) {
    const int nwarps = (nthreads + 31) / 32;
    return (size_t)((kSlots + 1 + 2 * kScBuffers) * Np + kScBuffers * kScPad + tab) * real_bytes + (size_t)Np * sizeof(int) +
           (size_t)(nwarps * kSampleBatch * 2 + nwarps + 32) * sizeof(double) + 36 * sizeof(int);
}

}  // namespace dbsgym
