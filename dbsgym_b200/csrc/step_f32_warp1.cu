// step-kernel instantiations: float, 8 x 8 x 4 half grid (N = 256), spectral contraction, one warp per environment with one
// octant point per lane (warp1_kernel.cuh)
#ifndef DBSGYM_WARP1_ENVS
#define DBSGYM_WARP1_ENVS 16         // environments (warps) per CTA; one persistent CTA per SM, 128 registers per thread
#endif
#include "warp1_kernel.cuh"
#include "step_launch.h"
namespace dbsgym {
constexpr int kWarp1Envs = DBSGYM_WARP1_ENVS;
int warp1_envs_per_cta() { return kWarp1Envs; }

// compiled rank lists (modes per parity sector s = 4 [odd y] + 2 [odd z] + [odd x]); the shipped cos(distance) kernel on
// the 8 x 8 x 4 grid has 26 modes above 1e-9 |lambda_max| and 32 above 1e-10
using HalfRanks26 = RankSet<5, 4, 4, 2, 4, 4, 2, 1>;
using HalfRanks32 = RankSet<9, 4, 4, 3, 4, 4, 3, 1>;
static const int kHalfRankSets[2][8] = {{5, 4, 4, 2, 4, 4, 2, 1}, {9, 4, 4, 3, 4, 4, 3, 1}};

template <class RK>
static cudaError_t launch_w1(int num_sms, const StepParams& p, cudaStream_t s) {
    using L = Warp1Layout<RK>;
    auto kern = warp1_step_kernel<RK>;
    int warps = kWarp1Envs;                                  // as many as the 227 KB of an SM hold
    while (warps > 1 && (size_t)warps * L::bytes_aligned > (size_t)227 * 1024) --warps;
    const size_t smem = (size_t)warps * L::bytes_aligned;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas = (p.n_launch + warps - 1) / warps;
    if (ctas > num_sms) ctas = num_sms;
    kern<<<ctas, warps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

// index of the smallest compiled rank list that covers the requested ranks (-1: none does)
int warp1_kernel_rank_set(const int* ranks8, int* compiled8) {
    for (int k = 0; k < 2; ++k) {
        bool ok = true;
        for (int i = 0; i < 8; ++i) ok = ok && ranks8[i] <= kHalfRankSets[k][i];
        if (ok) {
            if (compiled8) for (int i = 0; i < 8; ++i) compiled8[i] = kHalfRankSets[k][i];
            return k;
        }
    }
    return -1;
}

cudaError_t launch_f32_warp1(int rank_set, int num_sms, const StepParams& p, cudaStream_t s) {
    if (rank_set == 0) return launch_w1<HalfRanks26>(num_sms, p, s);
    if (rank_set == 1) return launch_w1<HalfRanks32>(num_sms, p, s);
    return cudaErrorInvalidConfiguration;
}
}  // namespace dbsgym
