// step-kernel instantiations: float, 8 x 8 x 8 grid, spectral contraction, one warp per environment (warp_kernel.cuh)
#ifndef DBSGYM_WARP_ENVS
#define DBSGYM_WARP_ENVS 8           // environments (warps) per CTA; one persistent CTA per SM
#endif
#ifndef DBSGYM_WARP_MAXNREG
#define DBSGYM_WARP_MAXNREG (65536 / (DBSGYM_WARP_ENVS * 32) / 8 * 8 > 255 ? 255 : 65536 / (DBSGYM_WARP_ENVS * 32) / 8 * 8)
#endif
#include "warp_kernel.cuh"
#include "step_launch.h"
namespace dbsgym {
constexpr int kWarpEnvs = DBSGYM_WARP_ENVS;
int warp_envs_per_cta() { return kWarpEnvs; }

// compiled rank lists (modes per parity sector s = 4 [odd y] + 2 [odd z] + [odd x]); the shipped cos(distance) kernel on
// the 8 x 8 x 8 grid has 32 modes above 1e-9 |lambda_max| and 34 above 1e-10
using Ranks32 = RankSet<7, 4, 4, 4, 4, 4, 4, 1>;
using Ranks34 = RankSet<9, 4, 4, 4, 4, 4, 4, 1>;
static const int kRankSets[2][8] = {{7, 4, 4, 4, 4, 4, 4, 1}, {9, 4, 4, 4, 4, 4, 4, 1}};

template <class RK>
static cudaError_t launch_w(int num_sms, const StepParams& p, cudaStream_t s) {
    using L = WarpLayout<RK>;
    auto kern = warp_step_kernel<RK>;
    int warps = kWarpEnvs;                                   // as many as the 227 KB of an SM hold (larger rank lists: fewer)
    while (warps > 1 && L::cta_bytes(warps) > (size_t)227 * 1024) --warps;
    const size_t smem = L::cta_bytes(warps);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas = (p.n_launch + warps - 1) / warps;
    if (ctas > num_sms) ctas = num_sms;
    kern<<<ctas, warps * 32, smem, s>>>(p);
    return cudaGetLastError();
}

// index of the smallest compiled rank list that covers the requested ranks (-1: none does)
int warp_kernel_rank_set(const int* ranks8, int* compiled8) {
    for (int k = 0; k < 2; ++k) {
        bool ok = true;
        for (int i = 0; i < 8; ++i) ok = ok && ranks8[i] <= kRankSets[k][i];
        if (ok) {
            if (compiled8) for (int i = 0; i < 8; ++i) compiled8[i] = kRankSets[k][i];
            return k;
        }
    }
    return -1;
}

cudaError_t launch_f32_warp(int rank_set, int num_sms, const StepParams& p, cudaStream_t s) {
    if (rank_set == 0) return launch_w<Ranks32>(num_sms, p, s);
    if (rank_set == 1) return launch_w<Ranks34>(num_sms, p, s);
    return cudaErrorInvalidConfiguration;
}
}  // namespace dbsgym
