// step-kernel instantiations: float, regular grids of 1024 / 2048 / 4096 oscillators in octant order, spectral contraction with
// the eigenvectors in registers, one CTA per environment (oct_kernel.cuh)
#include "oct_kernel.cuh"
#include "step_launch.h"
namespace dbsgym {

// compiled (oscillator count, rank list) pairs: modes per parity sector s = 4 [odd y] + 2 [odd z] + [odd x] of the shipped
// cos(distance) kernel above 1e-9 |lambda_max| on the grids of BASELINE configs[4] (the first 4 / 8 / 16 z-planes of the
// 16 x 16 x 16 grid) and on the elongated 8 x 8 x gz grids of scripts/sweep_n.py; odd totals padded by one mode
struct OctSet { int n_osc; int ranks[8]; };
static const OctSet kOctSets[] = {
    {1024, {9, 6, 5, 3, 6, 4, 3, 2}},      // 16 x 16 x 4
    {1024, {8, 5, 5, 4, 5, 4, 4, 3}},      // 8 x 8 x 16
    {2048, {9, 8, 6, 4, 8, 4, 4, 5}},      // 16 x 16 x 8
    {2048, {9, 7, 8, 5, 7, 5, 5, 4}},      // 8 x 8 x 32
    {4096, {9, 9, 9, 4, 9, 4, 4, 4}},      // 16 x 16 x 16
    {4096, {11, 8, 11, 7, 8, 5, 7, 5}},    // 8 x 8 x 64
    {8192, {14, 10, 10, 8, 10, 9, 8, 5}},  // 32 x 32 x 8: a cluster of 2 CTAs per environment (1.49 -> 1.03 ms per 256-env step).
                                           // (32 x 32 x 16 as a cluster of 4 with its 86 modes was measured SLOWER than the block kernel --
                                           //  1.76 vs 1.61 ms, 320 bytes of spills per thread -- and is not compiled)
};
constexpr int kNumOctSets = sizeof(kOctSets) / sizeof(kOctSets[0]);

template <class RK, int NW, int NC = 1>
static cudaError_t launch_o(int num_sms, const StepParams& p, cudaStream_t s) {
    using L = OctLayout<RK, NW, NC>;
    auto kern = oct_step_kernel<RK, NW, NC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes);
    if (e != cudaSuccess) return e;
    if constexpr (NC == 1) {
        int ctas = num_sms * (16 / NW);
        if (ctas > p.n_launch) ctas = p.n_launch;
        kern<<<ctas, 32 * NW, L::bytes, s>>>(p);
        return cudaGetLastError();
    } else {                                               // one cluster of NC CTAs (one CTA per SM) per environment in flight
        int clusters = num_sms / NC;
        if (clusters > p.n_launch) clusters = p.n_launch;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(clusters * NC));
        cfg.blockDim = dim3(32 * NW);
        cfg.dynamicSmemBytes = L::bytes;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, p);
    }
}

// index of the compiled rank list for this oscillator count that covers the requested ranks with the fewest modes (-1: none)
int oct_kernel_rank_set(int n_osc, const int* ranks8, int* compiled8) {
    int best = -1, best_modes = 1 << 30;
    for (int k = 0; k < kNumOctSets; ++k) {
        if (kOctSets[k].n_osc != n_osc) continue;
        bool ok = true;
        int modes = 0;
        for (int i = 0; i < 8; ++i) { ok = ok && ranks8[i] <= kOctSets[k].ranks[i]; modes += kOctSets[k].ranks[i]; }
        if (ok && modes < best_modes) { best = k; best_modes = modes; }
    }
    if (best >= 0 && compiled8) for (int i = 0; i < 8; ++i) compiled8[i] = kOctSets[best].ranks[i];
    return best;
}

cudaError_t launch_f32_oct(int rank_set, int num_sms, const StepParams& p, cudaStream_t s) {
    switch (rank_set) {
        case 0: return launch_o<RankSet<9, 6, 5, 3, 6, 4, 3, 2>, 4>(num_sms, p, s);
        case 1: return launch_o<RankSet<8, 5, 5, 4, 5, 4, 4, 3>, 4>(num_sms, p, s);
        case 2: return launch_o<RankSet<9, 8, 6, 4, 8, 4, 4, 5>, 8>(num_sms, p, s);
        case 3: return launch_o<RankSet<9, 7, 8, 5, 7, 5, 5, 4>, 8>(num_sms, p, s);
        case 4: return launch_o<RankSet<9, 9, 9, 4, 9, 4, 4, 4>, 16>(num_sms, p, s);
        case 5: return launch_o<RankSet<11, 8, 11, 7, 8, 5, 7, 5>, 16>(num_sms, p, s);
        case 6: return launch_o<RankSet<14, 10, 10, 8, 10, 9, 8, 5>, 16, 2>(num_sms, p, s);
    }
    return cudaErrorInvalidConfiguration;
}
}  // namespace dbsgym
