// step-kernel instantiations: float, 8 x 8 x 8 grid, spectral contraction (generalised mean-field identity)
#include "step_launch.h"
#ifndef DBSGYM_SPECTRAL_ENVS
#define DBSGYM_SPECTRAL_ENVS 8       // environments (64-thread workers) per CTA
#endif
namespace dbsgym {
constexpr int kSpEnvs = DBSGYM_SPECTRAL_ENVS;
int spectral_envs_per_cta() { return kSpEnvs; }

template <int RE, int RO>
static cudaError_t launch_spec(int num_sms, const StepParams& p, cudaStream_t s) {
    auto kern = step_kernel<float, CPL_SPECTRAL, kSpEnvs * kMwThreads, 1, 0, kSpEnvs, RE, RO>;
    const size_t smem = step_smem_bytes_spectral<RE, RO>(p.Np, kSpEnvs);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int ctas = (p.n_launch + kSpEnvs - 1) / kSpEnvs;
    if (ctas > num_sms) ctas = num_sms;
    kern<<<ctas, kSpEnvs * kMwThreads, smem, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_f32_spectral(int r_odd, int num_sms, const StepParams& p, cudaStream_t s) {
    if (r_odd <= 4) return launch_spec<9, 4>(num_sms, p, s);
    return launch_spec<9, 9>(num_sms, p, s);
}
}  // namespace dbsgym
