// step-kernel instantiation: float, 8 x 8 x 8 grid, multi-worker hosting of the exact sector contraction
#include "step_launch.h"
namespace dbsgym {
cudaError_t launch_f32_mw(int ctas, const StepParams& p, cudaStream_t s) {
    auto kern = step_kernel<float, CPL_GRID_SYM, kMwEnvs * kMwThreads, 1, 0, kMwEnvs>;
    const size_t smem = step_smem_bytes_mw(p.Np);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<ctas, kMwEnvs * kMwThreads, smem, s>>>(p);
    return cudaGetLastError();
}
}  // namespace dbsgym
