// step-kernel instantiations: float, CPL_LOWRANK (any coupling operator in its truncated eigenbasis), one CTA per
// environment, or one thread-block cluster per environment (N > 4096: every CTA owns 4096 consecutive oscillators)
#include "step_launch.h"
namespace dbsgym {
cudaError_t launch_f32_lowrank(int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    return launch_by_threads<float, CPL_LOWRANK, 0>(threads, smem, p, s);
}

template <int MAXT>
static cudaError_t launch_lr_cl(int threads, int cluster, int rank4, const StepParams& p, cudaStream_t s) {
    auto kern = step_kernel<float, CPL_LOWRANK, MAXT, 0, 1>;
    const size_t smem = step_smem_bytes_cluster_lr(threads, rank4);
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess && cluster > 8) e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p.n_launch * cluster));
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
}

cudaError_t launch_f32_lowrank_cluster(int threads, int cluster, int rank4, const StepParams& p, cudaStream_t s) {
    if (threads <= 64) return launch_lr_cl<64>(threads, cluster, rank4, p, s);
    if (threads <= 128) return launch_lr_cl<128>(threads, cluster, rank4, p, s);
    if (threads <= 256) return launch_lr_cl<256>(threads, cluster, rank4, p, s);
    return launch_lr_cl<512>(threads, cluster, rank4, p, s);
}
}  // namespace dbsgym
