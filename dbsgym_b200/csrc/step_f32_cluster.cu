// step-kernel instantiations: float, cluster mode (one environment = a thread-block cluster, N > 4096)
#include "step_launch.h"
namespace dbsgym {
template <int MAXT, int GEO>
static cudaError_t launch_cl(int threads, int cluster, const StepParams& p, cudaStream_t s) {
    auto kern = step_kernel<float, CPL_GRID_SYM, MAXT, GEO, 1>;
    const size_t smem = step_smem_bytes_cluster(threads, sizeof(float));
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

cudaError_t launch_f32_cluster(int geo, int t, int cluster, const StepParams& p, cudaStream_t s) {
    if (geo == 3) return t <= 256 ? launch_cl<256, 3>(t, cluster, p, s) : launch_cl<512, 3>(t, cluster, p, s);
    if (geo == 4) return t <= 256 ? launch_cl<256, 4>(t, cluster, p, s) : launch_cl<512, 4>(t, cluster, p, s);
    if (t <= 64) return launch_cl<64, 2>(t, cluster, p, s);
    if (t <= 128) return launch_cl<128, 2>(t, cluster, p, s);
    if (t <= 256) return launch_cl<256, 2>(t, cluster, p, s);
    return launch_cl<512, 2>(t, cluster, p, s);
}
}  // namespace dbsgym
