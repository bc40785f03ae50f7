// step-kernel instantiations: double, CPL_DENSE, 0 (one environment per CTA, 64 .. 1024 threads)
#include "step_launch.h"
namespace dbsgym {
cudaError_t launch_f64_dense(int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    return launch_by_threads<double, CPL_DENSE, 0>(threads, smem, p, s);
}
}  // namespace dbsgym
