// step-kernel instantiations: float, CPL_LOWRANK (DENSE operator in its truncated eigenbasis), one environment per CTA
#include "step_launch.h"
namespace dbsgym {
cudaError_t launch_f32_lowrank(int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    return launch_by_threads<float, CPL_LOWRANK, 0>(threads, smem, p, s);
}
}  // namespace dbsgym
