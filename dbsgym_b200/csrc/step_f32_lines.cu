// step-kernel instantiations: float, parity-sector contraction with lines of 16 (GEO 3) and 32 (GEO 4) oscillators
#include "step_launch.h"
namespace dbsgym {
cudaError_t launch_f32_sym_lines(int geo, int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    if (geo == 3) return launch_by_threads<float, CPL_GRID_SYM, 3, 64, 512>(threads, smem, p, s);
    return launch_by_threads<float, CPL_GRID_SYM, 4, 128, 512>(threads, smem, p, s);
}
}  // namespace dbsgym
