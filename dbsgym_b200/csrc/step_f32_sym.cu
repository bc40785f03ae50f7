// step-kernel instantiations: float, parity-sector contraction with run-time extents (GEO 0), the 8 x 8 x 8 grid (GEO 1)
// and 8 x 8 x gz grids (GEO 2); one environment per CTA
#include "step_launch.h"
namespace dbsgym {
cudaError_t launch_f32_sym_lines(int geo, int threads, size_t smem, const StepParams& p, cudaStream_t s);   // step_f32_lines.cu
cudaError_t launch_f32_sym(int geo, int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    if (geo == 1) return launch_one<float, CPL_GRID_SYM, 64, 1>(threads, smem, p, s);
    if (geo == 2) return launch_by_threads<float, CPL_GRID_SYM, 2, 64, 512>(threads, smem, p, s);
    if (geo == 3 || geo == 4) return launch_f32_sym_lines(geo, threads, smem, p, s);
    return launch_by_threads<float, CPL_GRID_SYM, 0>(threads, smem, p, s);
}
}  // namespace dbsgym
