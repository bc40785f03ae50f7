// Launch entry points of the step-kernel instantiations.  Every group of instantiations lives in its own translation
// unit (step_*.cu) so that the library builds in parallel and an edit of one kernel family recompiles one file.
#pragma once
#include "step_kernel.cuh"

namespace dbsgym {

// one environment per CTA, `threads` = N / 8 (<= 1024); geo: 0 run-time extents, 1 = 8 x 8 x 8, 2 = gx 8 fixed,
// 3 / 4 = lines of 16 / 32 (GRID_SYM float only)
cudaError_t launch_f32_grid(int threads, size_t smem, const StepParams& p, cudaStream_t s);
cudaError_t launch_f32_sym(int geo, int threads, size_t smem, const StepParams& p, cudaStream_t s);
cudaError_t launch_f32_dense(int threads, size_t smem, const StepParams& p, cudaStream_t s);
// DENSE operator in low-rank form (dbsgym_set_coupling_lowrank): smem includes 2 * p.lr_rank floats of mode coefficients
cudaError_t launch_f32_lowrank(int threads, size_t smem, const StepParams& p, cudaStream_t s);
cudaError_t launch_f32_lowrank_cluster(int threads, int cluster, int rank4, const StepParams& p, cudaStream_t s);
cudaError_t launch_f64_grid(int threads, size_t smem, const StepParams& p, cudaStream_t s);
cudaError_t launch_f64_sym(int threads, size_t smem, const StepParams& p, cudaStream_t s);
cudaError_t launch_f64_dense(int threads, size_t smem, const StepParams& p, cudaStream_t s);
// multi-worker kernel (8 x 8 x 8, exact sector contraction): kMwEnvs environments per CTA
cudaError_t launch_f32_mw(int ctas, const StepParams& p, cudaStream_t s);
// spectral kernel (8 x 8 x 8): r_odd <= 4 selects the (9, 4) instantiation, otherwise (9, 9); *workers = environments per CTA
cudaError_t launch_f32_spectral(int r_odd, int num_sms, const StepParams& p, cudaStream_t s);
int spectral_envs_per_cta();
// spectral kernel, one warp per environment (warp_kernel.cuh): rank_set = index returned by warp_kernel_rank_set for the
// eight sector ranks (-1: no compiled rank list covers them); compiled8 receives the ranks the tables must be padded to
int warp_kernel_rank_set(const int* ranks8, int* compiled8);
cudaError_t launch_f32_warp(int rank_set, int num_sms, const StepParams& p, cudaStream_t s);
int warp_envs_per_cta();
// the same for the 8 x 8 x 4 half grid, one octant point per lane (warp1_kernel.cuh)
int warp1_kernel_rank_set(const int* ranks8, int* compiled8);
cudaError_t launch_f32_warp1(int rank_set, int num_sms, const StepParams& p, cudaStream_t s);
int warp1_envs_per_cta();
// regular grids of 1024 / 2048 / 4096 oscillators in octant order, eigenvectors in registers, one CTA per environment (oct_kernel.cuh)
int oct_kernel_rank_set(int n_osc, const int* ranks8, int* compiled8);
cudaError_t launch_f32_oct(int rank_set, int num_sms, const StepParams& p, cudaStream_t s);
// cluster mode: one environment = `cluster` CTAs
cudaError_t launch_f32_cluster(int geo, int threads, int cluster, const StepParams& p, cudaStream_t s);

// ---- helpers for the translation units -----------------------------------------------------------------------
template <typename real, int CPL, int MAXT, int GEO>
cudaError_t launch_one(int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    auto kern = step_kernel<real, CPL, MAXT, GEO>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<p.n_launch, threads, smem, s>>>(p);
    return cudaGetLastError();
}

template <typename real, int CPL, int GEO, int MINT = 64, int MAXTT = 1024>
cudaError_t launch_by_threads(int threads, size_t smem, const StepParams& p, cudaStream_t s) {
    if constexpr (MINT <= 64) if (threads <= 64) return launch_one<real, CPL, 64, GEO>(threads, smem, p, s);
    if constexpr (MINT <= 128 && MAXTT >= 128) if (threads <= 128) return launch_one<real, CPL, 128, GEO>(threads, smem, p, s);
    if constexpr (MAXTT >= 256) if (threads <= 256) return launch_one<real, CPL, 256, GEO>(threads, smem, p, s);
    if constexpr (MAXTT >= 512) if (threads <= 512) return launch_one<real, CPL, 512, GEO>(threads, smem, p, s);
    if constexpr (MAXTT >= 1024) return launch_one<real, CPL, 1024, GEO>(threads, smem, p, s);
    return cudaErrorInvalidConfiguration;
}

}  // namespace dbsgym
