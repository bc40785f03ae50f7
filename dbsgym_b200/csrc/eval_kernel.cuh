// Evaluation metric on the device (sm_100a): the beta-band power the paper reports,
// reference aDBS_RL/evaluate_HF_DBS.py:122-135 (calc_psd_for_simple_eval), for every environment of a batch from
// the TRUE-LFP trace (theta_mean, env.py:441) the step kernel's tail records while evaluation runs:
//     sig_f = filtfilt(butter(2, [12, 30] Hz), sig)                 (utils.py:794-816, order 2)
//     ft    = |rfft(sig_f) / n|^2 * 2
//     ft    = filtfilt([1]*12, 5, ft)                               (spectrum smoothing)
//     bbpow = sum of ft over the bins with 12.5 Hz < f < 21 Hz
// The smoothing and the band sum are linear in ft, so bbpow = sum_k w_k ft_k with weights w the host obtains by
// running scipy's own filtfilt on unit vectors (dbsgym_b200/evaluation.py); only the ~130 bins with w_k != 0 are
// evaluated.  The band-pass filtfilt is restated exactly as scipy runs it: odd extension by padlen samples, direct
// form II transposed lfilter started from zi * first sample, forward, then backward.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dbsgym {

struct EvalParams {
    const double* trace; int cap; int n; int B;
    double* scratch;              // [B][n + 2 * pad] extended / filtered signal
    int pad;
    double b[5], a[5], zi[4];
    int k_lo, n_k;
    const double* weights;        // [n_k]
    double* out;                  // [B]
};

// one thread per environment: the recurrence is sequential in time
__global__ void __launch_bounds__(32) eval_filtfilt_kernel(const EvalParams p) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.B) return;
    const double* x = p.trace + (size_t)env * p.cap;
    const int n = p.n, pad = p.pad, m = n + 2 * pad;
    double* y = p.scratch + (size_t)env * m;
    auto ext = [&](int i) -> double {      // scipy.signal._arraytools.odd_ext
        if (i < pad) return 2.0 * x[0] - x[pad - i];
        if (i < pad + n) return x[i - pad];
        return 2.0 * x[n - 1] - x[n - 2 - (i - pad - n)];
    };
    const double b0 = p.b[0], b1 = p.b[1], b2 = p.b[2], b3 = p.b[3], b4 = p.b[4];
    const double a1 = p.a[1], a2 = p.a[2], a3 = p.a[3], a4 = p.a[4];
    double z0, z1, z2, z3;
    {
        const double x0 = ext(0);
        z0 = p.zi[0] * x0; z1 = p.zi[1] * x0; z2 = p.zi[2] * x0; z3 = p.zi[3] * x0;
    }
    auto step = [&](double xi) -> double {  // one sample of lfilter, direct form II transposed (scipy's recurrence)
        const double yi = b0 * xi + z0;
        z0 = b1 * xi + z1 - a1 * yi;
        z1 = b2 * xi + z2 - a2 * yi;
        z2 = b3 * xi + z3 - a3 * yi;
        z3 = b4 * xi - a4 * yi;
        return yi;
    };
    // forward pass; the interior is unrolled so that the (independent) loads of 8 samples are in flight together --
    // the recurrence itself is sequential, one thread per environment
    for (int i = 0; i < pad; ++i) y[i] = step(ext(i));
#pragma unroll 8
    for (int j = 0; j < n; ++j) y[pad + j] = step(x[j]);
    for (int i = pad + n; i < m; ++i) y[i] = step(ext(i));
    {
        const double x0 = y[m - 1];
        z0 = p.zi[0] * x0; z1 = p.zi[1] * x0; z2 = p.zi[2] * x0; z3 = p.zi[3] * x0;
    }
#pragma unroll 8
    for (int i = m - 1; i >= 0; --i) y[i] = step(y[i]);   // backward pass, in place
}

// one CTA per environment, one bin per thread: X_k = sum_t sig_f[t] e^{-j 2 pi k t / n} with a rotation recurrence
// re-seeded every 64 samples from the exactly reduced phase (k t mod n); every thread reads the same sample (broadcast)
constexpr int kEvalThreads = 128;
__global__ void __launch_bounds__(kEvalThreads) eval_band_power_kernel(const EvalParams p) {
    __shared__ double part[kEvalThreads / 32];
    const int env = blockIdx.x, tid = threadIdx.x;
    const int n = p.n;
    const double* sig = p.scratch + (size_t)env * (n + 2 * p.pad) + p.pad;
    double acc = 0.0;
    for (int kb = tid; kb < p.n_k; kb += kEvalThreads) {
        const long long k = p.k_lo + kb;
        double cd, sd;
        sincospi(2.0 * (double)k / (double)n, &sd, &cd);
        double re = 0.0, im = 0.0;
        for (int t0 = 0; t0 < n; t0 += 64) {
            double c, s;
            sincospi(2.0 * (double)((k * t0) % n) / (double)n, &s, &c);
            const int t1 = t0 + 64 < n ? t0 + 64 : n;
            for (int t = t0; t < t1; ++t) {
                const double x = sig[t];
                re = fma(x, c, re);
                im = fma(x, s, im);
                const double c2 = c * cd - s * sd;
                s = fma(c, sd, s * cd);
                c = c2;
            }
        }
        const double inv = 1.0 / (double)n;
        re *= inv; im *= inv;
        acc = fma(p.weights[kb], (re * re + im * im) * 2.0, acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) part[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int w = 0; w < kEvalThreads / 32; ++w) v += part[w];
        p.out[env] = v;
    }
}

}  // namespace dbsgym
