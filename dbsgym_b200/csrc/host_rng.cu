// Host-side replay of numpy's LEGACY global random stream (np.random.* of RandomState / MT19937) for batched resets.
//
// The reference draws everything a reset needs from the process-global np.random (environment/env.py:291, :483-598;
// SURVEY.md Appendix C), one environment after the other.  A batch of thousands of environments that wants the same
// numbers as a sequential DummyVecEnv of reference environments has to consume that one stream in the same order; doing
// it through thousands of small numpy calls costs ~75 us per environment in interpreter overhead alone.  This file
// restates the three generators involved so that the whole batch is drawn in one call, bit for bit:
//   * MT19937 as numpy seeds and steps it (state = np.random.get_state(): key[624], pos),
//   * legacy_gauss: Marsaglia's polar method on two 53-bit doubles per try, second deviate cached (has_gauss, gauss),
//     returned FIRST the f * x2 then (cached) f * x1 -- np.random.randn / standard_normal / normal(loc, scale),
//   * RandomState.randint(0, n) as used by np.random.choice: masked rejection on 32-bit words (smallest 2^k - 1 >= n - 1).
// The sequence of draws per environment is the reference's reset() (cited at each step below).  numpy's legacy stream is
// frozen by its compatibility policy (NEP 19); tests/test_host_rng.py checks bit equality against numpy itself.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/dbsgym.h"

namespace {

struct Mt {
    uint32_t* key;
    int pos;
    int has_gauss;
    double gauss;
};

inline void mt_gen(Mt& s) {
    constexpr int N = 624, M = 397;
    constexpr uint32_t MATRIX_A = 0x9908b0dfU, UPPER = 0x80000000U, LOWER = 0x7fffffffU;
    uint32_t y;
    int i;
    for (i = 0; i < N - M; i++) {
        y = (s.key[i] & UPPER) | (s.key[i + 1] & LOWER);
        s.key[i] = s.key[i + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    for (; i < N - 1; i++) {
        y = (s.key[i] & UPPER) | (s.key[i + 1] & LOWER);
        s.key[i] = s.key[i + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    }
    y = (s.key[N - 1] & UPPER) | (s.key[0] & LOWER);
    s.key[N - 1] = s.key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
    s.pos = 0;
}

inline uint32_t mt_next(Mt& s) {
    if (s.pos == 624) mt_gen(s);
    uint32_t y = s.key[s.pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680U;
    y ^= (y << 15) & 0xefc60000U;
    y ^= (y >> 18);
    return y;
}

inline double mt_double(Mt& s) {
    const int32_t a = (int32_t)(mt_next(s) >> 5), b = (int32_t)(mt_next(s) >> 6);
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

inline double legacy_gauss(Mt& s) {
    if (s.has_gauss) {
        const double t = s.gauss;
        s.has_gauss = 0;
        s.gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * mt_double(s) - 1.0;
        x2 = 2.0 * mt_double(s) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = std::sqrt(-2.0 * std::log(r2) / r2);
    s.gauss = f * x1;
    s.has_gauss = 1;
    return f * x2;
}

// RandomState.randint(0, n) for 1 <= n <= 2^32 (np.random.choice(n) / choice(list of n)): no draw for n == 1
inline uint32_t bounded(Mt& s, uint32_t n) {
    const uint32_t rng = n - 1;
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (mt_next(s) & mask)) > rng) {}
    return v;
}

// ---- legacy_gauss in two phases --------------------------------------------------------------------------------
// Evaluated one deviate at a time, the polar method is a chain of data-dependent branches (the accept test) interleaved
// with a logarithm, a division and a square root: ~39 ns per deviate.  GaussPipe consumes the stream in the same order but
// splits the work: phase 1 turns MT19937 words into accepted tries (x1, x2, r2) and remembers where the two deviates of
// each try go -- the second one is what numpy would keep in its cache for the NEXT call, wherever that call comes from --
// and phase 2 evaluates f = sqrt(-2 log(r2) / r2) for a whole block of tries in one tight loop (the libm calls pipeline):
// ~15-22 ns per deviate.  Same arithmetic per value as legacy_gauss, hence the same bits.
struct GaussPipe {
    static constexpr int kBlock = 8192;
    Mt& s;
    const double loc, scale;
    double r2_thr = 1.0;               // tries with r2 above this cannot produce loc + scale * g <= 0
    double x1[kBlock], x2[kBlock], r2[kBlock];
    double* d2[kBlock]; double* d1[kBlock];
    uint8_t a2[kBlock], a1[kBlock];
    int n = 0;
    bool half = false;                 // the last queued try still has its second deviate (f * x1) unassigned

    GaussPipe(Mt& s_, double loc_, double scale_) : s(s_), loc(loc_), scale(scale_) {
        if (scale > 0.0 && loc > scale) {                          // f(r2) = loc / scale by bisection, with a safety margin
            const double R = loc / scale * 0.999;
            double lo = 1e-300, hi = 1.0;
            for (int it = 0; it < 200; ++it) {
                const double mid = 0.5 * (lo + hi);
                if (fval(mid) >= R) lo = mid; else hi = mid;
            }
            r2_thr = hi;
        }
    }

    static inline double fval(double r) { return std::sqrt(-2.0 * std::log(r) / r); }

    // can loc + scale * (f x) be <= 0?  Only with x < 0 and f >= loc / scale, i.e. r2 below a threshold (f decreases
    // with r2): everything else is answered without evaluating the logarithm
    inline int maybe_bad(double r, double x) const {
        if (x >= 0.0 || r > r2_thr) return 0;
        return loc + scale * (fval(r) * x) <= 0.0;
    }

    // Next deviate of the stream into *dst (affine: loc + scale * g, otherwise g).  Returns 1 when an affine value came out
    // non-positive (decided at once, in the rare case it can happen at all), else 0.
    inline int next(double* dst, bool affine) {
        if (s.has_gauss) {                                          // numpy's own cached deviate from before this call
            const double g = s.gauss;
            s.has_gauss = 0; s.gauss = 0.0;
            *dst = affine ? loc + scale * g : g;
            return affine && *dst <= 0.0;
        }
        if (half) {                                                 // second deviate of the last try: f * x1
            const int k = n - 1;
            d1[k] = dst; a1[k] = affine;
            half = false;
            const int bad = affine ? maybe_bad(r2[k], x1[k]) : 0;
            if (n == kBlock) flush();
            return bad;
        }
        double u, v, r;
        do {
            u = 2.0 * mt_double(s) - 1.0;
            v = 2.0 * mt_double(s) - 1.0;
            r = u * u + v * v;
        } while (r >= 1.0 || r == 0.0);
        x1[n] = u; x2[n] = v; r2[n] = r; d2[n] = dst; d1[n] = nullptr; a2[n] = affine; a1[n] = 0;
        ++n;
        half = true;
        return affine ? maybe_bad(r, v) : 0;
    }

    void flush() {                                                  // (never called with an open try)
        for (int i = 0; i < n; ++i) r2[i] = fval(r2[i]);            // phase 2: the tight loop
        for (int i = 0; i < n; ++i) {
            const double f = r2[i];
            *d2[i] = a2[i] ? loc + scale * (f * x2[i]) : f * x2[i];
            if (d1[i]) *d1[i] = a1[i] ? loc + scale * (f * x1[i]) : f * x1[i];
        }
        n = 0;
    }

    // end of the call: whatever is queued is evaluated; an open second deviate becomes numpy's cached one
    void finish() {
        double cached = 0.0;
        if (half) { d1[n - 1] = &cached; a1[n - 1] = 0; }
        flush();
        if (half) { s.gauss = cached; s.has_gauss = 1; half = false; }
    }
};

}  // namespace

extern "C" {

int dbsgym_np_gauss(DbsGymNpState* st, int64_t n, double loc, double scale, double* out) {
    if (!st || (n > 0 && !out) || n < 0 || st->pos < 0 || st->pos > 624) return DBSGYM_EINVAL;
    Mt s{st->key, st->pos, st->has_gauss, st->gauss};
    for (int64_t i = 0; i < n; ++i) out[i] = loc + scale * legacy_gauss(s);
    st->pos = s.pos; st->has_gauss = s.has_gauss; st->gauss = s.gauss;
    return DBSGYM_OK;
}

int dbsgym_np_choice(DbsGymNpState* st, int64_t n, uint32_t pop_size, int32_t* out) {
    if (!st || (n > 0 && !out) || n < 0 || pop_size == 0 || st->pos < 0 || st->pos > 624) return DBSGYM_EINVAL;
    Mt s{st->key, st->pos, st->has_gauss, st->gauss};
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)bounded(s, pop_size);
    st->pos = s.pos;
    return DBSGYM_OK;
}

int dbsgym_np_reset_draws(DbsGymNpState* st, const DbsGymResetPlan* plan, const uint8_t* flags, const int32_t* freq,
                          int32_t* elec_coords, int32_t* next_inc, int32_t* spatial_pick, const int32_t* n_fix,
                          double* fix_noise, double* walk_noise, double* init_state, int32_t* refix_env,
                          double* refix_noise, int32_t* n_refix) {
    if (!st || !plan || !flags || !freq || !elec_coords || !next_inc || !spatial_pick || !n_fix || !init_state || !n_refix)
        return DBSGYM_EINVAL;
    if (plan->struct_bytes != sizeof(DbsGymResetPlan) || st->pos < 0 || st->pos > 624) return DBSGYM_EINVAL;
    DbsGymNpState work = *st;                                       // the caller's state changes only on success
    Mt s{work.key, work.pos, work.has_gauss, work.gauss};
    const int B = plan->n_envs, N = plan->n_osc, M = plan->walk_len;
    GaussPipe* gpp = new GaussPipe(s, plan->init_mean, plan->init_sd);      // (~400 KB of block arrays: heap, not stack)
    struct Guard { GaussPipe* p; ~Guard() { delete p; } } guard{gpp};
    GaussPipe& gp = *gpp;
    int64_t fix_at = 0, walk_at = 0;
    int refix_at = 0, refix_rows = 0;
    for (int e = 0; e < B; ++e) {
        const uint8_t f = flags[e];
        int32_t* inc = next_inc + 3 * e;
        inc[0] = inc[1] = inc[2] = 0;
        spatial_pick[e] = -1;
        if (f & DBSGYM_RESET_ELECTRODE_MOVE) {                       // env.py:485-498
            const int base = freq[3 * e];
            inc[0] = plan->random_freq_update ? base + ((int)bounded(s, 3) - 1) : base;      // calc_next_event(f, [-1, 0, 1])
            int32_t* c = elec_coords + 3 * e;
            int32_t moved[3];
            bool ok;
            do {                                                    // until every coordinate is inside [lo, hi]
                ok = true;
                for (int a = 0; a < 3; ++a) {
                    const int sgn = bounded(s, 2) ? 1 : -1;         // np.random.choice([-1, 1])
                    const int on = (int)bounded(s, 2);              // np.random.choice([0, 1])
                    moved[a] = c[a] + sgn * on;
                    if (moved[a] < plan->coord_lo || moved[a] > plan->coord_hi) ok = false;
                }
            } while (!ok);
            c[0] = moved[0]; c[1] = moved[1]; c[2] = moved[2];
        }
        if (f & DBSGYM_RESET_ENCAPSULATION) {                        // env.py:506-509: calc_next_event(f, [-2 .. 2])
            const int base = freq[3 * e + 1];
            inc[1] = plan->random_freq_update ? base + ((int)bounded(s, 5) - 2) : base;
        }
        if (f & DBSGYM_RESET_PLASTICITY) {                           // env.py:519-523: calc_next_temp_event(f, [0, 1])
            const int base = freq[3 * e + 2];
            inc[2] = plan->random_freq_update ? base + (int)bounded(s, 2) : base;
        }
        if (f & DBSGYM_RESET_WALK_REGEN) {                           // env.py:532-541 -> generate_perturbations, env.py:21-57
            if (!walk_noise) return DBSGYM_EINVAL;
            double* w = walk_noise + walk_at;
            for (int64_t i = 0; i < (int64_t)M * N; ++i) gp.next(w + i, false);
            walk_at += (int64_t)M * N;
        }
        if (f & DBSGYM_RESET_SPATIAL) spatial_pick[e] = (int32_t)bounded(s, (uint32_t)plan->table_len);     // env.py:544-552
        for (int i = 0; i < n_fix[e]; ++i) {                         // utils.py:819-823 remove_negative_w0(w0): randn(k)
            if (!fix_noise) return DBSGYM_EINVAL;
            gp.next(fix_noise + fix_at++, false);
        }
        double* y = init_state + (size_t)e * N;                     // env.py:595: np.random.normal(mean, sd, N)
        int bad = 0;
        for (int i = 0; i < N; ++i) bad += gp.next(y + i, true);
        if (bad) {                                                  // env.py:598 remove_negative_w0(init_state): randn(bad)
            if (!refix_env || !refix_noise || refix_rows >= plan->refix_cap_rows || refix_at + bad > plan->refix_cap_noise)
                return DBSGYM_ESTATE;                               // (caller's stream untouched: it takes the per-env path)
            refix_env[2 * refix_rows] = e; refix_env[2 * refix_rows + 1] = bad;
            ++refix_rows;
            for (int i = 0; i < bad; ++i) gp.next(refix_noise + refix_at++, false);
        }
    }
    gp.finish();
    *n_refix = refix_rows;
    work.pos = s.pos; work.has_gauss = s.has_gauss; work.gauss = s.gauss;
    *st = work;
    return DBSGYM_OK;
}

}  // extern "C"
