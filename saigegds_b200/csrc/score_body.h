// Single-variant score test with saddle-point approximation: the per-variant body, written once for the device.
// Replaces single_test_quant / single_test_bin (src/saige_main.cpp:188-407), the SPA routines Korg, K1_adj, K2,
// getroot_K1_fast, get_saddle_prob_fast, Saddle_Prob_Fast (src/SPATest.cpp:39-374) and f64_af_ac_impute
// (src/vectorization.cpp:186-205).
//
// One thread block works on one variant; `Env` supplies the block (thread index, sums over the block, an exclusive scan)
// so that the same source also compiles for one host thread in tests/native/score_body_check.cpp, where it is compared
// with the oracle before it ever sees a GPU.  Everything a thread decides on comes out of a block-wide sum that all
// threads receive bit-identically, so the control flow is uniform.
//
// The reference has two algebraically equal branches (dense for MAF >= 0.05, index lists below).  Here there is one: with
// coef = (X'VX)^-1 X'V G and B = X coef, a single pass over the samples with G != 0 gives
//   S = sum (y-mu)(G-B) = G'(y-mu) - S_a' coef,      var2 = sum w (G-B)^2 = G'WG - 2 coef'(X'WG) + coef' (X'WX) coef,
// (2K + 3 model values per visited sample, no second pass), which are the numbers of saige_main.cpp:218-262 / :322-350.  The SPA step needs
// the adjusted genotype of all samples only for the two one-sided sums g_pos / g_neg (SPATest.cpp:320-325) and for the
// samples with G != 0; q, m1 and var2 follow from the sums above.
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SGB_HD __host__ __device__ __forceinline__
#else
#define SGB_HD inline
#endif

namespace sgb {
namespace score {

constexpr int kOutCols = 8;   // AF, mac, num, beta, SE, pval, pval_noadj, converged  (the vector built at saige_main.cpp:277-284 / :398-406)

// The arrays of .init_nullmod (R/assoc_single.r:17-67) the test reads; K x n matrices in R's column-major order = [n][K].
struct Model {
    int trait;                  // 0 binary, 1 quantitative
    int64_t n;
    int K;
    double tau0;
    const double *y_mu, *mu, *mu2;          // [n]
    const double *t_XVX_inv_XV, *t_X;       // [n][K]
    const double *XVX;                      // [K][K]
    const double *S_a;                      // [K] = colSums(X * (y - mu))
    const double *X_mu;                     // [K] = colSums(X * mu)   (derived at init)
    double varRatio, thr_maf, thr_mac, thr_missing, thr_pval_spa;
};

SGB_HD double sq(double v) { return v * v; }
SGB_HD int sign(double v) { return (v > 0) ? 1 : ((v < 0) ? -1 : 0); }
SGB_HD double nan_value() { return NAN; }

// R's distribution functions for one degree of freedom, closed forms
SGB_HD double pchisq1_upper(double x) { return isnan(x) ? x : (x <= 0 ? 1.0 : erfc(sqrt(x * 0.5))); }
SGB_HD double pnorm_lower(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }
SGB_HD double pnorm_upper(double z) { return 0.5 * erfc(z * 0.70710678118654752440); }

// qnorm(p): Wichura's algorithm AS 241 (PPND16), the one R's qnorm5 uses
SGB_HD double qnorm_as241(double p) {
    if (isnan(p) || p < 0 || p > 1) return nan_value();
    if (p == 0) return -INFINITY;
    if (p == 1) return INFINITY;
    const double q = p - 0.5;
    double r, val;
    if (fabs(q) <= 0.425) {
        r = .180625 - q * q;
        return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                        45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                      133.14166789178437745) * r + 3.387132872796366608) /
               (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                    21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                 42.313330701600911252) * r + 1.);
    }
    r = (q < 0) ? p : 1 - p;
    r = sqrt(-log(r));
    if (r <= 5.) {
        r += -1.6;
        val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
                   1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
                4.6303378461565452959) * r + 1.42343711074968357734) /
              (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
                   .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
                2.05319162663775882187) * r + 1.);
    } else {
        r += -5.;
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
                   .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
                5.4637849111641143699) * r + 6.6579046435011037772) /
              (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                   7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
                .59983224954277312477) * r + 1.);
    }
    return (q < 0.0) ? -val : val;
}

// ---- genotype sources: value of sample i, NaN = missing -----------------------------------------------------------
struct PackedRow {   // 2-bit codes, sample 4j+k in bits 2k..2k+1 of byte j, 3 = missing (the store's own format)
    const uint8_t *row;
    SGB_HD double operator()(int64_t i) const {
        const unsigned c = (row[i >> 2] >> (2 * (int)(i & 3))) & 3u;
        return c < 3 ? (double)c : nan_value();
    }
};
struct DosageRow {   // REALSXP dosages as get_ds hands them on (saige_main.cpp:166-186)
    const double *row;
    SGB_HD double operator()(int64_t i) const {
        const double v = row[i];
        return isfinite(v) ? v : nan_value();
    }
};

// Genotype after mean imputation and the flip to the minor allele (vectorization.cpp:198-203, saige_main.cpp:208-213)
template <class Geno>
struct Coded {
    Geno g;
    double imputed;
    bool minus;
    SGB_HD double operator()(int64_t i) const {
        double v = g(i);
        if (isnan(v)) v = imputed;
        return minus ? 2 - v : v;
    }
};

// Samples are dealt to threads in groups of four (one packed byte, one 32-byte sector of every model vector); the
// assignment is the same in every pass, which is what makes the compaction offsets of pass 1 valid in the SPA pass.
#define SGB_SCORE_FOR_SAMPLES(env, n, i)                                                         \
    for (int64_t _g = (env).tid(), _ng = ((n) + 3) >> 2; _g < _ng; _g += (env).nthr())           \
        for (int64_t i = _g << 2, _e = (i + 4 < (n)) ? i + 4 : (n); i < _e; i++)

// ---- SPATest.cpp:39-80 on the compacted (g, mu) pairs of the samples with G != 0 --------------------------------------
template <class Env>
SGB_HD double Korg(Env &env, double t, int64_t nnz, const double *g, const double *mu) {
    double s = 0;
#pragma unroll 4
    for (int64_t k = env.tid(); k < nnz; k += env.nthr()) s += log(1 - mu[k] + mu[k] * exp(g[k] * t));
    return env.sum(s);
}
template <class Env>
SGB_HD double K1_adj(Env &env, double t, int64_t nnz, const double *g, const double *mu, double q) {
    double s = 0;
    for (int64_t k = env.tid(); k < nnz; k += env.nthr()) s += mu[k] * g[k] / ((1 - mu[k]) * exp(-g[k] * t) + mu[k]);
    return env.sum(s) - q;
}
template <class Env>
SGB_HD double K2(Env &env, double t, int64_t nnz, const double *g, const double *mu) {
    double s = 0;
#pragma unroll 4
    for (int64_t k = env.tid(); k < nnz; k += env.nthr()) {
        const double m = mu[k], om = 1 - m, gi = g[k], e = exp(-gi * t);
        const double v = (om * m * gi * gi * e) / sq(om * e + m);
        if (isfinite(v)) s += v;
    }
    return env.sum(s);
}

// K1_adj and K2 at the same t from one exponential per sample (Newton's step needs both; SPATest.cpp:143-150 evaluates
// them in two passes).  K1 term = mu g / d, K2 term = (1-mu) mu g^2 e / d^2 with e = exp(-g t), d = (1-mu) e + mu.
template <class Env>
SGB_HD void K1_K2(Env &env, double t, int64_t nnz, const double *g, const double *mu, double q, double &k1, double &k2) {
    double s1 = 0, s2 = 0;
#pragma unroll 4   // several (g, mu) pairs in flight: the loop is bound by the latency of these loads otherwise
    for (int64_t k = env.tid(); k < nnz; k += env.nthr()) {
        const double m = mu[k], om = 1 - m, gi = g[k], e = exp(-gi * t);
        const double inv = 1 / (om * e + m), mg = m * gi * inv;
        s1 += mg;
        const double v = om * e * gi * inv * mg;
        if (isfinite(v)) s2 += v;
    }
    k1 = env.sum(s1) - q;
    k2 = env.sum(s2);
}

// SPATest.cpp:134-181
template <class Env>
SGB_HD void getroot_K1_fast(Env &env, double g_pos, double g_neg, double &root, bool &converged, int64_t nnz, const double *g,
                            const double *mu, double q, double NAmu, double NAsigma) {
    const double root_tol = 1.220703125e-4;   // DBL_EPSILON ^ 0.25, SPATest.cpp:26
    if (q >= g_pos || q <= g_neg) {
        root = INFINITY;
        converged = true;
        return;
    }
    double t = root = 0;
    double k1, k2;
    K1_K2(env, t, nnz, g, mu, q, k1, k2);
    double K1_eval = k1 + NAmu + NAsigma * t;
    double prevJump = INFINITY;
    converged = false;
    for (int it = 1; it <= 1000; it++) {
        const double K2_eval = k2 + NAsigma;          // K2 at the current t, from the pass that produced K1_eval
        double tnew = t - K1_eval / K2_eval;
        if (!isfinite(tnew)) break;
        if (fabs(tnew - t) < root_tol) {
            converged = true;
            break;
        }
        K1_K2(env, tnew, nnz, g, mu, q, k1, k2);
        double newK1 = k1 + NAmu + NAsigma * tnew;
        if (sign(K1_eval) != sign(newK1)) {
            if (fabs(tnew - t) > prevJump - root_tol) {
                tnew = t + sign(newK1 - K1_eval) * prevJump * 0.5;
                K1_K2(env, tnew, nnz, g, mu, q, k1, k2);
                newK1 = k1 + NAmu + NAsigma * tnew;
                prevJump *= 0.5;
            } else {
                prevJump = fabs(tnew - t);
            }
        }
        root = t = tnew;
        K1_eval = newK1;
    }
}

// SPATest.cpp:210-230
template <class Env>
SGB_HD double get_saddle_prob_fast(Env &env, double t, int64_t nnz, const double *g, const double *mu, double q, double NAmu,
                                   double NAsigma) {
    if (!isfinite(t)) return 0;
    const double K = Korg(env, t, nnz, g, mu) + NAmu * t + 0.5 * NAsigma * t * t;
    const double k2 = K2(env, t, nnz, g, mu) + NAsigma;
    double pval = 0;
    if (isfinite(K) && isfinite(k2)) {
        const double w = sign(t) * sqrt(2 * (t * q - K));
        const double v = t * sqrt(k2);
        const double z = w + log(v / w) / w;
        pval = (z > 0) ? pnorm_upper(z) : -pnorm_lower(z);
    }
    return pval;
}

// Saddle_Prob_Fast, SPATest.cpp:298-374, after its set-up loop: the two one-sided sums g_pos / g_neg over all samples and
// the (g, mu) pairs of the samples kept exactly, the rest summarised by the normal part (NAmu, NAsigma).  With every
// sample kept and NAmu = NAsigma = 0 this is Saddle_Prob, SPATest.cpp:232-296.  p_noadj: the normal-approximation p-value.
template <class Env>
SGB_HD double saddle_prob(Env &env, double q, double m1, double var1, double g_pos, double g_neg, int64_t nnz, const double *g,
                          const double *mu, double NAmu, double NAsigma, double cutoff, bool &converged, double &p_noadj) {
    const double sdiff = q - m1, qinv = -sdiff + m1;
    p_noadj = pchisq1_upper(sdiff * sdiff / var1);
    double pval;
    while (true) {
        converged = true;
        if (cutoff < 0.1) cutoff = 0.1;
        if (fabs(q - m1) / sqrt(var1) < cutoff) {
            pval = p_noadj;
        } else {
            double root1, root2;
            bool conv1, conv2;
            getroot_K1_fast(env, g_pos, g_neg, root1, conv1, nnz, g, mu, q, NAmu, NAsigma);
            getroot_K1_fast(env, g_pos, g_neg, root2, conv2, nnz, g, mu, qinv, NAmu, NAsigma);
            if (conv1 && conv2) {
                const double p1 = get_saddle_prob_fast(env, root1, nnz, g, mu, q, NAmu, NAsigma);
                const double p2 = get_saddle_prob_fast(env, root2, nnz, g, mu, qinv, NAmu, NAsigma);
                pval = fabs(p1) + fabs(p2);
            } else {
                pval = p_noadj;
                converged = false;
                break;
            }
        }
        if (pval != 0 && p_noadj / pval > 1000)
            cutoff *= 2;
        else
            break;
    }
    return pval;
}

// ---- the same, both roots at once ------------------------------------------------------------------------------------------------------
// Saddle_Prob_Fast solves K1(t) = q and K1(t) = qinv by two independent Newton iterations; each step is a pass over the (g, mu) pairs.
// The sums of a pass depend on t only, so one pass serves the current step of both iterations (and the first pass, at t = 0, is the
// same for both): the pairs are read half as often.  Every root follows exactly the sequence of getroot_K1_fast.
struct RootIter {
    double q, t, root, K1_eval, k2, prevJump, tnew, t_eval;
    int it, phase;          // phase 0: wants the sums at t = 0; 1: at the Newton step tnew; 2: at the halved jump; 3: finished
    bool converged;
    SGB_HD void start(double q_, double g_pos, double g_neg) {
        q = q_; t = root = 0; K1_eval = k2 = tnew = t_eval = 0; prevJump = INFINITY; it = 0; converged = false; phase = 0;
        if (q >= g_pos || q <= g_neg) { root = INFINITY; converged = true; phase = 3; }
    }
    SGB_HD bool active() const { return phase != 3; }
    // top of an iteration of the loop at SPATest.cpp:143
    SGB_HD void advance(double NAsigma) {
        const double root_tol = 1.220703125e-4;
        if (++it > 1000) { phase = 3; return; }
        const double K2_eval = k2 + NAsigma;
        tnew = t - K1_eval / K2_eval;
        if (!isfinite(tnew)) { phase = 3; return; }
        if (fabs(tnew - t) < root_tol) { converged = true; phase = 3; return; }
        phase = 1;
        t_eval = tnew;
    }
    // s1, s2: the raw sums of K1_K2 at t_eval
    SGB_HD void consume(double s1, double s2, double NAmu, double NAsigma) {
        const double root_tol = 1.220703125e-4;
        const double k1 = s1 - q;
        if (phase == 0) {
            K1_eval = k1 + NAmu + NAsigma * t;
            k2 = s2;
            advance(NAsigma);
            return;
        }
        const double newK1 = k1 + NAmu + NAsigma * tnew;
        k2 = s2;
        if (phase == 1) {
            if (sign(K1_eval) != sign(newK1)) {
                if (fabs(tnew - t) > prevJump - root_tol) {
                    tnew = t + sign(newK1 - K1_eval) * prevJump * 0.5;
                    phase = 2;
                    t_eval = tnew;
                    return;
                }
                prevJump = fabs(tnew - t);
            }
        } else {
            prevJump *= 0.5;
        }
        root = t = tnew;
        K1_eval = newK1;
        advance(NAsigma);
    }
};

// The loops below read the next (g, mu) pair of the thread before they work on the current one: a pair costs two or four
// exponentials, during which the next load is in flight (the plain loops sat in the latency of these loads with 16 warps per SM).
// The order of a thread's additions is unchanged.
template <class Env>
SGB_HD void K1_K2_dual(Env &env, double ta, bool on_a, double tb, bool on_b, int64_t nnz, const double *g, const double *mu,
                       double &s1a, double &s2a, double &s1b, double &s2b) {
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
    const int64_t step = env.nthr();
    int64_t k = env.tid();
    double gn = 0, mn = 0;
    if (k < nnz) { gn = g[k]; mn = mu[k]; }
    if (on_a && on_b && ta == tb) {          // the first pass (t = 0): one evaluation serves both
        while (k < nnz) {
            const double m = mn, om = 1 - m, gi = gn;
            k += step;
            if (k < nnz) { gn = g[k]; mn = mu[k]; }
            const double e = exp(-gi * ta);
            const double inv = 1 / (om * e + m), mg = m * gi * inv;
            a1 += mg;
            const double v = om * e * gi * inv * mg;
            if (isfinite(v)) a2 += v;
        }
        s1a = s1b = env.sum(a1);
        s2a = s2b = env.sum(a2);
        return;
    }
    while (k < nnz) {
        const double m = mn, om = 1 - m, gi = gn;
        k += step;
        if (k < nnz) { gn = g[k]; mn = mu[k]; }
        if (on_a) {
            const double e = exp(-gi * ta), inv = 1 / (om * e + m), mg = m * gi * inv;
            a1 += mg;
            const double v = om * e * gi * inv * mg;
            if (isfinite(v)) a2 += v;
        }
        if (on_b) {
            const double e = exp(-gi * tb), inv = 1 / (om * e + m), mg = m * gi * inv;
            b1 += mg;
            const double v = om * e * gi * inv * mg;
            if (isfinite(v)) b2 += v;
        }
    }
    if (on_a) { s1a = env.sum(a1); s2a = env.sum(a2); }
    if (on_b) { s1b = env.sum(b1); s2b = env.sum(b2); }
}

// Korg and K2 of get_saddle_prob_fast at both roots from one pass
template <class Env>
SGB_HD void saddle_sums_dual(Env &env, double ta, bool on_a, double tb, bool on_b, int64_t nnz, const double *g, const double *mu,
                             double &Ka, double &k2a, double &Kb, double &k2b) {
    double a0 = 0, a2 = 0, b0 = 0, b2 = 0;
    const int64_t step = env.nthr();
    int64_t k = env.tid();
    double gn = 0, mn = 0;
    if (k < nnz) { gn = g[k]; mn = mu[k]; }
    while (k < nnz) {
        const double m = mn, om = 1 - m, gi = gn;
        k += step;
        if (k < nnz) { gn = g[k]; mn = mu[k]; }
        if (on_a) {
            a0 += log(1 - m + m * exp(gi * ta));
            const double e = exp(-gi * ta);
            const double v = (om * m * gi * gi * e) / sq(om * e + m);
            if (isfinite(v)) a2 += v;
        }
        if (on_b) {
            b0 += log(1 - m + m * exp(gi * tb));
            const double e = exp(-gi * tb);
            const double v = (om * m * gi * gi * e) / sq(om * e + m);
            if (isfinite(v)) b2 += v;
        }
    }
    if (on_a) { Ka = env.sum(a0); k2a = env.sum(a2); }
    if (on_b) { Kb = env.sum(b0); k2b = env.sum(b2); }
}

SGB_HD double saddle_pval_from_sums(double t, double Korg_t, double K2_t, double q, double NAmu, double NAsigma) {
    if (!isfinite(t)) return 0;
    const double K = Korg_t + NAmu * t + 0.5 * NAsigma * t * t;
    const double k2 = K2_t + NAsigma;
    double pval = 0;
    if (isfinite(K) && isfinite(k2)) {
        const double w = sign(t) * sqrt(2 * (t * q - K));
        const double v = t * sqrt(k2);
        const double z = w + log(v / w) / w;
        pval = (z > 0) ? pnorm_upper(z) : -pnorm_lower(z);
    }
    return pval;
}

// saddle_prob with the two roots advanced together.  The roots do not depend on the cutoff, so the cutoff-doubling loop of
// SPATest.cpp:298-374 is replayed on the values computed once.  Same result as saddle_prob (bit-identical for a one-thread Env).
template <class Env>
SGB_HD double saddle_prob_dual(Env &env, double q, double m1, double var1, double g_pos, double g_neg, int64_t nnz, const double *g,
                               const double *mu, double NAmu, double NAsigma, double cutoff, bool &converged, double &p_noadj) {
    const double sdiff = q - m1, qinv = -sdiff + m1;
    p_noadj = pchisq1_upper(sdiff * sdiff / var1);
    bool have = false, conv = false;
    double pval_roots = 0;
    double pval;
    while (true) {
        converged = true;
        if (cutoff < 0.1) cutoff = 0.1;
        if (fabs(q - m1) / sqrt(var1) < cutoff) {
            pval = p_noadj;
        } else {
            if (!have) {
                RootIter A, B;
                A.start(q, g_pos, g_neg);
                B.start(qinv, g_pos, g_neg);
                while (A.active() || B.active()) {
                    double s1a = 0, s2a = 0, s1b = 0, s2b = 0;
                    const bool on_a = A.active(), on_b = B.active();
                    K1_K2_dual(env, A.t_eval, on_a, B.t_eval, on_b, nnz, g, mu, s1a, s2a, s1b, s2b);
                    if (on_a) A.consume(s1a, s2a, NAmu, NAsigma);
                    if (on_b) B.consume(s1b, s2b, NAmu, NAsigma);
                }
                conv = A.converged && B.converged;
                if (conv) {
                    const bool fa = isfinite(A.root), fb = isfinite(B.root);
                    double Ka = 0, k2a = 0, Kb = 0, k2b = 0;
                    if (fa || fb) saddle_sums_dual(env, A.root, fa, B.root, fb, nnz, g, mu, Ka, k2a, Kb, k2b);
                    const double p1 = saddle_pval_from_sums(A.root, Ka, k2a, q, NAmu, NAsigma);
                    const double p2 = saddle_pval_from_sums(B.root, Kb, k2b, qinv, NAmu, NAsigma);
                    pval_roots = fabs(p1) + fabs(p2);
                }
                have = true;
            }
            if (conv) {
                pval = pval_roots;
            } else {
                pval = p_noadj;
                converged = false;
                break;
            }
        }
        if (pval != 0 && p_noadj / pval > 1000)
            cutoff *= 2;
        else
            break;
    }
    return pval;
}

// Filters of saige_main.cpp:197-205 / :297-305 from the allele count AC over Num called samples.
SGB_HD bool variant_passes(const Model &M, double AC, int Num, double &AF, double &mac) {
    AF = (Num > 0) ? (AC / (2 * Num)) : nan_value();
    const double maf = fmin(AF, 1 - AF);
    mac = fmin(AC, 2 * Num - AC);
    const double missing = double(M.n - Num) / M.n;
    return (Num > 0) && (maf > 0) && (maf >= M.thr_maf) && (mac >= M.thr_mac) && (missing <= M.thr_missing);
}

// Score statistic and its normal-approximation p-value from the per-variant sums over the samples with G != 0:
// coef = sum G a_i (a = row of t_XVX_inv_XV), xwg = sum G w_i x_i, SyG = sum G (y-mu)_i, SwGG = sum G^2 w_i.
// With B = X coef:  sum_i (y-mu)_i (G-B)_i = G'(y-mu) - S_a'coef  and
// sum_i w_i (G-B)_i^2 = G'WG - 2 coef'(X'WG) + coef'(X'WX)coef.  (W = mu(1-mu) of the mixed model, while coef is weighted
// with the V of the fixed-effects-only fit, so the cross term does not collapse.)  The reference reaches the same two
// numbers through per-sample residuals, saige_main.cpp:218-262 / :322-350.
template <int KMAX>
SGB_HD void score_stats(const Model &M, const double (&coef)[KMAX], const double (&xwg)[KMAX], double SyG, double SwGG, double mac,
                        double &S, double &var2, double &coef_xmu, double &pval_noadj, double &beta) {
    const int K = M.K;
    double quad = 0, cross = 0, sa_coef = 0;
    coef_xmu = 0;
#pragma unroll
    for (int a = 0; a < KMAX; a++)
        if (a < K) {
            double r = 0;
#pragma unroll
            for (int b = 0; b < KMAX; b++)
                if (b < K) r += coef[b] * M.XVX[a * K + b];
            quad += coef[a] * r;
            cross += coef[a] * xwg[a];
            sa_coef += M.S_a[a] * coef[a];
            coef_xmu += coef[a] * M.X_mu[a];
        }
    var2 = SwGG - 2 * cross + quad;
    S = SyG - sa_coef;
    const double inv_sqrt_mac = 1.0 / sqrt(mac), inv_mac = 1.0 / mac;
    if (M.trait == 0) {
        const double var1 = var2 * M.varRatio;
        pval_noadj = pchisq1_upper(S * S / var1);
        beta = S / var1;
    } else {
        const double var1 = var2 * inv_mac * M.varRatio;
        const double Tstat = S * inv_sqrt_mac / M.tau0;
        pval_noadj = pchisq1_upper(Tstat * Tstat / var1);
        beta = Tstat / var1 * inv_sqrt_mac;
    }
}

// The saddle-point step of one variant whose normal-approximation p-value calls for it (saige_main.cpp:353-394): G = the coded
// genotypes, coef = (X'VX)^-1 X'V G, my_nnz = this thread's number of samples with G != 0 under SGB_SCORE_FOR_SAMPLES, the rest
// as score_stats returned them.  DUAL: saddle_prob_dual.  Replaces pval, beta, converged.
template <int KMAX, bool DUAL, class Env, class CodedGeno>
SGB_HD void spa_adjust(Env &env, const Model &M, const CodedGeno &G, const double (&coef)[KMAX], int my_nnz, double AC, int Num, bool minus,
                       double S, double var2, double coef_xmu, double gmu, double pval_noadj, double *spa_g, double *spa_mu, double &pval,
                       double &beta, bool &converged) {
    const int64_t n = M.n;
    const int K = M.K;
    const double AC2 = minus ? (2 * Num - AC) : AC;
    const double sc = 1 / sqrt(AC2);
    // adjusted genotype g = (G - B) / sqrt(AC2):  q - m1 = sum (y-mu) g,  m1 = sum mu g,  var2 = sum mu(1-mu) g^2
    const double m1 = (gmu - coef_xmu) * sc;
    const double svar2 = var2 * sc * sc, svar1 = svar2 * M.varRatio;
    const double Tstat = S * sc;
    const double q = Tstat / sqrt(svar1) * sqrt(svar2) + m1;      // "qtilde"
    // one pass over all samples: one-sided sums, and the (g, mu) pairs of the samples with G != 0, compacted in
    // thread order (offset = exclusive scan of the per-thread counts of pass 1)
    int64_t nnz = 0;
    int64_t at = env.excl_scan(my_nnz, nnz);
    double g_pos = 0, g_neg = 0, sub_mu = 0, sub_sigma = 0;
    SGB_SCORE_FOR_SAMPLES(env, n, i) {
        const double v = G(i);
        const double *x = M.t_X + (size_t)i * K;
        double B = 0;
#pragma unroll
        for (int c = 0; c < KMAX; c++)
            if (c < K) B += coef[c] * x[c];
        const double g = (v - B) * sc;
        if (g > 0) g_pos += g; else g_neg += g;
        if (v != 0) {
            const double m = M.mu[i];
            spa_g[at] = g;
            spa_mu[at] = m;
            at++;
            sub_mu += g * m;
            sub_sigma += g * g * m * (1 - m);
        }
    }
    g_pos = env.sum(g_pos);
    g_neg = env.sum(g_neg);
    const double NAmu = m1 - env.sum(sub_mu);
    const double NAsigma = svar2 - env.sum(sub_sigma);
    env.sync();   // the compacted pairs are read by other threads from here on

    // Saddle_Prob_Fast with (q, m1, var1) = (qtilde, m1, svar2) and cutoff 2 (saige_main.cpp:386-388)
    double p_na;
    pval = DUAL ? saddle_prob_dual(env, q, m1, svar2, g_pos, g_neg, nnz, spa_g, spa_mu, NAmu, NAsigma, 2.0, converged, p_na)
                : saddle_prob(env, q, m1, svar2, g_pos, g_neg, nnz, spa_g, spa_mu, NAmu, NAsigma, 2.0, converged, p_na);
    if (pval == 0 && pval_noadj > 0) {
        pval = pval_noadj;
        converged = false;
    }
    beta = (Tstat / svar1) / sqrt(AC2);
    env.sync();   // scratch is reused by the block's next variant
}

// One variant.  KMAX >= M.K bounds the per-thread coefficient registers.  spa_g / spa_mu: scratch of n doubles each, owned
// by this block.  out: kOutCols doubles; returns (to every thread) whether the variant passed the filters.
template <int KMAX, class Env, class Geno, bool DUAL = false>
SGB_HD bool test_variant(Env &env, const Model &M, const Geno &geno, double *spa_g, double *spa_mu, double *out) {
    const int64_t n = M.n;
    const int K = M.K;
    const bool bin = (M.trait == 0);

    // ---- f64_af_ac_impute: allele frequency, allele count, number of calls
    double s = 0;
    int cnt = 0;
    SGB_SCORE_FOR_SAMPLES(env, n, i) {
        const double v = geno(i);
        if (!isnan(v)) { s += v; cnt++; }
    }
    const double AC = env.sum(s);
    const int Num = (int)env.sum((double)cnt);
    double AF, mac;
    if (!variant_passes(M, AC, Num, AF, mac)) {
        if (env.tid() == 0)
            for (int k = 0; k < kOutCols; k++) out[k] = nan_value();
        return false;
    }
    const bool minus = (AF > 0.5);
    const Coded<Geno> G{geno, AF * 2, minus};

    // ---- one pass over the samples with G != 0: coef = (X'VX)^-1 X'V G, G'(y-mu), G'WG, G'mu
    double coef[KMAX], xwg[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; c++) coef[c] = xwg[c] = 0;
    double SyG = 0, SwGG = 0, gmu = 0;
    int my_nnz = 0;
    SGB_SCORE_FOR_SAMPLES(env, n, i) {
        const double v = G(i);
        if (v != 0) {
            my_nnz++;
            const double *a = M.t_XVX_inv_XV + (size_t)i * K, *x = M.t_X + (size_t)i * K;
            const double vw = bin ? v * M.mu2[i] : v;
#pragma unroll
            for (int c = 0; c < KMAX; c++)
                if (c < K) { coef[c] += v * a[c]; xwg[c] += vw * x[c]; }
            SyG += v * M.y_mu[i];
            SwGG += v * vw;
            gmu += v * M.mu[i];
        }
    }
#pragma unroll
    for (int c = 0; c < KMAX; c++)
        if (c < K) { coef[c] = env.sum(coef[c]); xwg[c] = env.sum(xwg[c]); }
    SyG = env.sum(SyG);
    SwGG = env.sum(SwGG);
    gmu = env.sum(gmu);
    double S, var2, coef_xmu, pval_noadj, beta;
    score_stats<KMAX>(M, coef, xwg, SyG, SwGG, mac, S, var2, coef_xmu, pval_noadj, beta);

    // ---- saddle-point approximation, saige_main.cpp:353-394 + Saddle_Prob_Fast
    double pval = pval_noadj;
    bool converged = isfinite(pval_noadj);
    if (bin && converged && (pval_noadj <= M.thr_pval_spa))
        spa_adjust<KMAX, DUAL>(env, M, G, coef, my_nnz, AC, Num, minus, S, var2, coef_xmu, gmu, pval_noadj, spa_g, spa_mu, pval, beta,
                               converged);
    if (minus) beta = -beta;
    if (env.tid() == 0) {
        out[0] = AF; out[1] = mac; out[2] = (double)Num; out[3] = beta;
        out[4] = fabs(beta / qnorm_as241(pval / 2));
        out[5] = pval; out[6] = pval_noadj; out[7] = converged ? 1.0 : 0.0;
    }
    return true;
}

}  // namespace score
}  // namespace sgb
