// Device-resident N-vector algebra for the PCG / IRLS / AI-REML loops (replaces the Armadillo
// expressions of saige_fitnull.cpp:581-862).  All reductions are deterministic: per-block partial
// sums in a fixed tree, then the last block to finish adds the partials in block order.
#pragma once
#include "ctx.h"

namespace sgb {

constexpr int kMaxPairs = 48;
struct DotArgs {
    const double *a[kMaxPairs];
    const double *b[kMaxPairs];
    int q;
};
constexpr int kMaxCoef = 40;
struct LinArgs {
    double coef[kMaxCoef];
    int n;
};

// out_host[q] = sum_n a_q[n] * b_q[n]   (q pairs; synchronises the stream)
void dot_pairs(Context &c, const std::vector<const double *> &a, const std::vector<const double *> &b, double *out_host);
// out[n] = s * base[n] + sum_c coef[c] * cols[n + c*ld]     (base may be NULL)
void lincomb(Context &c, double *out, double s, const double *base, const double *cols, int64_t ld,
             const std::vector<double> &coef);
// family: mu = linkinv(eta), Y = eta - offset + (y - mu)/mu.eta(eta), W = mu.eta^2 / variance(mu)
//   (saige_fitnull.cpp:791-794, :804-807; R family.c for logit).  add_offset: eta += offset first (:803).
void family_update(Context &c, int family, double *eta, const double *offset, const double *y, double *mu, double *Y,
                   double *W, bool add_offset);
// W only, from given eta and mu (saige_fitnull.cpp:1281-1284)
void family_weights(Context &c, int family, const double *eta, const double *mu, double *W);
// eta = Y - tau0 * (Sigma_iY - Sigma_iX alpha) / w      (saige_fitnull.cpp:757)
void eta_update(Context &c, double *eta, const double *Y, const double *Sigma_iY, const double *Sigma_iX, int64_t ld,
                const std::vector<double> &alpha, double tau0, const double *w);
// minv = 1 / max(tau0 / w + tau1 * diag, 1e-4)           (get_diag_sigma :542-558 and :587)
void diag_sigma(Context &c, const double *w, double tau0, double tau1, double *out, bool invert);
void expand_rademacher(Context &c, const int8_t *bits_device, double *out, int64_t count);  // 0/1 -> -1/+1  (:649)
// var2 terms: out_host[c] = sum_n wgt[n] * (G[n,c])^2 / ac[c]    (wgt == NULL -> 1)   (:1325 / :1436)
void weighted_sumsq(Context &c, const double *wgt, const double *G, int64_t ld, int k, double *out_host);
// dosage post-processing of f64_af_ac_impute + flip (vectorization.cpp:186-205, saige_fitnull.cpp:1310-1315)
void impute_flip(Context &c, double *G0, double impute_value, bool flip);

// ---- batched PCG state machine (PCG_diag_sigma, saige_fitnull.cpp:581-614) ----
struct PcgWork {
    DevBuf<double> minv, r, z, p, Ap, gp;
    DevBuf<double> scal;      // [4][K]: rz, rz_new, pAp, rr
    DevBuf<int> cols;         // active column list
    DevBuf<double> partial;
    DevBuf<unsigned int> counter;
};
void pcg_solve(Context &c, PcgWork &ws, const double *w, double tau0, double tau1, const double *b, int k, int maxiter,
               double tol, double *x, int *iters_host);

}  // namespace sgb
