// CPU oracle, part 2: the single-variant score test with saddle-point approximation.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  A literal, R-free restatement of
//   src/saige_main.cpp:188-407      single_test_quant / single_test_bin
//   src/SPATest.cpp:39-374          Korg, K1_adj, K2, getroot_K1_fast, get_saddle_prob_fast, Saddle_Prob_Fast
//   src/vectorization.cpp:186-205   f64_af_ac_impute, and the small mat-vec helpers at :279-582
//   R/assoc_single.r:17-67          .init_nullmod (the derived model arrays; built by the Python caller)
// R's distribution functions are replaced by their closed forms for one degree of freedom:
//   pchisq(x, 1, lower=FALSE) = erfc(sqrt(x / 2)),  pnorm(z) = erfc(-z / sqrt 2) / 2,
//   qnorm(p) = Wichura's AS 241 (PPND16), the algorithm R's qnorm5 uses.
// It exists to check the <= 1e-6 p-value criterion: the null model fitted on the GPU and the one fitted by the
// oracle are both pushed through this test and compared, and the restatement itself is pinned against the
// reference's golden p-values (inst/unitTests/saige_pval.rds, saige_pval_quant.rds) in tests/test_oracle_golden.py.
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <limits>
#include <vector>

namespace {

inline double sq(double v) { return v * v; }
inline int sign(double v) { return (v > 0) ? 1 : ((v < 0) ? -1 : 0); }
const double kInf = std::numeric_limits<double>::infinity();
const double kNaN = std::numeric_limits<double>::quiet_NaN();

inline double pchisq1_upper(double x) { return std::isnan(x) ? kNaN : (x <= 0 ? 1.0 : erfc(sqrt(x * 0.5))); }
inline double pnorm_lower(double z) { return 0.5 * erfc(-z * M_SQRT1_2); }
inline double pnorm_upper(double z) { return 0.5 * erfc(z * M_SQRT1_2); }

// Wichura (1988) AS 241, PPND16: the algorithm behind R's qnorm5(p, 0, 1, lower=TRUE, log=FALSE)
double qnorm_as241(double p) {
    if (std::isnan(p) || p < 0 || p > 1) return kNaN;
    if (p == 0) return -kInf;
    if (p == 1) return kInf;
    const double q = p - 0.5;
    double r, val;
    if (fabs(q) <= 0.425) {
        r = .180625 - q * q;
        val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                       45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                     133.14166789178437745) * r + 3.387132872796366608) /
              (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                   21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                42.313330701600911252) * r + 1.);
        return val;
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

// ---- SPATest.cpp:39-80
double Korg(double t, size_t n, const double *mu, const double *g) {
    double sum = 0;
    for (size_t i = 0; i < n; i++) sum += log(1 - mu[i] + mu[i] * exp(g[i] * t));
    return sum;
}
double K1_adj(double t, size_t n, const double *mu, const double *g, double q) {
    double sum = 0;
    for (size_t i = 0; i < n; i++) sum += mu[i] * g[i] / ((1 - mu[i]) * exp(-g[i] * t) + mu[i]);
    return sum - q;
}
double K2(double t, size_t n, const double *mu, const double *g) {
    double sum = 0;
    for (size_t i = 0; i < n; i++) {
        const double m = mu[i], om = 1 - m, gi = g[i], e = exp(-gi * t);
        const double v = (om * m * gi * gi * e) / sq(om * e + m);
        if (std::isfinite(v)) sum += v;
    }
    return sum;
}

const double root_tol = sqrt(sqrt(DBL_EPSILON));
const int MaxNumIter = 1000;

// SPATest.cpp:134-181
void getroot_K1_fast(double g_pos, double g_neg, double &root, bool &converged, double init, size_t n_nonzero,
                     const double *mu, const double *g, double q, double NAmu, double NAsigma) {
    if (q >= g_pos || q <= g_neg) {
        root = kInf;
        converged = true;
        return;
    }
    double t = root = init;
    double K1_eval = K1_adj(t, n_nonzero, mu, g, q) + NAmu + NAsigma * t;
    double prevJump = kInf;
    converged = false;
    for (int it = 1; it <= MaxNumIter; it++) {
        const double K2_eval = K2(t, n_nonzero, mu, g) + NAsigma;
        double tnew = t - K1_eval / K2_eval;
        if (!std::isfinite(tnew)) break;
        if (fabs(tnew - t) < root_tol) {
            converged = true;
            break;
        }
        double newK1 = K1_adj(tnew, n_nonzero, mu, g, q) + NAmu + NAsigma * tnew;
        if (sign(K1_eval) != sign(newK1)) {
            if (fabs(tnew - t) > prevJump - root_tol) {
                tnew = t + sign(newK1 - K1_eval) * prevJump * 0.5;
                newK1 = K1_adj(tnew, n_nonzero, mu, g, q) + NAmu + NAsigma * tnew;
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
double get_saddle_prob_fast(double t, size_t n_nonzero, const double *mu, const double *g, double q, double NAmu,
                            double NAsigma) {
    if (!std::isfinite(t)) return 0;
    const double K = Korg(t, n_nonzero, mu, g) + NAmu * t + 0.5 * NAsigma * t * t;
    const double k2 = K2(t, n_nonzero, mu, g) + NAsigma;
    double pval = 0;
    if (std::isfinite(K) && std::isfinite(k2)) {
        const double w = sign(t) * sqrt(2 * (t * q - K));
        const double v = t * sqrt(k2);
        const double z = w + log(v / w) / w;
        pval = (z > 0) ? pnorm_upper(z) : -pnorm_lower(z);
    }
    return pval;
}

// SPATest.cpp:298-374
double Saddle_Prob_Fast(double q, double m1, double var1, size_t n_g, const double *mu, const double *g, size_t n_nonzero,
                        const int *nonzero_idx, double cutoff, bool &converged, double *buf_spa) {
    const double s = q - m1;
    const double qinv = -s + m1;
    const double pval_noadj = pchisq1_upper(s * s / var1);
    double pval;
    double NAmu = 0, NAsigma = 0, g_pos = 0, g_neg = 0;
    bool init = false;
    while (true) {
        converged = true;
        if (cutoff < 0.1) cutoff = 0.1;
        if (fabs(q - m1) / sqrt(var1) < cutoff) {
            pval = pval_noadj;
        } else {
            if (!init) {
                init = true;
                for (size_t i = 0; i < n_g; i++) {
                    const double v = g[i];
                    if (v > 0) g_pos += v; else g_neg += v;
                }
                NAmu = m1, NAsigma = var1;
                for (size_t i = 0; i < n_nonzero; i++) {
                    const size_t k = nonzero_idx[i];
                    const double g_k = g[k], mu_k = mu[k];
                    buf_spa[i] = g_k;
                    buf_spa[i + n_nonzero] = mu_k;
                    NAmu -= g_k * mu_k;
                    NAsigma -= g_k * g_k * mu_k * (1 - mu_k);
                }
                g = &buf_spa[0];
                mu = &buf_spa[n_nonzero];
            }
            double root1, root2;
            bool conv1, conv2;
            getroot_K1_fast(g_pos, g_neg, root1, conv1, 0, n_nonzero, mu, g, q, NAmu, NAsigma);
            getroot_K1_fast(g_pos, g_neg, root2, conv2, 0, n_nonzero, mu, g, qinv, NAmu, NAsigma);
            if (conv1 && conv2) {
                const double p1 = get_saddle_prob_fast(root1, n_nonzero, mu, g, q, NAmu, NAsigma);
                const double p2 = get_saddle_prob_fast(root2, n_nonzero, mu, g, qinv, NAmu, NAsigma);
                pval = fabs(p1) + fabs(p2);
            } else {
                pval = pval_noadj;
                converged = false;
                break;
            }
        }
        if (pval != 0 && pval_noadj / pval > 1000)
            cutoff *= 2;
        else
            break;
    }
    return pval;
}

// SPATest.cpp:90-131, the full saddle-point root finder (all samples, no normal part)
void getroot_K1(double g_pos, double g_neg, double &root, bool &converged, double init, size_t n_g, const double *mu,
                const double *g, double q) {
    if (q >= g_pos || q <= g_neg) {
        root = kInf;
        converged = true;
        return;
    }
    double t = root = init;
    double K1_eval = K1_adj(t, n_g, mu, g, q);
    double prevJump = kInf;
    converged = false;
    for (int it = 1; it <= MaxNumIter; it++) {
        const double K2_eval = K2(t, n_g, mu, g);
        double tnew = t - K1_eval / K2_eval;
        if (!std::isfinite(tnew)) break;
        if (fabs(tnew - t) < root_tol) {
            converged = true;
            break;
        }
        double newK1 = K1_adj(tnew, n_g, mu, g, q);
        if (sign(K1_eval) != sign(newK1)) {
            if (fabs(tnew - t) > prevJump - root_tol) {
                tnew = t + sign(newK1 - K1_eval) * prevJump * 0.5;
                newK1 = K1_adj(tnew, n_g, mu, g, q);
                prevJump *= 0.5;
            } else {
                prevJump = fabs(tnew - t);
            }
        }
        root = t = tnew;
        K1_eval = newK1;
    }
}

// SPATest.cpp:185-207
double get_saddle_prob(double t, size_t n_g, const double *mu, const double *g, double q) {
    if (!std::isfinite(t)) return 0;
    const double K = Korg(t, n_g, mu, g);
    const double k2 = K2(t, n_g, mu, g);
    double pval = 0;
    if (std::isfinite(K) && std::isfinite(k2)) {
        const double w = sign(t) * sqrt(2 * (t * q - K));
        const double v = t * sqrt(k2);
        const double z = w + log(v / w) / w;
        pval = (z > 0) ? pnorm_upper(z) : -pnorm_lower(z);
    }
    return pval;
}

// SPATest.cpp:232-296
double Saddle_Prob(double q, double m1, double var1, size_t n_g, const double *mu, const double *g, double cutoff,
                   bool &converged, double *p_noadj) {
    const double s = q - m1;
    const double qinv = -s + m1;
    const double pval_noadj = pchisq1_upper(s * s / var1);
    double pval;
    double g_pos = 0, g_neg = 0;
    bool init = false;
    if (p_noadj) *p_noadj = pval_noadj;
    while (true) {
        converged = true;
        if (cutoff < 0.1) cutoff = 0.1;
        if (fabs(q - m1) / sqrt(var1) < cutoff) {
            pval = pval_noadj;
        } else {
            if (!init) {
                init = true;
                for (size_t i = 0; i < n_g; i++) {
                    const double v = g[i];
                    if (v > 0) g_pos += v; else g_neg += v;
                }
            }
            double root1, root2;
            bool conv1, conv2;
            getroot_K1(g_pos, g_neg, root1, conv1, 0, n_g, mu, g, q);
            getroot_K1(g_pos, g_neg, root2, conv2, 0, n_g, mu, g, qinv);
            if (conv1 && conv2) {
                const double p1 = get_saddle_prob(root1, n_g, mu, g, q);
                const double p2 = get_saddle_prob(root2, n_g, mu, g, qinv);
                pval = fabs(p1) + fabs(p2);
            } else {
                pval = pval_noadj;
                converged = false;
                break;
            }
        }
        if (pval != 0 && pval_noadj / pval > 1000)
            cutoff *= 2;
        else
            break;
    }
    return pval;
}

// The model arrays of .init_nullmod (R/assoc_single.r:17-67); all matrices are K x n, column-major (R layout)
struct Model {
    int trait;  // 0 binary, 1 quantitative
    size_t n;
    int K;
    const double *tau, *y, *mu, *y_mu, *mu2, *t_XXVX_inv, *XV, *t_XVX_inv_XV, *XVX, *t_X, *S_a;
    double varRatio, thr_maf, thr_mac, thr_missing, thr_pval_spa;
};

struct Out {
    double AF, mac, beta, SE, pval, pval_noadj;
    int num, converged;
};

// vectorization.cpp:186-205
void af_ac_impute(double *ds, size_t n, double &AF, double &AC, int &Num, std::vector<int> &idx) {
    double sum = 0;
    int num = 0;
    size_t nmiss = 0;
    for (size_t i = 0; i < n; i++) {
        if (std::isfinite(ds[i])) { sum += ds[i]; num++; } else idx[nmiss++] = (int)i;
    }
    AF = (num > 0) ? (sum / (2 * num)) : kNaN;
    AC = sum;
    Num = num;
    if (num < (int)n) {
        const double d = AF * 2;
        for (size_t k = 0; k < nmiss; k++) ds[idx[k]] = d;
    }
}

// src/saige_main.cpp:188-285 (quantitative) and :288-407 (binary)
bool single_test(const Model &M, double *G, Out &o, std::vector<int> &buf_index, std::vector<double> &buf_coeff,
                 std::vector<double> &buf_adj_g, std::vector<double> &buf_B, std::vector<double> &buf_g_tilde,
                 std::vector<double> &buf_X1, std::vector<double> &buf_spa) {
    const size_t n = M.n;
    const int K = M.K;
    double AF, AC;
    int Num;
    af_ac_impute(G, n, AF, AC, Num, buf_index);
    const double maf = std::min(AF, 1 - AF);
    const double mac = std::min(AC, 2 * Num - AC);
    const double missing = double(n - Num) / n;
    if (!((Num > 0) && (maf > 0) && (maf >= M.thr_maf) && (mac >= M.thr_mac) && (missing <= M.thr_missing))) return false;
    const bool minus = (AF > 0.5);
    if (minus) for (size_t i = 0; i < n; i++) G[i] = 2 - G[i];
    const bool is_sparse = maf < 0.05;
    const bool bin = (M.trait == 0);
    size_t nnz = 0;
    double pval_noadj, beta;
    const double inv_sqrt_mac = 1.0 / sqrt(mac), inv_mac = 1.0 / mac;
    auto adj_g_dense = [&](bool sparse_g) {
        // buf_coeff = XV * G ; buf_adj_g = G - XXVX_inv * buf_coeff
        std::fill(buf_coeff.begin(), buf_coeff.end(), 0.0);
        if (sparse_g) {
            for (size_t k = 0; k < nnz; k++) {
                const size_t i = buf_index[k];
                for (int c = 0; c < K; c++) buf_coeff[c] += G[i] * M.XV[(size_t)K * i + c];
            }
        } else {
            for (size_t i = 0; i < n; i++)
                if (G[i] != 0) for (int c = 0; c < K; c++) buf_coeff[c] += G[i] * M.XV[(size_t)K * i + c];
        }
        for (size_t i = 0; i < n; i++) {
            double s = 0;
            for (int c = 0; c < K; c++) s += buf_coeff[c] * M.t_XXVX_inv[(size_t)K * i + c];
            buf_adj_g[i] = G[i] - s;
        }
    };
    if (is_sparse) {
        for (size_t j = 0; j < n; j++) if (G[j] != 0) buf_index[nnz++] = (int)j;
        // buf_coeff = XVX_inv_XV * G
        std::fill(buf_coeff.begin(), buf_coeff.end(), 0.0);
        for (size_t k = 0; k < nnz; k++) {
            const size_t i = buf_index[k];
            for (int c = 0; c < K; c++) buf_coeff[c] += G[i] * M.t_XVX_inv_XV[(size_t)K * i + c];
        }
        // buf_B = t(X) * buf_coeff ; g_tilde = G - B
        for (size_t k = 0; k < nnz; k++) {
            const size_t i = buf_index[k];
            double s = 0;
            for (int c = 0; c < K; c++) s += buf_coeff[c] * M.t_X[(size_t)K * i + c];
            buf_B[k] = s;
            buf_g_tilde[k] = G[i] - s;
        }
        double var2 = 0;
        for (int a = 0; a < K; a++)
            for (int b = 0; b < K; b++) var2 += buf_coeff[a] * buf_coeff[b] * M.XVX[(size_t)K * a + b];
        for (size_t k = 0; k < nnz; k++) {
            const double d = sq(buf_g_tilde[k]) - sq(buf_B[k]);
            var2 += bin ? d * M.mu2[buf_index[k]] : d;
        }
        double S1 = 0;
        for (size_t k = 0; k < nnz; k++) S1 += M.y_mu[buf_index[k]] * buf_g_tilde[k];
        std::fill(buf_X1.begin(), buf_X1.end(), 0.0);
        for (size_t k = 0; k < nnz; k++) {
            const size_t i = buf_index[k];
            for (int c = 0; c < K; c++) buf_X1[c] += M.y_mu[i] * M.t_X[(size_t)K * i + c];
        }
        double S2 = 0;
        for (int c = 0; c < K; c++) S2 += (buf_X1[c] - M.S_a[c]) * buf_coeff[c];
        if (bin) {
            const double var1 = var2 * M.varRatio;
            const double S = S1 + S2;
            pval_noadj = pchisq1_upper(S * S / var1);
            beta = S / var1;
        } else {
            const double var1 = var2 * inv_mac * M.varRatio;
            const double Tstat = (S1 + S2) * inv_sqrt_mac / M.tau[0];
            pval_noadj = pchisq1_upper(Tstat * Tstat / var1);
            beta = Tstat / var1 * inv_sqrt_mac;
        }
    } else {
        adj_g_dense(false);
        double S = 0, var = 0;
        for (size_t i = 0; i < n; i++) {
            S += M.y_mu[i] * buf_adj_g[i];
            var += bin ? M.mu2[i] * buf_adj_g[i] * buf_adj_g[i] : buf_adj_g[i] * buf_adj_g[i];
        }
        if (bin) {
            var *= M.varRatio;
            pval_noadj = pchisq1_upper(S * S / var);
            beta = S / var;
        } else {
            const double Tstat = S * inv_sqrt_mac / M.tau[0];
            var *= inv_mac * M.varRatio;
            pval_noadj = pchisq1_upper(Tstat * Tstat / var);
            beta = Tstat / var * inv_sqrt_mac;
        }
    }
    double pval = pval_noadj;
    bool converged = std::isfinite(pval_noadj);
    if (bin && converged && (pval_noadj <= M.thr_pval_spa)) {
        if (is_sparse) adj_g_dense(true);
        const double AC2 = minus ? (2 * Num - AC) : AC;
        const double sc = 1 / sqrt(AC2);
        for (size_t i = 0; i < n; i++) buf_adj_g[i] *= sc;
        double q = 0, m1 = 0, var2 = 0;
        for (size_t i = 0; i < n; i++) q += M.y[i] * buf_adj_g[i];
        for (size_t i = 0; i < n; i++) {
            m1 += M.mu[i] * buf_adj_g[i];
            var2 += M.mu2[i] * buf_adj_g[i] * buf_adj_g[i];
        }
        const double var1 = var2 * M.varRatio;
        const double Tstat = q - m1;
        const double qtilde = Tstat / sqrt(var1) * sqrt(var2) + m1;
        if (!is_sparse) {
            nnz = 0;
            for (size_t j = 0; j < n; j++) if (G[j] != 0) buf_index[nnz++] = (int)j;
        }
        pval = Saddle_Prob_Fast(qtilde, m1, var2, n, M.mu, buf_adj_g.data(), nnz, buf_index.data(), 2, converged, buf_spa.data());
        if (pval == 0 && pval_noadj > 0) {
            pval = pval_noadj;
            converged = false;
        }
        beta = (Tstat / var1) / sqrt(AC2);
    }
    if (minus) beta = -beta;
    const double SE = fabs(beta / qnorm_as241(pval / 2));
    o.AF = AF; o.mac = mac; o.num = Num; o.beta = beta; o.SE = SE; o.pval = pval; o.pval_noadj = pval_noadj;
    o.converged = converged ? 1 : 0;
    return true;
}

}  // namespace

extern "C" {

// dosage: [n_var][n] doubles (NaN = missing), overwritten (imputation / allele flip) like the reference's buffer.
// out: [n_var][8] = AF, mac, num, beta, SE, pval, pval_noadj, converged; valid[n_var] = passed the filters.
int orc_score_test(int trait, long n, int K, const double *tau, const double *y, const double *mu, const double *y_mu,
                   const double *mu2, const double *t_XXVX_inv, const double *XV, const double *t_XVX_inv_XV,
                   const double *XVX, const double *t_X, const double *S_a, double varRatio, double thr_maf, double thr_mac,
                   double thr_missing, double thr_pval_spa, long n_var, double *dosage, double *out, int *valid) {
    Model M{trait, (size_t)n, K, tau, y, mu, y_mu, mu2, t_XXVX_inv, XV, t_XVX_inv_XV, XVX, t_X, S_a,
            varRatio, thr_maf, thr_mac, thr_missing, thr_pval_spa};
    // saige_score_test_init (saige_main.cpp:101-108): non-finite thresholds switch the filter off
    if (!std::isfinite(M.thr_maf)) M.thr_maf = -1;
    if (!std::isfinite(M.thr_mac)) M.thr_mac = -1;
    if (!std::isfinite(M.thr_missing)) M.thr_missing = 1;
    if (!std::isfinite(M.thr_pval_spa)) M.thr_pval_spa = 0.05;
#pragma omp parallel
    {
        std::vector<int> buf_index(n);
        std::vector<double> buf_coeff(K), buf_adj_g(n), buf_B(n), buf_g_tilde(n), buf_X1(K), buf_spa(2 * n);
#pragma omp for schedule(dynamic, 16)
        for (long v = 0; v < n_var; v++) {
            Out o{};
            const bool ok = single_test(M, dosage + (size_t)v * n, o, buf_index, buf_coeff, buf_adj_g, buf_B, buf_g_tilde, buf_X1, buf_spa);
            valid[v] = ok ? 1 : 0;
            double *r = out + (size_t)v * 8;
            if (ok) {
                r[0] = o.AF; r[1] = o.mac; r[2] = o.num; r[3] = o.beta; r[4] = o.SE; r[5] = o.pval; r[6] = o.pval_noadj; r[7] = o.converged;
            } else {
                for (int k = 0; k < 8; k++) r[k] = kNaN;
            }
        }
    }
    return 0;
}

double orc_qnorm(double p) { return qnorm_as241(p); }

// Saddle_Prob (SPATest.cpp:232-296), used by saige_GxG_snp_bin
double orc_saddle_prob(double q, double m1, double var1, long n, const double *mu, const double *g, double cutoff,
                       int *converged, double *p_noadj) {
    bool conv = false;
    const double p = Saddle_Prob(q, m1, var1, (size_t)n, mu, g, cutoff, conv, p_noadj);
    *converged = conv ? 1 : 0;
    return p;
}

}  // extern "C"
