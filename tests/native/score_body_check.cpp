// TEST HARNESS (not product code): compiles the device body of the score test, saigegds_b200/csrc/score_body.h, for ONE
// host thread so that its arithmetic can be compared with the oracle and the reference's golden p-values on a machine
// without a GPU.  The block primitives collapse to the identity; the CUDA versions of them live in csrc/score.cu and are
// exercised by the -m gpu tests.  Built on the fly by tests/test_score_test.py.
#include <cmath>
#include <cstddef>
#include <vector>

#include "../../saigegds_b200/csrc/score_body.h"

namespace {
struct HostEnv {
    int tid() const { return 0; }
    int nthr() const { return 1; }
    void sync() {}
    double sum(double v) { return v; }
    int64_t excl_scan(int v, int64_t &total) { total = v; return 0; }
};
int g_dual = 0;   // 1: saddle_prob_dual (both Newton iterations advanced by the same passes) instead of saddle_prob
}  // namespace

extern "C" void score_body_set_dual(int on) { g_dual = on; }

extern "C" int score_body_check(int trait, long n, int K, double tau0, const double *mu, const double *y_mu, const double *mu2,
                                const double *t_XVX_inv_XV, const double *XVX, const double *t_X, const double *S_a,
                                double varRatio, double thr_maf, double thr_mac, double thr_missing, double thr_pval_spa,
                                long n_var, const double *dosage, const unsigned char *packed, long nb, double *out, int *valid) {
    using namespace sgb::score;
    std::vector<double> X_mu(K, 0.0), spa(2 * (size_t)n);
    for (long i = 0; i < n; i++)
        for (int c = 0; c < K; c++) X_mu[c] += t_X[(size_t)i * K + c] * mu[i];
    Model M{trait, n, K, tau0, y_mu, mu, mu2, t_XVX_inv_XV, t_X, XVX, S_a, X_mu.data(), varRatio,
            std::isfinite(thr_maf) ? thr_maf : -1, std::isfinite(thr_mac) ? thr_mac : -1,
            std::isfinite(thr_missing) ? thr_missing : 1, std::isfinite(thr_pval_spa) ? thr_pval_spa : 0.05};
    if (K > 32) return 1;
    HostEnv env;
    for (long v = 0; v < n_var; v++) {
        bool ok;
        if (packed && g_dual)
            ok = test_variant<32, HostEnv, PackedRow, true>(env, M, PackedRow{packed + (size_t)v * nb}, spa.data(), spa.data() + n, out + v * kOutCols);
        else if (packed)
            ok = test_variant<32>(env, M, PackedRow{packed + (size_t)v * nb}, spa.data(), spa.data() + n, out + v * kOutCols);
        else if (g_dual)
            ok = test_variant<32, HostEnv, DosageRow, true>(env, M, DosageRow{dosage + (size_t)v * n}, spa.data(), spa.data() + n, out + v * kOutCols);
        else
            ok = test_variant<32>(env, M, DosageRow{dosage + (size_t)v * n}, spa.data(), spa.data() + n, out + v * kOutCols);
        valid[v] = ok ? 1 : 0;
    }
    return 0;
}
