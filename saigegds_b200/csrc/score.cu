// Single-variant score test with saddle-point approximation on the GPU: the per-variant loop of
// seqAssocGLMM_SPA.  Replaces saige_score_test_init / saige_score_test_bin / saige_score_test_quant
// (src/saige_main.cpp:101-155, 188-407) and the SPA routines of src/SPATest.cpp; the arithmetic is in score_body.h.
//
// One block of 256 threads per variant, variants handed out through an atomic counter (a saddle-point variant costs many
// passes over its non-zero genotypes, a plain one three), blocks persistent.  The model vectors (y - mu, mu, mu(1-mu),
// the two n x K matrices; 8 n (3 + 2K) bytes, 79 MB at n = 430K, K = 10) are read by every variant and stay in L2; the
// genotypes stream through once: 2 bits per sample from a packed batch, or 8 bytes per sample for real-valued dosages.
#include <algorithm>

#include "ctx.h"
#include "score_body.h"

namespace sgb {

struct ScoreState {
    score::Model M{};
    DevBuf<double> y_mu, mu, mu2, t_XVX_inv_XV, t_X, XVX, S_a, X_mu;
    DevBuf<double> spa;                  // [grid][2][n] compacted (g, mu) pairs of the saddle-point step
    DevBuf<unsigned long long> counter;  // next variant
    DevBuf<double> out;                  // [n_var][8]
    DevBuf<int32_t> valid;
    DevBuf<uint8_t> geno;                // staged batch (packed bytes or dosages)
    int grid = 0;
};

namespace {

constexpr int kThreads = 256;

struct BlockEnv {
    double *red;   // shared, one slot per warp
    int *wsum;     // shared, one slot per warp
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nthr() const { return blockDim.x; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
    // Sum over the block; every thread receives the same bits (fixed butterfly inside a warp, warps added in order).
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        __syncthreads();   // readers of the previous sum are done with `red`
        if ((threadIdx.x & 31) == 0) red[w] = v;
        __syncthreads();
        double t = 0;
        for (int k = 0; k < nw; k++) t += red[k];
        return t;
    }
    // Exclusive scan of one int per thread, in thread order; total to every thread.
    __device__ __forceinline__ int64_t excl_scan(int v, int64_t &total) {
        const int l = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += t;
        }
        __syncthreads();
        if (l == 31) wsum[w] = inc;
        __syncthreads();
        int64_t base = 0, tot = 0;
        for (int k = 0; k < nw; k++) {
            if (k < w) base += wsum[k];
            tot += wsum[k];
        }
        total = tot;
        return base + inc - v;
    }
};

struct PackedSrc {
    const uint8_t *base;
    size_t pitch;
    __device__ __forceinline__ score::PackedRow row(int64_t v) const { return score::PackedRow{base + (size_t)v * pitch}; }
};
struct DosageSrc {
    const double *base;
    size_t n;
    __device__ __forceinline__ score::DosageRow row(int64_t v) const { return score::DosageRow{base + (size_t)v * n}; }
};

template <int KMAX, class Src>
__global__ void __launch_bounds__(kThreads) score_test_kernel(score::Model M, Src src, int64_t n_var, double *spa,
                                                              unsigned long long *__restrict__ counter,
                                                              double *__restrict__ out, int32_t *__restrict__ valid) {
    __shared__ double red[kThreads / 32];
    __shared__ int wsum[kThreads / 32];
    __shared__ unsigned long long next;
    BlockEnv env{red, wsum};
    double *spa_g = spa + (size_t)blockIdx.x * 2 * (size_t)M.n, *spa_mu = spa_g + M.n;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) next = atomicAdd(counter, 1ULL);
        __syncthreads();
        const int64_t v = (int64_t)next;
        if (v >= n_var) break;
        const bool ok = score::test_variant<KMAX>(env, M, src.row(v), spa_g, spa_mu, out + v * score::kOutCols);
        if (threadIdx.x == 0) valid[v] = ok ? 1 : 0;
    }
}

template <class Src>
void launch(Context &c, ScoreState &s, const Src &src, int64_t n_var) {
    const int grid = (int)std::min<int64_t>(n_var, s.grid);
    SGB_CUDA(cudaMemsetAsync(s.counter.get(), 0, sizeof(unsigned long long), c.stream));
    c.prof_begin();
    const int K = s.M.K;
#define SGB_SCORE_LAUNCH(KMAX)                                                                                         \
    score_test_kernel<KMAX, Src><<<grid, kThreads, 0, c.stream>>>(s.M, src, n_var, s.spa.get(), s.counter.get(),       \
                                                                  s.out.get(), s.valid.get())
    if (K <= 4) SGB_SCORE_LAUNCH(4);
    else if (K <= 8) SGB_SCORE_LAUNCH(8);
    else if (K <= 16) SGB_SCORE_LAUNCH(16);
    else SGB_SCORE_LAUNCH(32);
#undef SGB_SCORE_LAUNCH
    SGB_CHECK_LAUNCH();
    c.prof_end("score_test_kernel");
    c.stats.n_kernel_launches++;
}

void fetch(Context &c, ScoreState &s, int64_t n_var, double *out, int32_t *valid) {
    if (out) c.d2h(out, s.out.get(), sizeof(double) * score::kOutCols * n_var);
    if (valid) c.d2h(valid, s.valid.get(), sizeof(int32_t) * n_var);
    c.sync();
}

ScoreState &state(Context &c, int64_t n_var) {
    if (!c.score) throw Error(SGB_ERR_STATE, "no model: call sgb_score_test_init first");
    if (n_var < 1) throw Error(SGB_ERR_INVALID, "no variants");
    ScoreState &s = *c.score;
    s.out.ensure((size_t)score::kOutCols * n_var);
    s.valid.ensure((size_t)n_var);
    return s;
}

}  // namespace

void score_release(Context &c) {
    delete c.score;
    c.score = nullptr;
}

void score_init(Context &c, const sgb_score_model *m, double maf, double mac, double missing, double spa_pval) {
    if (!m || !m->tau || !m->mu || !m->y_mu || !m->mu2 || !m->t_XVX_inv_XV || !m->XVX || !m->t_X || !m->S_a)
        throw Error(SGB_ERR_INVALID, "score model: NULL array");
    if (m->n < 1 || m->K < 1) throw Error(SGB_ERR_INVALID, "score model: empty");
    if (m->K > 32) throw Error(SGB_ERR_INVALID, "score model: more than 32 fixed-effect columns are not supported");
    if (m->trait != 0 && m->trait != 1) throw Error(SGB_ERR_INVALID, "score model: trait must be 0 (binary) or 1 (quantitative)");
    score_release(c);
    ScoreState *s = new ScoreState();
    c.score = s;
    const size_t n = (size_t)m->n, K = (size_t)m->K;
    auto up = [&](DevBuf<double> &d, const double *h, size_t count) {
        d.ensure(count);
        c.h2d(d.get(), h, sizeof(double) * count);
    };
    up(s->y_mu, m->y_mu, n); up(s->mu, m->mu, n); up(s->mu2, m->mu2, n);
    up(s->t_XVX_inv_XV, m->t_XVX_inv_XV, n * K); up(s->t_X, m->t_X, n * K);
    up(s->XVX, m->XVX, K * K); up(s->S_a, m->S_a, K);
    std::vector<double> x_mu(K, 0.0);   // colSums(X * mu): gives m1 of the saddle-point step without a pass over all samples
    for (size_t i = 0; i < n; i++)
        for (size_t k = 0; k < K; k++) x_mu[k] += m->t_X[i * K + k] * m->mu[i];
    up(s->X_mu, x_mu.data(), K);
    c.sync();
    s->grid = c.sm_count * 4;
    s->spa.ensure((size_t)s->grid * 2 * n);
    s->counter.ensure(1);
    score::Model &M = s->M;
    M.trait = m->trait; M.n = m->n; M.K = m->K; M.tau0 = m->tau[0];
    M.y_mu = s->y_mu.get(); M.mu = s->mu.get(); M.mu2 = s->mu2.get();
    M.t_XVX_inv_XV = s->t_XVX_inv_XV.get(); M.t_X = s->t_X.get(); M.XVX = s->XVX.get(); M.S_a = s->S_a.get(); M.X_mu = s->X_mu.get();
    M.varRatio = m->var_ratio;
    // saige_score_test_init, saige_main.cpp:106-113: a non-finite threshold switches its filter off
    M.thr_maf = std::isfinite(maf) ? maf : -1;
    M.thr_mac = std::isfinite(mac) ? mac : -1;
    M.thr_missing = std::isfinite(missing) ? missing : 1;
    M.thr_pval_spa = std::isfinite(spa_pval) ? spa_pval : 0.05;
}

void score_test_packed(Context &c, const uint8_t *packed, int64_t nb, int64_t n_var, double *out, int32_t *valid) {
    ScoreState &s = state(c, n_var);
    if (!packed) throw Error(SGB_ERR_INVALID, "packed is NULL");
    if (nb != (s.M.n + 3) / 4) throw Error(SGB_ERR_INVALID, "n_bytes_per_variant must equal ceil(n_samp/4) of the model");
    s.geno.ensure((size_t)nb * n_var);
    c.h2d(s.geno.get(), packed, (size_t)nb * n_var);
    launch(c, s, PackedSrc{s.geno.get(), (size_t)nb}, n_var);
    fetch(c, s, n_var, out, valid);
}

void score_test_dosage(Context &c, const double *dosage, int64_t n_var, double *out, int32_t *valid) {
    ScoreState &s = state(c, n_var);
    if (!dosage) throw Error(SGB_ERR_INVALID, "dosage is NULL");
    const size_t bytes = sizeof(double) * (size_t)s.M.n * n_var;
    s.geno.ensure(bytes);
    c.h2d(s.geno.get(), dosage, bytes);
    launch(c, s, DosageSrc{(const double *)s.geno.get(), (size_t)s.M.n}, n_var);
    fetch(c, s, n_var, out, valid);
}

void score_test_stored(Context &c, int64_t first, int64_t n_var, double *out, int32_t *valid, float *kernel_ms) {
    ScoreState &s = state(c, n_var);
    c.require_stored();
    if (c.N != s.M.n) throw Error(SGB_ERR_INVALID, "the stored genotypes and the model differ in the number of samples");
    if (first < 0 || first + n_var > c.M) throw Error(SGB_ERR_INVALID, "variant range outside the stored shard");
    SGB_CUDA(cudaEventRecord(c.ev0, c.stream));
    launch(c, s, PackedSrc{c.packed.get() + (size_t)first * c.pitch, c.pitch}, n_var);
    SGB_CUDA(cudaEventRecord(c.ev1, c.stream));
    fetch(c, s, n_var, out, valid);
    if (kernel_ms) SGB_CUDA(cudaEventElapsedTime(kernel_ms, c.ev0, c.ev1));
}

}  // namespace sgb
