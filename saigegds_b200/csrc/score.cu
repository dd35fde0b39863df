// Single-variant score test with saddle-point approximation on the GPU: the per-variant loop of
// seqAssocGLMM_SPA.  Replaces saige_score_test_init / saige_score_test_bin / saige_score_test_quant
// (src/saige_main.cpp:101-155, 188-407) and the SPA routines of src/SPATest.cpp; the arithmetic is in score_body.h.
//
// Two kernels.
//  * score_tiled_kernel: the score statistic of every variant.  A block owns 16 variants (8 warps x 2) and walks the
//    samples in tiles of 256; the model values of a tile -- 2K + 3 doubles per sample: the row of (X'VX)^-1 X'V, the row of
//    WX, y - mu, w, mu -- are staged once per block in shared memory (cp.async, double buffered, stored at init in the
//    tile-major order the lanes read conflict-free), so the 79 MB of model data (n = 430K, K = 10) leave L2 once per 16
//    variants instead of once per variant; a lane keeps the 2K + 3 running sums of its two variants in registers.
//    Variants whose normal-approximation p-value calls for the saddle-point step are appended to a list.
//  * score_test_kernel: one block of 256 threads per listed variant (atomic work counter; a saddle-point variant costs
//    a pass over all samples plus tens of passes over its non-zero genotypes), the whole test of score_body.h.
// SGB_SCORE_PER_VARIANT (sgb_score_test_set_path) runs every variant through the second kernel alone.
#include <algorithm>
#include <cstdlib>

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
    // tiled kernel: model values in tile order [tile][row][32 * spl], rows = a(K), w*x(K), y-mu, w, mu
    DevBuf<double> mt;
    int rows = 0, spl = 8;
    int path = SGB_SCORE_TILED;
    DevBuf<int32_t> spa_list;
    DevBuf<unsigned int> spa_count;
    PinBuf<unsigned int> h_count;
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
__global__ void __launch_bounds__(kThreads, 2) score_test_kernel(score::Model M, Src src, int64_t n_var,
                                                              const int32_t *__restrict__ list, double *spa,
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
        if ((int64_t)next >= n_var) break;
        const int64_t v = list ? (int64_t)list[next] : (int64_t)next;   // n_var counts list entries when a list is given
        const bool ok = score::test_variant<KMAX>(env, M, src.row(v), spa_g, spa_mu, out + v * score::kOutCols);
        if (threadIdx.x == 0) valid[v] = ok ? 1 : 0;
    }
}

// ---- tiled kernel ----------------------------------------------------------------------------------------------------
constexpr int kTileWarps = 16;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;   // fixed butterfly: every lane holds the same bits
}

// A lane's genotypes of one tile: SPL consecutive samples starting at tile * 32 SPL + lane * SPL.
// tab[c] = value of code c after mean imputation and the flip to the minor allele (tab[3]: missing); all zero for a
// filtered variant, which then adds nothing.
template <int SPL>
struct PackedFrag {
    uint32_t bits;
    __device__ __forceinline__ double value(int k) const {
        const unsigned c = (bits >> (2 * k)) & 3u;
        return c < 3 ? (double)c : score::nan_value();
    }
    __device__ __forceinline__ double coded(int k, const double (&tab)[4], bool, bool) const {
        const unsigned c = (bits >> (2 * k)) & 3u;
        return c == 0 ? tab[0] : (c == 1 ? tab[1] : (c == 2 ? tab[2] : tab[3]));
    }
};
struct PackedTiles {
    const uint8_t *base;
    size_t pitch;
    int64_t nb;
    template <int SPL>
    __device__ __forceinline__ PackedFrag<SPL> fetch(int64_t v, int64_t tile, int lane) const {
        const uint8_t *row = base + (size_t)v * pitch;
        const int64_t b0 = tile * (8 * SPL) + lane * (SPL / 4);
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < SPL / 4; j++)
            if (b0 + j < nb) bits |= (uint32_t)row[b0 + j] << (8 * j);
        return PackedFrag<SPL>{bits};
    }
};
template <int SPL>
struct DosageFrag {
    double d[SPL];
    __device__ __forceinline__ double value(int k) const { return isfinite(d[k]) ? d[k] : score::nan_value(); }
    __device__ __forceinline__ double coded(int k, const double (&tab)[4], bool ok, bool minus) const {
        if (!ok) return 0.0;
        if (!isfinite(d[k])) return tab[3];
        return minus ? 2 - d[k] : d[k];
    }
};
struct DosageTiles {
    const double *base;
    int64_t n;
    template <int SPL>
    __device__ __forceinline__ DosageFrag<SPL> fetch(int64_t v, int64_t tile, int lane) const {
        const double *row = base + (size_t)v * n;
        const int64_t i0 = (tile * 32 + lane) * SPL;
        DosageFrag<SPL> f;
#pragma unroll
        for (int k = 0; k < SPL; k++) f.d[k] = (i0 + k < n) ? row[i0 + k] : 0.0;
        return f;
    }
};

// EXACT: the model has exactly KMAX columns (loops and row offsets fold at compile time); otherwise KMAX bounds M.K.
template <int KMAX, bool EXACT, int R, int SPL, class Src>
__global__ void __launch_bounds__(kTileWarps * 32)
score_tiled_kernel(score::Model M, Src src, int64_t n_var, const double *__restrict__ mt, int rows, double *__restrict__ out,
                   int32_t *__restrict__ valid, int32_t *__restrict__ spa_list, unsigned int *__restrict__ spa_count) {
    constexpr int T = 32 * SPL;
    extern __shared__ __align__(16) double tile_smem[];   // [2][rows][T]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = EXACT ? KMAX : M.K;
    const int64_t n = M.n, tiles = (n + T - 1) / T;
    const int64_t v0 = ((int64_t)blockIdx.x * kTileWarps + warp) * R;
    const int tile_doubles = rows * T;

    // ---- allele counts of this warp's variants (f64_af_ac_impute), filters, coded allele
    bool ok[R], minus[R];
    double AF[R], AC[R], mac[R], imputed[R];
    int Num[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        ok[r] = false; minus[r] = false; AF[r] = AC[r] = mac[r] = imputed[r] = 0; Num[r] = 0;
        const int64_t v = v0 + r;
        if (v < n_var) {
            double s = 0;
            int cnt = 0;
            for (int64_t t = 0; t < tiles; t++) {
                const auto f = src.template fetch<SPL>(v, t, lane);
#pragma unroll
                for (int k = 0; k < SPL; k++) {
                    const double x = f.value(k);
                    if ((t * 32 + lane) * SPL + k < n && !isnan(x)) { s += x; cnt++; }
                }
            }
            AC[r] = warp_sum(s);
            Num[r] = (int)warp_sum((double)cnt);
            ok[r] = score::variant_passes(M, AC[r], Num[r], AF[r], mac[r]);
            minus[r] = AF[r] > 0.5;
            imputed[r] = AF[r] * 2;
        }
    }

    // ---- sweep over the sample tiles
    double coef[R][KMAX], xwg[R][KMAX], SyG[R], SwGG[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        SyG[r] = SwGG[r] = 0;
#pragma unroll
        for (int c = 0; c < KMAX; c++) coef[r][c] = xwg[r][c] = 0;
    }
    auto stage = [&](int64_t t) {
        double *dst = tile_smem + (size_t)(t & 1) * tile_doubles;
        const double *srcp = mt + (size_t)t * tile_doubles;
        for (int i = threadIdx.x; i < tile_doubles / 2; i += kTileWarps * 32) cp_async16(dst + 2 * i, srcp + 2 * i);
    };
    double tab[R][4];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const double miss = ok[r] ? imputed[r] : 0.0, one = ok[r] ? 1.0 : 0.0;
        tab[r][0] = minus[r] ? 2 * one : 0.0;
        tab[r][1] = one;
        tab[r][2] = minus[r] ? 0.0 : 2 * one;
        tab[r][3] = minus[r] ? 2 * one - miss : miss;
    }
    using Frag = decltype(src.template fetch<SPL>(0, 0, 0));
    Frag cur[R], nxt[R];
#pragma unroll
    for (int r = 0; r < R; r++) cur[r] = nxt[r] = src.template fetch<SPL>(ok[r] ? v0 + r : 0, 0, lane);
    stage(0);
    cp_async_commit();
    for (int64_t t = 0; t < tiles; t++) {
        if (t + 1 < tiles) {
            stage(t + 1);
            // the genotypes of the next tile travel while this one is being used
#pragma unroll
            for (int r = 0; r < R; r++) nxt[r] = src.template fetch<SPL>(ok[r] ? v0 + r : 0, t + 1, lane);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double *m = tile_smem + (size_t)(t & 1) * tile_doubles + lane;
#pragma unroll
        for (int k = 0; k < SPL; k++) {
            double g[R];
            bool any = false;
#pragma unroll
            for (int r = 0; r < R; r++) {
                g[r] = cur[r].coded(k, tab[r], ok[r], minus[r]);
                any |= (g[r] != 0);
            }
            if (any) {   // samples past n hold zeros in every model row
                const double *mk = m + k * 32;
#pragma unroll
                for (int c = 0; c < KMAX; c++)
                    if (c < K) {
                        const double a = mk[c * T], xw = mk[(K + c) * T];
#pragma unroll
                        for (int r = 0; r < R; r++) { coef[r][c] += g[r] * a; xwg[r][c] += g[r] * xw; }
                    }
                const double ym = mk[2 * K * T], w = mk[(2 * K + 1) * T];
#pragma unroll
                for (int r = 0; r < R; r++) { SyG[r] += g[r] * ym; SwGG[r] += g[r] * g[r] * w; }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) cur[r] = nxt[r];
        __syncthreads();   // this buffer is refilled by the copy issued in the next iteration
    }
    cp_async_wait<0>();

    // ---- per-variant statistics; every lane holds the same sums, lane 0 writes
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int64_t v = v0 + r;
        if (v >= n_var) continue;
        double *o = out + v * score::kOutCols;
        if (!ok[r]) {
            if (lane == 0) {
                for (int k = 0; k < score::kOutCols; k++) o[k] = score::nan_value();
                valid[v] = 0;
            }
            continue;
        }
#pragma unroll
        for (int c = 0; c < KMAX; c++)
            if (c < K) { coef[r][c] = warp_sum(coef[r][c]); xwg[r][c] = warp_sum(xwg[r][c]); }
        const double sy = warp_sum(SyG[r]), sw = warp_sum(SwGG[r]);
        double S, var2, coef_xmu, pval_noadj, beta;
        score::score_stats<KMAX>(M, coef[r], xwg[r], sy, sw, mac[r], S, var2, coef_xmu, pval_noadj, beta);
        if (lane == 0) {
            const bool fin = isfinite(pval_noadj);
            if (minus[r]) beta = -beta;
            o[0] = AF[r]; o[1] = mac[r]; o[2] = (double)Num[r]; o[3] = beta;
            o[4] = fabs(beta / score::qnorm_as241(pval_noadj / 2));
            o[5] = pval_noadj; o[6] = pval_noadj; o[7] = fin ? 1.0 : 0.0;
            valid[v] = 1;
            // saige_main.cpp:353-355: binary trait and a small enough p-value -> saddle-point approximation (second kernel)
            if (M.trait == 0 && fin && pval_noadj <= M.thr_pval_spa) spa_list[atomicAdd(spa_count, 1u)] = (int32_t)v;
        }
    }
}

template <class Src>
void launch_per_variant(Context &c, ScoreState &s, const Src &src, int64_t n_items, const int32_t *list) {
    const int grid = (int)std::min<int64_t>(n_items, s.grid);
    SGB_CUDA(cudaMemsetAsync(s.counter.get(), 0, sizeof(unsigned long long), c.stream));
    c.prof_begin();
    const int K = s.M.K;
#define SGB_SCORE_LAUNCH(KMAX)                                                                                         \
    score_test_kernel<KMAX, Src><<<grid, kThreads, 0, c.stream>>>(s.M, src, n_items, list, s.spa.get(), s.counter.get(), \
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

template <int KMAX, bool EXACT, int R, int SPL, class Tiles>
void launch_tiled_kernel(Context &c, ScoreState &s, const Tiles &tiles, int64_t n_var) {
    auto kern = score_tiled_kernel<KMAX, EXACT, R, SPL, Tiles>;
    const size_t smem = (size_t)2 * s.rows * 32 * SPL * sizeof(double);
    SGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t per_block = (int64_t)kTileWarps * R;
    kern<<<(unsigned)((n_var + per_block - 1) / per_block), kTileWarps * 32, smem, c.stream>>>(
        s.M, tiles, n_var, s.mt.get(), s.rows, s.out.get(), s.valid.get(), s.spa_list.get(), s.spa_count.get());
    SGB_CHECK_LAUNCH();
}

// All variants through the tiled kernel, then the listed ones through the per-variant kernel (Src = the same genotypes).
template <class Tiles, class Src>
void launch(Context &c, ScoreState &s, const Tiles &tiles, const Src &src, int64_t n_var) {
    if (s.path == SGB_SCORE_PER_VARIANT) {
        launch_per_variant(c, s, src, n_var, nullptr);
        return;
    }
    if (n_var > 0x7fffffff) throw Error(SGB_ERR_INVALID, "more than 2^31 - 1 variants in one batch");
    s.spa_list.ensure((size_t)n_var);
    SGB_CUDA(cudaMemsetAsync(s.spa_count.get(), 0, sizeof(unsigned int), c.stream));
    c.prof_begin();
    switch (s.M.K) {   // one variant per warp, 16 warps: the 2K + 2 running sums must leave room for 512 threads per SM
#define SGB_TILED_EXACT(KX) case KX: launch_tiled_kernel<KX, true, 1, 8>(c, s, tiles, n_var); break
        SGB_TILED_EXACT(1); SGB_TILED_EXACT(2); SGB_TILED_EXACT(3); SGB_TILED_EXACT(4);
        SGB_TILED_EXACT(5); SGB_TILED_EXACT(6); SGB_TILED_EXACT(7); SGB_TILED_EXACT(8);
        SGB_TILED_EXACT(9); SGB_TILED_EXACT(10); SGB_TILED_EXACT(11); SGB_TILED_EXACT(12);
        SGB_TILED_EXACT(13); SGB_TILED_EXACT(14); SGB_TILED_EXACT(15); SGB_TILED_EXACT(16);
#undef SGB_TILED_EXACT
        default: launch_tiled_kernel<32, false, 1, 4>(c, s, tiles, n_var); break;
    }
    c.prof_end("score_tiled_kernel");
    c.stats.n_kernel_launches++;
    c.d2h(s.h_count.p, s.spa_count.get(), sizeof(unsigned int));
    c.sync();
    const int64_t n_spa = *s.h_count.p;
    if (n_spa > 0) launch_per_variant(c, s, src, n_spa, s.spa_list.get());
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

// Saddle_Prob (SPATest.cpp:232-296) on device vectors g, mu of length n: one block; result[3] = p-value, normal
// approximation, converged.
namespace {
constexpr int kSaddleThreads = 1024;
__global__ void __launch_bounds__(kSaddleThreads) saddle_dense_kernel(const double *g, const double *mu, int64_t n, double q,
                                                                      double m1, double var1, double cutoff, double *result) {
    __shared__ double red[kSaddleThreads / 32];
    __shared__ int wsum[kSaddleThreads / 32];
    BlockEnv env{red, wsum};
    double gp = 0, gn = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double v = g[i];
        if (v > 0) gp += v; else gn += v;
    }
    gp = env.sum(gp);
    gn = env.sum(gn);
    bool converged;
    double p_noadj;
    const double pval = score::saddle_prob(env, q, m1, var1, gp, gn, n, g, mu, 0.0, 0.0, cutoff, converged, p_noadj);
    if (threadIdx.x == 0) {
        result[0] = pval;
        result[1] = p_noadj;
        result[2] = converged ? 1.0 : 0.0;
    }
}
}  // namespace

void saddle_prob_dense(Context &c, const double *g_device, const double *mu_device, int64_t n, double q, double m1, double var1,
                       double cutoff, double *pval, double *p_noadj, bool *converged) {
    c.red_out.ensure(256);
    c.prof_begin();
    saddle_dense_kernel<<<1, kSaddleThreads, 0, c.stream>>>(g_device, mu_device, n, q, m1, var1, cutoff, c.red_out.get());
    SGB_CHECK_LAUNCH();
    c.prof_end("saddle_dense_kernel");
    c.stats.n_kernel_launches++;
    double h[3];
    c.d2h(h, c.red_out.get(), sizeof(h));
    c.sync();
    *pval = h[0];
    *p_noadj = h[1];
    *converged = h[2] != 0;
}

double qnorm_host(double p) { return score::qnorm_as241(p); }

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
    // model values in the order the tiled kernel reads them: [tile][row][T], sample s of a tile at (s % spl) * 32 + s / spl
    {
        const bool bin = (m->trait == 0);
        s->spl = (K <= 16) ? 8 : 4;
        s->rows = (int)(2 * K + 3);
        const size_t T = (size_t)32 * s->spl, tiles = (n + T - 1) / T, R = (size_t)s->rows;
        std::vector<double> mt(tiles * R * T, 0.0);
        for (size_t i = 0; i < n; i++) {
            const size_t sidx = i % T, pos = (sidx % s->spl) * 32 + sidx / s->spl;
            double *b = mt.data() + (i / T) * R * T + pos;
            const double w = bin ? m->mu2[i] : 1.0;
            for (size_t k = 0; k < K; k++) {
                b[k * T] = m->t_XVX_inv_XV[i * K + k];
                b[(K + k) * T] = w * m->t_X[i * K + k];
            }
            b[2 * K * T] = m->y_mu[i];
            b[(2 * K + 1) * T] = w;
            b[(2 * K + 2) * T] = m->mu[i];
        }
        up(s->mt, mt.data(), mt.size());
        c.sync();
    }
    s->spa_count.ensure(1);
    s->h_count.ensure(1);
    const char *env_path = getenv("SGB_SCORE_PATH");
    s->path = (env_path && std::string(env_path) == "per_variant") ? SGB_SCORE_PER_VARIANT : SGB_SCORE_TILED;
    s->grid = c.sm_count * 4;
    // saddle-point scratch (2 n doubles per block of the candidate kernel): binary traits only -- a quantitative trait never
    // takes the saddle-point branch (saige_main.cpp:322-350)
    if (m->trait == 0) s->spa.ensure((size_t)s->grid * 2 * n);
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

void score_set_path(Context &c, int path) {
    if (!c.score) throw Error(SGB_ERR_STATE, "no model: call sgb_score_test_init first");
    if (path != SGB_SCORE_TILED && path != SGB_SCORE_PER_VARIANT) throw Error(SGB_ERR_INVALID, "unknown score-test path");
    c.score->path = path;
}

void score_test_packed(Context &c, const uint8_t *packed, int64_t nb, int64_t n_var, double *out, int32_t *valid) {
    ScoreState &s = state(c, n_var);
    if (!packed) throw Error(SGB_ERR_INVALID, "packed is NULL");
    if (nb != (s.M.n + 3) / 4) throw Error(SGB_ERR_INVALID, "n_bytes_per_variant must equal ceil(n_samp/4) of the model");
    s.geno.ensure((size_t)nb * n_var);
    c.h2d(s.geno.get(), packed, (size_t)nb * n_var);
    launch(c, s, PackedTiles{s.geno.get(), (size_t)nb, nb}, PackedSrc{s.geno.get(), (size_t)nb}, n_var);
    fetch(c, s, n_var, out, valid);
}

void score_test_dosage(Context &c, const double *dosage, int64_t n_var, double *out, int32_t *valid) {
    ScoreState &s = state(c, n_var);
    if (!dosage) throw Error(SGB_ERR_INVALID, "dosage is NULL");
    const size_t bytes = sizeof(double) * (size_t)s.M.n * n_var;
    s.geno.ensure(bytes);
    c.h2d(s.geno.get(), dosage, bytes);
    launch(c, s, DosageTiles{(const double *)s.geno.get(), s.M.n}, DosageSrc{(const double *)s.geno.get(), (size_t)s.M.n}, n_var);
    fetch(c, s, n_var, out, valid);
}

void score_test_stored(Context &c, int64_t first, int64_t n_var, double *out, int32_t *valid, float *kernel_ms) {
    ScoreState &s = state(c, n_var);
    c.require_stored();
    if (c.N != s.M.n) throw Error(SGB_ERR_INVALID, "the stored genotypes and the model differ in the number of samples");
    if (first < 0 || first + n_var > c.M) throw Error(SGB_ERR_INVALID, "variant range outside the stored shard");
    SGB_CUDA(cudaEventRecord(c.ev0, c.stream));
    launch(c, s, PackedTiles{c.packed.get() + (size_t)first * c.pitch, c.pitch, c.NB},
           PackedSrc{c.packed.get() + (size_t)first * c.pitch, c.pitch}, n_var);
    SGB_CUDA(cudaEventRecord(c.ev1, c.stream));
    fetch(c, s, n_var, out, valid);
    if (kernel_ms) SGB_CUDA(cudaEventElapsedTime(kernel_ms, c.ev0, c.ev1));
}

}  // namespace sgb
