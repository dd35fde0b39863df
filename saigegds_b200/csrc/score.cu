// Single-variant score test with saddle-point approximation on the GPU: the per-variant loop of
// seqAssocGLMM_SPA.  Replaces saige_score_test_init / saige_score_test_bin / saige_score_test_quant
// (src/saige_main.cpp:101-155, 188-407) and the SPA routines of src/SPATest.cpp; the arithmetic is in score_body.h.
//
// Three paths for the score statistics of all variants of a block, then one kernel for the saddle-point candidates.
//  * tensor (default for 2-bit packed genotypes): the statistics are sums of model columns over the samples of each genotype class,
//    so a block of variants is three integer GEMMs on the tcgen05 pair kernel of the batched GRM product (grm_umma.cuh) -- A = low
//    bit / high bit / low & high of the 2-bit codes, B = the digit planes of the 2K + 4 model columns (a, w x, y - mu, w, mu, 1),
//    quantised once per model -- followed by score_finish_kernel (one thread per variant: exact class sums -> allele counts,
//    filters, score_stats).  9,472 variants x 430K samples: 3 x 0.33 ms instead of 32 ms for the tiled kernel.
//  * score_tiled_kernel (dosages, and SGB_SCORE_TILED): the score statistic of every variant.  A block owns 16 variants (8 warps x 2) and walks the
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
    int grid = 0, grid_spa = 0;
    // tiled kernel: model values in tile order [tile][row][32 * spl], rows = a(K), w*x(K), y-mu, w, mu
    DevBuf<double> mt;
    int rows = 0, spl = 8;
    int path = SGB_SCORE_TILED;
    // tensor path: digit planes of the model columns in groups of <= 32, per-column scalars and exact totals, class-sum limbs
    bool tensor_ok = false;
    int ncols = 0;                       // 2K + 4: a (K), w x (K), y - mu, w, mu, 1
    DevBuf<double> cand;                 // [n_var][kCandCols + K]: what the saddle-point kernel needs of a candidate
    DevBuf<double> xt, mup;              // column-major covariates [K][ldx] and mu [ldx], zero-padded to ldx = 4 ceil(n / 4)
    int64_t ldx = 0;
    int64_t cpad = 0;                    // contraction length of a digit row = 4 * pitch of a packed block
    DevBuf<int8_t> cdig;                 // [groups][192][cpad]
    DevBuf<double> cscal;                // [ncols][8]
    DevBuf<long long> ctot;              // [ncols][2]
    DevBuf<unsigned long long> c_lo, c_hi;   // [3 modes][ncols][n_var]
    DevBuf<int> cerr;
    PinBuf<int> h_cerr;
    DevBuf<int32_t> spa_list;
    DevBuf<unsigned int> spa_count;
    PinBuf<unsigned int> h_count;
};

namespace {

constexpr int kThreads = 256;
constexpr int kSpaThreads = 512;   // saddle-point candidates of the tensor scan: one block per candidate
constexpr int kSpaBuckets = 8;     // candidate lists by minor-allele frequency: the kernel takes the costliest (most non-zero genotypes) first
constexpr int kCandCols = 8;       // AC, Num, AF, S, var2, coef_xmu, G'mu, (spare), then coef[K]

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

// ---- tensor path: class sums -> statistics ------------------------------------------------------------------------------------------
// lo / hi: [3][ncols][n_var] limbs of sum_i bit(code(v, i)) * q(col, i) for bit = low, high, low & high; tot: the same over all i < n.
// Class sums in exact integer arithmetic (S1 = L - M, S2 = H - M, S3 = M, S0 = tot - L - H + M), one rounding each.
template <int KMAX>
__global__ void __launch_bounds__(64) score_finish_kernel(score::Model M, int64_t n_var, int ncols, const unsigned long long *__restrict__ lo,
                                                          const unsigned long long *__restrict__ hi, const long long *__restrict__ tot,
                                                          const double *__restrict__ scal, double *__restrict__ out,
                                                          int32_t *__restrict__ valid, int32_t *__restrict__ spa_list,
                                                          unsigned int *__restrict__ spa_count, double *__restrict__ cand) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_var) return;
    const int K = M.K;
    const size_t plane = (size_t)ncols * n_var;
    // S[k] of column c, converted with the column's unit
    auto classes = [&](int c, double (&S)[4]) {
        const size_t o = (size_t)c * n_var + v;
        const long long Ll = (long long)lo[o], Lh = (long long)hi[o];
        const long long Hl = (long long)lo[plane + o], Hh = (long long)hi[plane + o];
        const long long Ml = (long long)lo[2 * plane + o], Mh = (long long)hi[2 * plane + o];
        const double unit = scal[(size_t)c * kClassScal + 2];
        auto val = [&](long long l, long long h) { return unit == 0 ? 0.0 : unit * ((double)l + 16777216.0 * (double)h); };
        S[1] = val(Ll - Ml, Lh - Mh);
        S[2] = val(Hl - Ml, Hh - Mh);
        S[3] = val(Ml, Mh);
        S[0] = val(tot[2 * c] - Ll - Hl + Ml, tot[2 * c + 1] - Lh - Hh + Mh);
    };
    double S[4];
    classes(2 * K + 3, S);                                   // the column of ones: class counts (exact integers)
    const double AC = S[1] + 2 * S[2];
    const int Num = (int)(M.n - (int64_t)S[3]);
    double AF, mac;
    double *o = out + v * score::kOutCols;
    if (!score::variant_passes(M, AC, Num, AF, mac)) {
        for (int k = 0; k < score::kOutCols; k++) o[k] = score::nan_value();
        valid[v] = 0;
        return;
    }
    const bool minus = AF > 0.5;
    const double miss = AF * 2;
    const double tab[4] = {minus ? 2.0 : 0.0, 1.0, minus ? 0.0 : 2.0, minus ? 2.0 - miss : miss};
    auto gsum = [&](int c) {
        double T[4];
        classes(c, T);
        return tab[0] * T[0] + tab[1] * T[1] + tab[2] * T[2] + tab[3] * T[3];
    };
    double coef[KMAX], xwg[KMAX];
#pragma unroll
    for (int c = 0; c < KMAX; c++) {
        coef[c] = xwg[c] = 0;
        if (c < K) { coef[c] = gsum(c); xwg[c] = gsum(K + c); }
    }
    const double sy = gsum(2 * K);
    classes(2 * K + 1, S);
    const double sw = tab[0] * tab[0] * S[0] + tab[1] * tab[1] * S[1] + tab[2] * tab[2] * S[2] + tab[3] * tab[3] * S[3];
    double Sc, var2, coef_xmu, pval_noadj, beta;
    score::score_stats<KMAX>(M, coef, xwg, sy, sw, mac, Sc, var2, coef_xmu, pval_noadj, beta);
    const bool fin = isfinite(pval_noadj);
    if (minus) beta = -beta;
    o[0] = AF; o[1] = mac; o[2] = (double)Num; o[3] = beta;
    o[4] = fabs(beta / score::qnorm_as241(pval_noadj / 2));
    o[5] = pval_noadj; o[6] = pval_noadj; o[7] = fin ? 1.0 : 0.0;
    valid[v] = 1;
    if (M.trait == 0 && fin && pval_noadj <= M.thr_pval_spa) {
        // saddle-point candidate: hand the sums over so that spa_candidate_kernel does not repeat the passes that produced them
        const double maf = fmin(AF, 1 - AF);
        const int bucket = kSpaBuckets - 1 - min(kSpaBuckets - 1, (int)(maf * (2 * kSpaBuckets)));
        spa_list[(size_t)bucket * n_var + atomicAdd(spa_count + bucket, 1u)] = (int32_t)v;
        double *r = cand + (size_t)v * (kCandCols + K);
        r[0] = AC; r[1] = (double)Num; r[2] = AF; r[3] = Sc; r[4] = var2; r[5] = coef_xmu; r[6] = gsum(2 * K + 2);
        for (int c = 0; c < K; c++) r[kCandCols + c] = coef[c];
    }
}

// One block per saddle-point candidate of the tensor scan (atomic work counter).  The sums that called for the saddle-point step come
// from the finishing kernel; what is left of score::test_variant is the pass over all samples that forms the adjusted genotype
// g = (G - x'coef) / sqrt(AC) -- one-sided sums, and the (g, mu) pairs of the samples with G != 0 compacted in thread order -- and
// the root finding on those pairs (score::saddle_prob_dual: both Newton iterations advanced by the same passes).  Same arithmetic
// per sample as score::spa_adjust; the covariates are read from a column-major copy (xt: [K][ldx], 32 bytes per thread and column for
// its four samples) so that ten 32-byte loads per thread are in flight instead of one row.  Fills beta, SE, pval, converged.
template <int KMAX>
__global__ void __launch_bounds__(kSpaThreads, 2) spa_candidate_kernel(score::Model M, PackedSrc src, int64_t n_cand,
                                                                     const int32_t *__restrict__ list, const unsigned int *__restrict__ counts,
                                                                     int64_t list_stride, const double *__restrict__ cand,
                                                                     const double *__restrict__ xt, const double *__restrict__ mup,
                                                                     int64_t ldx, double *spa, unsigned long long *__restrict__ counter,
                                                                     double *__restrict__ out) {
    __shared__ double red[kSpaThreads / 32];
    __shared__ int wsum[kSpaThreads / 32];
    __shared__ unsigned long long next;
    BlockEnv env{red, wsum};
    const int K = M.K;
    const int64_t n = M.n, ngrp = (n + 3) >> 2;
    double *spa_g = spa + (size_t)blockIdx.x * 2 * (size_t)n, *spa_mu = spa_g + n;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) next = atomicAdd(counter, 1ULL);
        __syncthreads();
        if ((int64_t)next >= n_cand) break;
        int64_t pos = (int64_t)next;
        int bucket = 0;
        while (bucket < kSpaBuckets - 1 && pos >= (int64_t)counts[bucket]) pos -= counts[bucket++];
        const int64_t v = list[(size_t)bucket * list_stride + pos];
        const double *r = cand + (size_t)v * (kCandCols + K);
        const double AC = r[0], AF = r[2], S = r[3], var2 = r[4], coef_xmu = r[5], gmu = r[6];
        const int Num = (int)r[1];
        const bool minus = AF > 0.5;
        double coef[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; c++) coef[c] = (c < K) ? r[kCandCols + c] : 0.0;
        // value of a code after mean imputation and the flip to the minor allele (score::Coded)
        const double imp = AF * 2;
        const double t0 = minus ? 2.0 : 0.0, t2 = minus ? 0.0 : 2.0, t3 = minus ? 2 - imp : imp;
        auto value = [&](unsigned code) { return code == 0 ? t0 : (code == 1 ? 1.0 : (code == 2 ? t2 : t3)); };
        const uint8_t *row = src.base + (size_t)v * src.pitch;
        int my_nnz = 0;
#pragma unroll 4
        for (int64_t g = threadIdx.x; g < ngrp; g += kSpaThreads) {
            const unsigned b = row[g];
            const int lim = (int)min((int64_t)4, n - 4 * g);
#pragma unroll
            for (int j = 0; j < 4; j++) my_nnz += (j < lim && value((b >> (2 * j)) & 3u) != 0);
        }
        double *o = out + v * score::kOutCols;
        const double pval_noadj = o[6];
        // ---- score::spa_adjust, with the pass over the samples restated for 32-byte loads
        const double AC2 = minus ? (2 * Num - AC) : AC;
        const double sc = 1 / sqrt(AC2);
        const double m1 = (gmu - coef_xmu) * sc;
        const double svar2 = var2 * sc * sc, svar1 = svar2 * M.varRatio;
        const double Tstat = S * sc;
        const double q = Tstat / sqrt(svar1) * sqrt(svar2) + m1;
        int64_t nnz = 0;
        int64_t at = env.excl_scan(my_nnz, nnz);
        double g_pos = 0, g_neg = 0, sub_mu = 0, sub_sigma = 0;
        unsigned bnext = (threadIdx.x < ngrp) ? row[threadIdx.x] : 0u;
        for (int64_t g = threadIdx.x; g < ngrp; g += kSpaThreads) {
            // the packed byte of the next iteration is requested before this one's covariates: the warp issues in order, and the
            // first use of the byte would otherwise hold back the loads behind it for a full memory latency
            const unsigned b = bnext;
            if (g + kSpaThreads < ngrp) bnext = row[g + kSpaThreads];
            const int64_t i0 = 4 * g;
            const int lim = (int)min((int64_t)4, n - i0);
            double B[4] = {0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < KMAX; c++)
                if (c < K) {
                    const double2 xa = __ldg(reinterpret_cast<const double2 *>(xt + (size_t)c * ldx + i0));
                    const double2 xb = __ldg(reinterpret_cast<const double2 *>(xt + (size_t)c * ldx + i0 + 2));
                    B[0] += coef[c] * xa.x; B[1] += coef[c] * xa.y; B[2] += coef[c] * xb.x; B[3] += coef[c] * xb.y;
                }
            const double2 ma = __ldg(reinterpret_cast<const double2 *>(mup + i0)), mb = __ldg(reinterpret_cast<const double2 *>(mup + i0 + 2));
            const double mm[4] = {ma.x, ma.y, mb.x, mb.y};
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (j < lim) {
                    const double gv = value((b >> (2 * j)) & 3u);
                    const double gg = (gv - B[j]) * sc;
                    if (gg > 0) g_pos += gg; else g_neg += gg;
                    if (gv != 0) {
                        const double m = mm[j];
                        spa_g[at] = gg;
                        spa_mu[at] = m;
                        at++;
                        sub_mu += gg * m;
                        sub_sigma += gg * gg * m * (1 - m);
                    }
                }
        }
        g_pos = env.sum(g_pos);
        g_neg = env.sum(g_neg);
        const double NAmu = m1 - env.sum(sub_mu);
        const double NAsigma = svar2 - env.sum(sub_sigma);
        env.sync();
        double p_na;
        bool converged = true;
        double pval = score::saddle_prob_dual(env, q, m1, svar2, g_pos, g_neg, nnz, spa_g, spa_mu, NAmu, NAsigma, 2.0, converged, p_na);
        if (pval == 0 && pval_noadj > 0) {
            pval = pval_noadj;
            converged = false;
        }
        double beta = (Tstat / svar1) / sqrt(AC2);
        if (minus) beta = -beta;
        if (threadIdx.x == 0) {
            o[3] = beta;
            o[4] = fabs(beta / score::qnorm_as241(pval / 2));
            o[5] = pval;
            o[7] = converged ? 1.0 : 0.0;
        }
    }
}

// statistics of n_var variants of a packed block (128-byte aligned base and pitch) through the tensor path
void launch_tensor(Context &c, ScoreState &s, const uint8_t *packed, size_t pitch, int64_t n_var) {
    const size_t plane = (size_t)s.ncols * n_var;
    s.c_lo.ensure(3 * plane); s.c_hi.ensure(3 * plane);
    c.prof_begin();
    for (int mode = 0; mode < 3; mode++)
        for (int g = 0, c0 = 0; c0 < s.ncols; g++, c0 += kClassMaxCols) {
            const int nc = std::min(kClassMaxCols, s.ncols - c0);
            umma_class_sums(c, packed, pitch, n_var, s.M.n, s.cdig.get() + (size_t)g * kClassDigitRows * s.cpad, s.cpad, nc, mode + 1,
                            s.c_lo.get() + mode * plane + (size_t)c0 * n_var, s.c_hi.get() + mode * plane + (size_t)c0 * n_var, s.cerr.get());
        }
    c.prof_end("umma_pair_kernel (class sums, 3 bit planes)");
    c.prof_begin();
    const unsigned grid = (unsigned)((n_var + 63) / 64);
    const int K = s.M.K;
    s.cand.ensure((size_t)n_var * (kCandCols + K));
#define SGB_FINISH(KMAX)                                                                                                              \
    score_finish_kernel<KMAX><<<grid, 64, 0, c.stream>>>(s.M, n_var, s.ncols, s.c_lo.get(), s.c_hi.get(), s.ctot.get(), s.cscal.get(), \
                                                         s.out.get(), s.valid.get(), s.spa_list.get(), s.spa_count.get(), s.cand.get())
    if (K <= 4) SGB_FINISH(4);
    else if (K <= 8) SGB_FINISH(8);
    else if (K <= 16) SGB_FINISH(16);
    else SGB_FINISH(32);
#undef SGB_FINISH
    SGB_CHECK_LAUNCH();
    c.prof_end("score_finish_kernel");
    c.stats.n_kernel_launches++;
    SGB_CUDA(cudaMemcpyAsync(s.h_cerr.p, s.cerr.get(), sizeof(int), cudaMemcpyDeviceToHost, c.stream));
}

void launch_candidates(Context &c, ScoreState &s, const PackedSrc &src, int64_t n_cand, int64_t n_var) {
    const int grid = (int)std::min<int64_t>(n_cand, s.grid_spa);
    s.spa.ensure((size_t)grid * 2 * s.M.n);                          // saddle-point scratch of the blocks of this launch
    SGB_CUDA(cudaMemsetAsync(s.counter.get(), 0, sizeof(unsigned long long), c.stream));
    c.prof_begin();
    const int K = s.M.K;
#define SGB_SPA_LAUNCH(KMAX)                                                                                                          \
    spa_candidate_kernel<KMAX><<<grid, kSpaThreads, 0, c.stream>>>(s.M, src, n_cand, s.spa_list.get(), s.spa_count.get(), n_var,      \
                                                                   s.cand.get(), s.xt.get(), s.mup.get(), s.ldx, s.spa.get(),         \
                                                                   s.counter.get(), s.out.get())
    if (K <= 4) SGB_SPA_LAUNCH(4);
    else if (K <= 8) SGB_SPA_LAUNCH(8);
    else if (K <= 16) SGB_SPA_LAUNCH(16);
    else SGB_SPA_LAUNCH(32);
#undef SGB_SPA_LAUNCH
    SGB_CHECK_LAUNCH();
    c.prof_end("spa_candidate_kernel");
    c.stats.n_kernel_launches++;
}

template <class Src>
void launch_per_variant(Context &c, ScoreState &s, const Src &src, int64_t n_items, const int32_t *list) {
    const int grid = (int)std::min<int64_t>(n_items, s.grid);
    if (s.M.trait == 0) s.spa.ensure((size_t)grid * 2 * s.M.n);      // saddle-point scratch of the blocks of this launch
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
void launch(Context &c, ScoreState &s, const Tiles &tiles, const Src &src, int64_t n_var, const uint8_t *tensor_base = nullptr,
            size_t tensor_pitch = 0) {
    if (s.path == SGB_SCORE_PER_VARIANT) {
        launch_per_variant(c, s, src, n_var, nullptr);
        return;
    }
    if (n_var > 0x7fffffff) throw Error(SGB_ERR_INVALID, "more than 2^31 - 1 variants in one batch");
    s.spa_list.ensure((size_t)n_var * kSpaBuckets);   // the tiled kernel fills list 0 only
    SGB_CUDA(cudaMemsetAsync(s.spa_count.get(), 0, sizeof(unsigned int) * kSpaBuckets, c.stream));
    const bool tensor = s.path == SGB_SCORE_TENSOR && s.tensor_ok && tensor_base != nullptr;
    if (tensor) {
        launch_tensor(c, s, tensor_base, tensor_pitch, n_var);
    } else {
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
    }
    c.d2h(s.h_count.p, s.spa_count.get(), sizeof(unsigned int) * kSpaBuckets);
    c.sync();
    if (tensor && *s.h_cerr.p != 0) {
        *s.h_cerr.p = 0;
        SGB_CUDA(cudaMemsetAsync(s.cerr.get(), 0, sizeof(int), c.stream));
        throw Error(SGB_ERR_CUDA, "the tensor-core score scan timed out waiting inside the GEMM kernel (SGB_WAIT_TIMEOUT_MS); "
                                  "sgb_score_test_set_path(ctx, SGB_SCORE_TILED) selects the kernel without such waits");
    }
    int64_t n_spa = 0;
    for (int b = 0; b < kSpaBuckets; b++) n_spa += s.h_count.p[b];
    if (n_spa > 0 && tensor) launch_candidates(c, s, PackedSrc{tensor_base, tensor_pitch}, n_spa, n_var);
    else if (n_spa > 0) launch_per_variant(c, s, src, n_spa, s.spa_list.get());
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
    s->spa_count.ensure(kSpaBuckets);
    s->h_count.ensure(kSpaBuckets);
    // tensor path: the 2K + 4 model columns [a | w x | y - mu | w | mu | 1], column-major, cut into digit planes once
    try {
        const bool bin = (m->trait == 0);
        s->ncols = (int)(2 * K + 4);
        const size_t nb = (n + 3) / 4, pitch = ((nb + 255) / 256) * 256;
        s->cpad = (int64_t)pitch * 4;
        const int groups = (s->ncols + kClassMaxCols - 1) / kClassMaxCols;
        std::vector<double> W((size_t)s->ncols * n);
        for (size_t i = 0; i < n; i++) {
            const double w = bin ? m->mu2[i] : 1.0;
            for (size_t k = 0; k < K; k++) {
                W[k * n + i] = m->t_XVX_inv_XV[i * K + k];
                W[(K + k) * n + i] = w * m->t_X[i * K + k];
            }
            W[2 * K * n + i] = m->y_mu[i];
            W[(2 * K + 1) * n + i] = w;
            W[(2 * K + 2) * n + i] = m->mu[i];
            W[(2 * K + 3) * n + i] = 1.0;
        }
        DevBuf<double> wdev;
        up(wdev, W.data(), W.size());
        s->ldx = (int64_t)((n + 3) / 4 * 4);
        std::vector<double> xt((size_t)K * s->ldx, 0.0), mup((size_t)s->ldx, 0.0);
        for (size_t i = 0; i < n; i++) {
            for (size_t k = 0; k < K; k++) xt[k * s->ldx + i] = m->t_X[i * K + k];
            mup[i] = m->mu[i];
        }
        up(s->xt, xt.data(), xt.size());
        up(s->mup, mup.data(), mup.size());
        s->cdig.ensure((size_t)groups * kClassDigitRows * s->cpad);
        s->cscal.ensure((size_t)s->ncols * kClassScal);
        s->ctot.ensure((size_t)s->ncols * 2);
        s->cerr.ensure(1);
        s->h_cerr.ensure(1);
        *s->h_cerr.p = 0;
        SGB_CUDA(cudaMemsetAsync(s->cerr.get(), 0, sizeof(int), c.stream));
        for (int g = 0, c0 = 0; c0 < s->ncols; g++, c0 += kClassMaxCols)
            umma_class_digits(c, wdev.get() + (size_t)c0 * n, (int64_t)n, std::min(kClassMaxCols, s->ncols - c0), s->cpad,
                              s->cdig.get() + (size_t)g * kClassDigitRows * s->cpad, s->cscal.get() + (size_t)c0 * kClassScal,
                              s->ctot.get() + (size_t)c0 * 2);
        c.sync();
        s->tensor_ok = true;
    } catch (const Error &e) {
        cudaGetLastError();
        s->tensor_ok = false;
        s->cdig.release();
        c.printf("note: tensor-core score scan unavailable (%s); using the shared-memory-tiled kernel\n", e.what());
    }
    const char *env_path = getenv("SGB_SCORE_PATH");
    const std::string ep = env_path ? env_path : "";
    s->path = ep == "per_variant" ? SGB_SCORE_PER_VARIANT : (ep == "tiled" || !s->tensor_ok) ? SGB_SCORE_TILED : SGB_SCORE_TENSOR;
    s->grid = c.sm_count * 4;
    s->grid_spa = c.sm_count * 2;   // blocks of spa_candidate_kernel (512 threads), scratch rows [0, grid_spa) of `spa`
    // saddle-point scratch (2 n doubles per block of the candidate kernel): binary traits only -- a quantitative trait never
    // takes the saddle-point branch (saige_main.cpp:322-350)
    // -- it is allocated by the kernel launches below for the blocks they really start (2.0 GB at n = 430K for the candidate kernel of
    // the tensor path, nothing for scans without candidates)
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
    if (path != SGB_SCORE_TILED && path != SGB_SCORE_PER_VARIANT && path != SGB_SCORE_TENSOR)
        throw Error(SGB_ERR_INVALID, "unknown score-test path");
    if (path == SGB_SCORE_TENSOR && !c.score->tensor_ok) throw Error(SGB_ERR_STATE, "the tensor-core score scan is not available for this model");
    c.score->path = path;
}

void score_test_packed(Context &c, const uint8_t *packed, int64_t nb, int64_t n_var, double *out, int32_t *valid) {
    ScoreState &s = state(c, n_var);
    if (!packed) throw Error(SGB_ERR_INVALID, "packed is NULL");
    if (nb != (s.M.n + 3) / 4) throw Error(SGB_ERR_INVALID, "n_bytes_per_variant must equal ceil(n_samp/4) of the model");
    if (s.path == SGB_SCORE_TENSOR && s.tensor_ok) {
        // rows at the 256-byte pitch the TMA boxes need; the bytes past nb meet zero digits
        const size_t pitch = (size_t)s.cpad / 4;
        s.geno.ensure(pitch * n_var);
        SGB_CUDA(cudaMemcpy2DAsync(s.geno.get(), pitch, packed, (size_t)nb, (size_t)nb, (size_t)n_var, cudaMemcpyHostToDevice, c.stream));
        launch(c, s, PackedTiles{s.geno.get(), pitch, nb}, PackedSrc{s.geno.get(), pitch}, n_var, s.geno.get(), pitch);
    } else {
        s.geno.ensure((size_t)nb * n_var);
        c.h2d(s.geno.get(), packed, (size_t)nb * n_var);
        launch(c, s, PackedTiles{s.geno.get(), (size_t)nb, nb}, PackedSrc{s.geno.get(), (size_t)nb}, n_var);
    }
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
    const bool tensor = (int64_t)c.pitch * 4 == s.cpad;
    launch(c, s, PackedTiles{c.packed.get() + (size_t)first * c.pitch, c.pitch, c.NB},
           PackedSrc{c.packed.get() + (size_t)first * c.pitch, c.pitch}, n_var, tensor ? c.packed.get() + (size_t)first * c.pitch : nullptr,
           c.pitch);
    SGB_CUDA(cudaEventRecord(c.ev1, c.stream));
    fetch(c, s, n_var, out, valid);
    if (kernel_ms) SGB_CUDA(cudaEventElapsedTime(kernel_ms, c.ev0, c.ev1));
}

}  // namespace sgb
