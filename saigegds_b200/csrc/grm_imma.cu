// int8 tensor-core implementation of get_crossprod_b_grm (saige_fitnull.cpp:435-536), exact-integer
// ("Ozaki"-sliced) arithmetic.  Why: the product costs 2 FP64 FMAs per genotype, the HBM roofline
// allows ~98 genotypes/clk/SM, the FP64 pipe 64 FMA/clk/SM and the issue ports 128 lanes/clk/SM
// (tools/microbench.cu, profiles/r01_microbench_b200.txt) -- a one-instruction-per-genotype kernel
// cannot get past ~30 % of the roofline.  Tensor cores consume 512 genotypes per instruction.
//
// Algebra.  With counts c in {0,1,2} (code 3 = missing), lut_j[c] = c*inv_j + lut0_j for valid calls:
//   dot_j = sum_n lut_j[c_nj] b_n         = inv_j * T_j + lut0_j * (Sb - U_j)
//   out_n = (1/M) sum_j dot_j lut_j[c_nj] = sum_j e_j c_nj + H - sum_{j in miss(n)} h_j
//   T_j = sum_n c_nj b_n, U_j = sum_{n in miss(j)} b_n, Sb = sum_n b_n,
//   e_j = dot_j inv_j / M, h_j = dot_j lut0_j / M, H = sum_j h_j.
// T = C'b and R = C e are integer-matrix x FP64-vector products.  b (resp. e) is quantised to a 56-bit
// fixed-point integer relative to its largest element and cut into eight signed base-128 digits; each
// digit plane is an exact u8 x s8 -> s32 tensor-core GEMM (mma.sync.m16n8k32, SASS IMMA.16832.U8.S8)
// with N = 8 digit planes.  The result equals the FP64 product of the quantised vector exactly; the
// quantisation error (2^-55 of the largest element) is below the rounding error of any FP64 dot product.
// Everything is integer and therefore order-independent and bit-reproducible.
//
// Operand formation costs one LOP3 per four genotypes: a packed 32-bit word holds 16 genotypes, and
// `w & (0x03030303 << 2t)` leaves four bytes each holding one genotype times 4^t -- a valid u8 operand.
// The factor 4^t is removed exactly from the int32 accumulator (one accumulator set per t).  Code 3 is
// NOT masked out: it contributes 3*b_n (resp. 3*e_j), which the sparse missing-genotype correction
// subtracts together with the mean-imputation terms (U_j and h_j above).
//
//   phase A (imma_dots_kernel):  CTA = 256 variants x a sample range; packed rows staged with cp.async;
//       A fragments straight from 128-bit shared loads; accumulators stay in registers over the sweep.
//   phase B (imma_apply_kernel): CTA = 1024 samples x a variant range; the same packed rows are read
//       through ldmatrix.m16n16.trans.b8 (SASS LDSM.8.MT1616), which transposes bytes in hardware so that
//       a register holds four *variants* of one sample byte.
// Two HBM passes over the packed matrix per product in this version (the fused single-pass kernel is
// the next step).
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include <algorithm>
#include <climits>
#include <type_traits>
#include <cmath>

#include <cuda.h>

#include "ctx.h"

namespace sgb {

struct ImmaPlan {
    // sparse lists of missing genotypes (samples < N only)
    DevBuf<int64_t> mv_ptr;  // [M+1] by variant -> sample ids
    DevBuf<int32_t> mv_idx;
    DevBuf<int64_t> ms_ptr;  // [N+1] by sample -> local variant ids
    DevBuf<int32_t> ms_idx;
    int64_t nnz = 0;
    int64_t ksteps = 0;      // ceil(N / 256)
    int64_t kblocks = 0;     // ceil(M / 32)
    int split_a = 1, split_b = 1;
    DevBuf<int8_t> dfrag;    // [ksteps][2048]  digits of b in phase-A fragment order
    DevBuf<int8_t> efrag;    // [kblocks][256]  digits of e in phase-B fragment order
    DevBuf<double> tq;       // [split_a][M]   T'_j partials in units of unit_b
    DevBuf<double> e, hm;    // [M]
    DevBuf<double> rpart;    // [split_b][N]   R_n partials in units of unit_e
    DevBuf<int64_t> mv_pos;  // [n_stiles * M + 1] tile-major start of (sample tile, variant)
    DevBuf<uint16_t> mv_i16; // [nnz] sample offset inside the tile
    DevBuf<int64_t> ms_pos;  // [n_vtiles * N + 1] tile-major start of (variant tile, sample)
    DevBuf<uint16_t> ms_i16; // [nnz] variant offset inside the tile
    // lane-interleaved ("ELL") copy of the two lists, what sparse_ell_sum_kernel reads: per (tile, group of 32 rows) a block
    // [k][lane] of 16-bit offsets padded to the longest row of the group with kSpTile (a zero slot of the vector tile)
    DevBuf<int64_t> mv_gstart, ms_gstart;   // [n_tiles * n_groups + 1] block starts (entries, multiples of 32)
    DevBuf<uint16_t> mv_ell, ms_ell;
    int um_pair = 2;         // batched GEMM kernel: 2 = pair kernel (clusters of two CTAs, cta_group::2 MMAs, one M-tile per CTA, eight A slots,
                             // four issuer threads: 4.6 ms per phase at K = 30, 3.5 ms for few columns), 0 = one CTA per MMA (5.0 / 4.0 ms),
                             // 1 = the single-CTA kernel's structure with pair MMAs (7.6 ms: hand-over loop too long); env SGB_UMMA_PAIR
    int um_fork = 0;         // batched path: sparse corrections on the side stream beside the GEMMs (env SGB_UMMA_FORK)
    int um_gather_w = 0;     // columns per pass of the row-gather kernel: 8, 16 or 32 (env SGB_UMMA_GATHER_W); 0 = by the column count
    int um_gather_v2 = -1;   // 16-byte loads in the row-gather kernel (env SGB_UMMA_GATHER_V2); -1 = with 32-column passes
    int um_gather_cols = 8;  // batched path: more columns than this take the row-gather sparse kernel (env SGB_UMMA_GATHER_COLS)
    bool sg_attr_set = false;   // imma_small_gemm_kernel: shared-memory attribute set / resident CTAs per SM on this device
    int sg_per_sm = 0;
    int um_small_cols = 4;   // batched path: up to this many columns the GEMMs run on mma.sync (imma_small_gemm_kernel; env SGB_UMMA_SMALL_COLS, 0 = never)
    int um_min_cols = 2;     // AUTO: batched path from this many columns (two columns: 5.3 ms with the mma.sync GEMMs against 2 x 3.4 ms
                             // for the fused single-RHS kernel and 8.3 ms on tcgen05 at N = 430K; env SGB_UMMA_MIN_COLS; 0 disables)
    bool use_csr = false;    // env SGB_SPARSE_CSR: the older row-per-thread kernel (comparison only)
    int n_stiles = 0, n_vtiles = 0;
    int opt_fork = 1, opt_fork_fused = -1, opt_grid_mult = 2, opt_stages = 3;   // tuning knobs (env: SGB_SPARSE_FORK, SGB_SPARSE_GRID_MULT, SGB_DOTS_STAGES)
    // fused single-pass kernel (grm_fused.cuh)
    bool fused_ok = false;
    int f_ks_per_cta = 0, f_grid = 0;   // half-steps (128 samples) per CTA slice, number of CTAs
    int64_t f_tiles = 0;
    DevBuf<int8_t> dfrag128;
    DevBuf<unsigned long long> f_acc;
    CUtensorMap f_tmap;      // packed matrix as a 2-D byte tensor [M][pitch], box 128 B x 32 rows, SWIZZLE_128B
    double f_efactor = 0;    // max_j |inv_j| sqrt(sum_n lut_j[c_nj]^2) / M_total (bound on |e_j| / |b|_2)
    int f_poll_ns = 30;
    int f_lag = 6;           // phase B runs this many tiles behind phase A (env SGB_FUSED_LAG)
    int64_t f_acc_stride = 0;
    DevBuf<double> f_rout, f_htotal, f_u, f_hpart;
    DevBuf<unsigned long long> f_edig;   // [f_tiles][32] digit blocks of e published by the tile owners
    DevBuf<int> f_err;
    PinBuf<int> f_herr;
    DevBuf<double> upart;    // [n_stiles][M] U_j per sample tile
    DevBuf<double> cpart;    // [n_vtiles][N] output correction per variant tile
    cudaStream_t side = nullptr;   // the sparse corrections run beside the tensor-core kernels
    cudaEvent_t ev_in = nullptr, ev_u = nullptr, ev_hm = nullptr, ev_corr = nullptr;
    ~ImmaPlan() {
        if (side) cudaStreamDestroy(side);
        for (cudaEvent_t e : {ev_in, ev_u, ev_hm, ev_corr}) if (e) cudaEventDestroy(e);
    }
    DevBuf<double> scal;     // [16] device scalars
    DevBuf<double> red;      // reduction partials
    DevBuf<unsigned int> counter;
    // batched K-RHS product on tcgen05 (grm_umma.cuh); prepared lazily by the first multi-column product
    struct Umma {
        bool ready = false, failed = false;
        DevBuf<uint8_t> pt;              // sample-major copy of the packed matrix [N][pitch_t]
        size_t pitch_t = 0;
        int64_t cpad_a = 0, cpad_b = 0;  // contraction capacity of phase A (samples) / phase B (variants), multiples of 512
        int boxes_a = 0, boxes_b = 0, split_a = 1, split_b = 1;
        DevBuf<int8_t> db, de;           // digit matrices [224][cpad]
        DevBuf<unsigned long long> t_lo, t_hi, r_lo, r_hi;   // exact limbs [32][M], [32][N]
        DevBuf<double> e, hm, u, corr;   // [32][M], [32][M], [32][M], [32][N]
        DevBuf<double> part;             // sparse partial sums of one column chunk: [4][tiles][rows]
        DevBuf<double> vt;               // transposed vector block [max(N, M)][32] of the row-gather kernel
        DevBuf<double> scal, red;
        DevBuf<unsigned int> counter;
        DevBuf<int> err;
        DevBuf<long long> prof;
        PinBuf<int> herr;
        CUtensorMap tmap_p, tmap_pt, tmap_p128, tmap_pt128, tmap_db[15], tmap_de[15];
        int split_a2 = 1, split_b2 = 1;   // splits of the pair kernel (a cluster = 256 rows on two SMs)
        bool have_d[15] = {false};
    } um;
};

namespace {

// scal slots
enum { S_MAXB = 0, S_SUMB = 1, S_UNITB = 2, S_MAXE = 3, S_H = 4, S_UNITE = 5, S_FESH = 6, S_FUNITE = 7, S_FEBOUND = 8 };

constexpr int kAThreads = 128;   // phase A: 4 warps x 4 row-blocks x 16 variants
constexpr int kARB = 4;
constexpr int kAVar = 4 * kARB * 16;          // 256 variants per CTA
constexpr int kAStageSteps = 2;               // K-steps (256 samples = 64 B per row) per pipeline stage
// pipeline depth is a template parameter (3 stages of 36 KB x 2 CTAs/SM = 221 KB is the default)
constexpr int kARowBytes = 64 * kAStageSteps; // 128 B per row per stage
constexpr int kAStageBytes = kAVar * kARowBytes + 2048 * kAStageSteps;

constexpr int kBThreads = 128;   // phase B: 4 warps x 4 row-blocks x 64 samples
constexpr int kBRB = 4;
constexpr int kBSamp = 4 * kBRB * 64;         // 1024 samples = 256 packed bytes per variant
constexpr int kBRowBytes = kBSamp / 4;        // 256
constexpr int kBPitch = kBRowBytes + 16;      // 272: 8 consecutive rows hit 8 different 16-byte bank groups
constexpr int kBStageVar = 64;                // variants per pipeline stage (2 K-blocks)
constexpr int kBStages = 4;
constexpr int kBStageBytes = kBStageVar * kBPitch + 256 * (kBStageVar / 32);

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void imma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// 56-bit fixed point -> eight signed base-128 digits (d_l in [-64, 63], top digit takes the rest)
__device__ __forceinline__ void to_digits(double v, int shift, int8_t (&d)[8]) {
    long long B = __double2ll_rn(scalbn(v, shift));
#pragma unroll
    for (int l = 0; l < 7; l++) {
        int dl = (int)(((B & 127) ^ 64) - 64);
        d[l] = (int8_t)dl;
        B = (B - dl) >> 7;
    }
    d[7] = (int8_t)B;
}

// shift so that |v| * 2^shift <= 2^55 for |v| <= maxabs; unit = 2^-shift (0 when the vector is all zero, NaN if not finite)
__device__ __forceinline__ int quant_shift(double maxabs, double *unit) {
    if (!(maxabs > 0) || !isfinite(maxabs)) {
        *unit = (maxabs == 0) ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
        return 0;
    }
    int sh = 54 - ilogb(maxabs);
    *unit = scalbn(1.0, -sh);
    return sh;
}

// ------------------------------------------------------------------------------------------------
// deterministic block reductions (fixed tree + last block adds partials in order)
template <int T>
__device__ __forceinline__ double block_reduce_sum(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0) for (int w = 0; w < T / 32; w++) t += sm[w];
    __syncthreads();
    return t;
}
template <int T>
__device__ __forceinline__ double block_reduce_max(double v, double *sm) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    if (threadIdx.x == 0) for (int w = 0; w < T / 32; w++) t = fmax(t, sm[w]);
    __syncthreads();
    return t;
}

// max|b| (NaN-propagating through the sum), sum(b) and sum(b^2): partial[0][G] = max, partial[1][G] = sum, partial[2][G] = sumsq.
// efactor > 0 (fused kernel): also the a-priori bound |e_j| <= efactor * |b|_2 and the fixed-point exponent of e derived
// from it (one bit of head room for the rounding of the bound itself).
__global__ void __launch_bounds__(256) absmax_sum_kernel(const double *__restrict__ b, int64_t N, double *partial,
                                                         unsigned int *counter, double *scal, double efactor) {
    __shared__ double sm[8];
    __shared__ bool last;
    double mx = 0, s = 0, s2 = 0;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < N; i += (int64_t)gridDim.x * 256) {
        double v = b[i];
        mx = fmax(mx, fabs(v));
        s += v;
        s2 += v * v;
    }
    mx = block_reduce_max<256>(mx, sm);
    s = block_reduce_sum<256>(s, sm);
    s2 = block_reduce_sum<256>(s2, sm);
    const int G = gridDim.x;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = mx;
        partial[G + blockIdx.x] = s;
        partial[2 * G + blockIdx.x] = s2;
        __threadfence();
        last = atomicInc(counter, G - 1) == (unsigned)(G - 1);
    }
    __syncthreads();
    if (!last) return;
    // the last block adds the partials with a fixed tree (deterministic), all threads taking part
    __threadfence();
    {
        const volatile double *p = partial;
        double m1 = 0, t1 = 0, t21 = 0;
        for (int i = threadIdx.x; i < G; i += 256) { m1 = fmax(m1, p[i]); t1 += p[G + i]; t21 += p[2 * G + i]; }
        mx = block_reduce_max<256>(m1, sm);
        s = block_reduce_sum<256>(t1, sm);
        s2 = block_reduce_sum<256>(t21, sm);
    }
    if (threadIdx.x == 0) {
        double m = mx, t = s, t2 = s2;
        if (!isfinite(t)) m = t;     // NaN / Inf anywhere in b poisons the product like it does in the reference
        scal[S_MAXB] = m;
        scal[S_SUMB] = t;
        double unit;
        quant_shift(m, &unit);
        scal[S_UNITB] = unit;
        if (efactor > 0) {
            // |b|_2, falling back to sqrt(N) max|b| when the sum of squares over- or underflowed
            double nrm = sqrt((double)N) * m;
            if (isfinite(t2) && m * m > 1e-290 && t2 >= 0.999 * m * m) nrm = fmin(nrm, sqrt(t2) * (1.0 + 1e-9));
            double eb = efactor * nrm;
            if (!isfinite(t)) eb = t;
            double eunit = 0, esh = 0;
            if (eb > 0 && isfinite(eb)) {
                const int sh = 53 - ilogb(eb);
                esh = (double)sh;
                eunit = scalbn(1.0, -sh);
            } else if (eb != 0) {
                eunit = __longlong_as_double(0x7ff8000000000000LL);
            }
            scal[S_FEBOUND] = eb;
            scal[S_FESH] = esh;
            scal[S_FUNITE] = eunit;
        }
    }
}

// digits of b in phase-A fragment order: [kstep][lane = l*4+tq][(pi*4+t0)*2+h][beta]
__global__ void digits_b_kernel(const double *__restrict__ b, int64_t N, int64_t Npad, const double *__restrict__ scal,
                                int8_t *__restrict__ dfrag) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Npad) return;
    double unit;
    const int sh = quant_shift(scal[S_MAXB], &unit);
    int8_t d[8];
    const double v = (n < N && unit > 0) ? b[n] : 0.0;
    to_digits(v, sh, d);
    const int64_t ks = n >> 8;
    const int r = (int)(n & 255), tq = r >> 6, r2 = r & 63, pi = r2 >> 5, h = (r2 >> 4) & 1, beta = (r2 >> 2) & 3, t0 = r2 & 3;
    int8_t *base = dfrag + ks * 2048 + tq * 64 + ((pi * 4 + t0) * 2 + h) * 4 + beta;
#pragma unroll
    for (int l = 0; l < 8; l++) base[l * 256] = d[l];
}

// digits of e in phase-B fragment order: [kblock][lane = l*4+tq][h][beta]
__global__ void digits_e_kernel(const double *__restrict__ e, int64_t M, int64_t Mpad, const double *__restrict__ scal,
                                int8_t *__restrict__ efrag) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Mpad) return;
    double unit;
    const int sh = quant_shift(scal[S_MAXE], &unit);
    int8_t d[8];
    const double v = (j < M && unit > 0) ? e[j] : 0.0;
    to_digits(v, sh, d);
    const int64_t kb = j >> 5;
    const int jj = (int)(j & 31), h = jj >> 4, tq = (jj >> 2) & 3, beta = jj & 3;
    int8_t *base = efrag + kb * 256 + tq * 8 + h * 4 + beta;
#pragma unroll
    for (int l = 0; l < 8; l++) base[l * 32] = d[l];
}

// ------------------------------------------------------------------------------------------------
// Phase A: T'_j = sum_n c'_nj * B_n  (c' = raw 2-bit code, B_n = quantised b) for 256 variants x a K-step range.
template <int kAStages>
__global__ void __launch_bounds__(kAThreads, 2) imma_dots_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M,
                                                                 int64_t ksteps, int split, const int8_t *__restrict__ dfrag,
                                                                 double *__restrict__ tq_out) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int64_t j0 = (int64_t)blockIdx.x * kAVar;
    const int sp = blockIdx.y;
    const int64_t ks_per = (ksteps + split - 1) / split;
    const int64_t ks_begin = (int64_t)sp * ks_per, ks_end = min(ksteps, ks_begin + ks_per);
    const int64_t n_stage = (ks_end > ks_begin) ? (ks_end - ks_begin + kAStageSteps - 1) / kAStageSteps : 0;

    auto issue = [&](int64_t st) {
        if (st < n_stage) {
            uint8_t *buf = smem + (size_t)(st % kAStages) * kAStageBytes;
            const int64_t ks0 = ks_begin + st * kAStageSteps;
            // packed rows: 256 rows x (4 * kAStageSteps) granules of 16 B.  A quarter-warp reads 4 granules of two
            // adjacent rows: with 64-byte rows that is one contiguous 128-byte line; with 128-byte rows the granule
            // index of odd rows is flipped by 4 so that the two rows use different bank halves.
            constexpr int kGran = kARowBytes / 16;
            for (int i = tid; i < kAVar * kGran; i += kAThreads) {
                const int r = i / kGran, gran = i % kGran;
                const int64_t j = min(j0 + r, M - 1);
                const int64_t ks = ks0 + (gran >> 2);
                const size_t off = (size_t)min(ks, ksteps - 1) * 64 + (size_t)(gran & 3) * 16;
                const int sw = (kAStageSteps == 2) ? (gran ^ ((r & 1) << 2)) : gran;
                cp_async16(buf + r * kARowBytes + (sw << 4), packed + (size_t)j * pitch + off);
            }
            uint8_t *dbuf = buf + kAVar * kARowBytes;
            for (int i = tid; i < 128 * kAStageSteps; i += kAThreads) {
                const int64_t ks = min(ks0 + (i >> 7), ksteps - 1);
                cp_async16(dbuf + i * 16, dfrag + ks * 2048 + (size_t)(i & 127) * 16);
            }
        }
        cp_async_commit();
    };

    int acc[kARB][4][4];
#pragma unroll
    for (int rb = 0; rb < kARB; rb++)
#pragma unroll
        for (int t = 0; t < 4; t++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[rb][t][q] = 0;
    // FP64 running totals (flushed from the int32 accumulators well before they can overflow)
    double tot[kARB][2];
#pragma unroll
    for (int rb = 0; rb < kARB; rb++) tot[rb][0] = tot[rb][1] = 0;

    auto flush = [&]() {
        // value of this thread's two digit columns (2tq, 2tq+1) for rows g and g+8, with the 4^t factor removed
        const double w0 = scalbn(1.0, 14 * tq), w1 = scalbn(1.0, 14 * tq + 7);
#pragma unroll
        for (int rb = 0; rb < kARB; rb++) {
            double lo = 0, hi = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                lo += w0 * (double)(acc[rb][t][0] >> (2 * t)) + w1 * (double)(acc[rb][t][1] >> (2 * t));
                hi += w0 * (double)(acc[rb][t][2] >> (2 * t)) + w1 * (double)(acc[rb][t][3] >> (2 * t));
                acc[rb][t][0] = acc[rb][t][1] = acc[rb][t][2] = acc[rb][t][3] = 0;
            }
            tot[rb][0] += lo;
            tot[rb][1] += hi;
        }
    };

    for (int s = 0; s < kAStages - 1; s++) issue(s);
    int since_flush = 0;
    for (int64_t st = 0; st < n_stage; st++) {
        cp_async_wait<kAStages - 2>();
        __syncthreads();
        issue(st + kAStages - 1);
        const uint8_t *buf = smem + (size_t)(st % kAStages) * kAStageBytes;
        const uint8_t *dbuf = buf + kAVar * kARowBytes;
        const int64_t ks0 = ks_begin + st * kAStageSteps;
#pragma unroll
        for (int s = 0; s < kAStageSteps; s++) {
            if (ks0 + s >= ks_end) break;
            // B fragments: 64 contiguous bytes per lane
            uint4 bf[4];
            const uint4 *bp = reinterpret_cast<const uint4 *>(dbuf + s * 2048 + lane * 64);
#pragma unroll
            for (int i = 0; i < 4; i++) bf[i] = bp[i];
            const uint32_t *bw = reinterpret_cast<const uint32_t *>(bf);   // [(pi*4+t0)*2+h]
#pragma unroll
            for (int rb = 0; rb < kARB; rb++) {
                const int r0 = warp * (kARB * 16) + rb * 16 + g;
                const int gran = (kAStageSteps == 2) ? ((s * 4 + tq) ^ ((r0 & 1) << 2)) : (s * 4 + tq);   // r0, r0+8: same parity
                const uint4 wa = *reinterpret_cast<const uint4 *>(buf + r0 * kARowBytes + (gran << 4));
                const uint4 wb = *reinterpret_cast<const uint4 *>(buf + (r0 + 8) * kARowBytes + (gran << 4));
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    const uint32_t m = 0x03030303u << (2 * t);
                    imma_u8s8(acc[rb][t], wa.x & m, wb.x & m, wa.y & m, wb.y & m, bw[(0 * 4 + t) * 2], bw[(0 * 4 + t) * 2 + 1]);
                    imma_u8s8(acc[rb][t], wa.z & m, wb.z & m, wa.w & m, wb.w & m, bw[(1 * 4 + t) * 2], bw[(1 * 4 + t) * 2 + 1]);
                }
            }
        }
        // one accumulator sees 64 terms of at most 192*64 per K-step: flush every 1024 K-steps (< 2^30)
        since_flush += kAStageSteps;
        if (since_flush >= 1024) { flush(); since_flush = 0; }
    }
    cp_async_wait<0>();
    flush();
    // add the four digit-column groups (tq = 0..3) of each row; lanes tq == 0 write
#pragma unroll
    for (int rb = 0; rb < kARB; rb++) {
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
            double v = tot[rb][hh];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const int64_t j = j0 + warp * (kARB * 16) + rb * 16 + g + hh * 8;
            if (tq == 0 && j < M) tq_out[(size_t)sp * M + j] = v;
        }
    }
}

// Sparse missing-genotype sums, tiled so that the gathered vector sits in shared memory:
//   part[t][r] = sum over the entries of row r whose column lies in tile t of vec[column]
// Rows are variants (vec = b, result U_j) or samples (vec = hm, result corr_n).  The entries are stored
// tile-major -- [tile][row][ascending column] -- as 16-bit column offsets inside the tile, so a warp working
// on 32 consecutive rows of one tile reads one contiguous stretch of memory, and the gathers hit shared
// memory (8 bytes per value) instead of L2 (a 32-byte sector per value).  pos[t * R + r] is the start of
// (tile t, row r); pos has n_tiles * R + 1 entries.
constexpr int kSpTile = 4096;      // columns per tile: 32 KB of doubles
constexpr int kSpRows = 1024;      // rows per work item
constexpr int kSpThreads = 512;
constexpr int kSpCap = 24576;      // index entries staged per piece (48 KB); a work item holds ~ kSpRows * 20 at 0.5 % missing
constexpr int kSpSmem = kSpTile * 8 + kSpCap * 2 + (kSpRows + 1) * 8;
// Persistent (grid <= number of SMs).  Per work item (tile t, kSpRows rows): the vector tile, the rows' start
// offsets and the whole contiguous index range of those rows are streamed into shared memory with coalesced
// 16-byte cp.async copies (deep memory-level parallelism, no dependent global loads); then thread <-> row walks its
// own segment out of shared memory.  Summation order inside a row is fixed (four interleaved partial sums).
__global__ void __launch_bounds__(kSpThreads) sparse_tile_sum_kernel(const int64_t *__restrict__ pos, const uint16_t *__restrict__ idx16,
                                                                     const double *__restrict__ vec, int64_t R, int64_t C,
                                                                     int n_tiles, double *__restrict__ part) {
    extern __shared__ __align__(16) uint8_t smem_sp[];
    double *sv = reinterpret_cast<double *>(smem_sp);
    uint16_t *sidx = reinterpret_cast<uint16_t *>(smem_sp + kSpTile * 8);
    int64_t *spos = reinterpret_cast<int64_t *>(smem_sp + kSpTile * 8 + kSpCap * 2);
    __shared__ int64_t s_bounds[2][2];     // index range of this / the next work item (fetched one item ahead)
    const int64_t n_chunks = (R + kSpRows - 1) / kSpRows;
    const int64_t n_work = n_chunks * n_tiles;
    auto item_range = [&](int64_t w, int64_t &lo, int64_t &hi) {
        const int64_t t = w / n_chunks, chunk = w % n_chunks;
        lo = (int64_t)t * R + chunk * kSpRows;
        hi = (int64_t)t * R + min(R, (chunk + 1) * kSpRows);
    };
    if (threadIdx.x == 0 && (int64_t)blockIdx.x < n_work) {
        int64_t lo, hi;
        item_range(blockIdx.x, lo, hi);
        s_bounds[0][0] = pos[lo];
        s_bounds[0][1] = pos[hi];
    }
    int it = 0;
    for (int64_t w = blockIdx.x; w < n_work; w += gridDim.x, it++) {
        const int t = (int)(w / n_chunks);
        const int64_t chunk = w % n_chunks;
        const int64_t c0 = (int64_t)t * kSpTile;
        const int64_t r0 = chunk * kSpRows, r1 = min(R, r0 + kSpRows);
        const int nrow = (int)(r1 - r0);
        const int64_t *pp = pos + (size_t)t * R + r0;
        __syncthreads();                                   // previous work item fully consumed; s_bounds[it & 1] visible
        const int64_t e0 = s_bounds[it & 1][0], e1 = s_bounds[it & 1][1];
        const int64_t a0 = e0 & ~(int64_t)7;               // 16-byte aligned start of the index range
        // everything this item needs is requested at once (one memory latency per item): row offsets, vector tile, indices
        for (int i = threadIdx.x; i <= nrow; i += kSpThreads) cp_async8(spos + i, pp + i);
        for (int i = threadIdx.x; i < kSpTile; i += kSpThreads) {
            if (c0 + i < C) cp_async8(sv + i, vec + c0 + i); else sv[i] = 0.0;
        }
        {
            const int64_t pe = min(e1, a0 + kSpCap);
            const int ngran = (int)((pe - a0 + 7) >> 3);
            for (int i = threadIdx.x; i < ngran; i += kSpThreads) cp_async16(sidx + i * 8, idx16 + a0 + (int64_t)i * 8);
        }
        cp_async_commit();
        if (threadIdx.x == 0 && w + gridDim.x < n_work) {   // bounds of the next item: their latency hides behind this item's loads
            int64_t lo, hi;
            item_range(w + gridDim.x, lo, hi);
            s_bounds[(it + 1) & 1][0] = pos[lo];
            s_bounds[(it + 1) & 1][1] = pos[hi];
        }
        double acc[kSpRows / kSpThreads];
#pragma unroll
        for (int k = 0; k < kSpRows / kSpThreads; k++) acc[k] = 0;
        for (int64_t pc = a0; pc < e1 || pc == a0; pc += kSpCap) {     // almost always a single piece
            const int64_t pe = min(e1, pc + kSpCap);
            if (pc != a0) {
                const int ngran = (int)((pe - pc + 7) >> 3);
                __syncthreads();
                for (int i = threadIdx.x; i < ngran; i += kSpThreads) cp_async16(sidx + i * 8, idx16 + pc + (int64_t)i * 8);
                cp_async_commit();
            }
            cp_async_wait<0>();
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kSpRows / kSpThreads; k++) {
                const int r = threadIdx.x + k * kSpThreads;
                if (r < nrow) {
                    // entries [lo, hi) of this row inside the staged piece, as 32-bit offsets; four indices per 64-bit shared
                    // load once the offset is 4-aligned (the shared-memory pipe, not HBM, bounds this kernel)
                    const int lo = (int)(max(spos[r], pc) - pc), hi = (int)(min(spos[r + 1], pe) - pc);
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                    int i = lo;
                    for (; i < hi && (i & 3); i++) s0 += sv[sidx[i]];
                    for (; i + 4 <= hi; i += 4) {
                        const uint2 q = *reinterpret_cast<const uint2 *>(sidx + i);
                        s0 += sv[q.x & 0xffffu]; s1 += sv[q.x >> 16]; s2 += sv[q.y & 0xffffu]; s3 += sv[q.y >> 16];
                    }
                    for (; i < hi; i++) s1 += sv[sidx[i]];
                    if (hi > lo) acc[k] += (s0 + s1) + (s2 + s3);
                }
            }
            if (e1 <= a0) break;
        }
#pragma unroll
        for (int k = 0; k < kSpRows / kSpThreads; k++) {
            const int r = threadIdx.x + k * kSpThreads;
            if (r < nrow) part[(size_t)t * R + r0 + r] = acc[k];
        }
    }
}

// ---- bulk-copy (TMA) helpers of the sparse kernel ---------------------------------------------------------------------
__device__ __forceinline__ unsigned sp_smem(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sp_mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sp_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void sp_mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sp_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sp_mbar_wait(unsigned long long *b, unsigned parity) {
    const unsigned a = sp_smem(b);
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void sp_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sp_smem(dst)), "l"(src), "r"(bytes), "r"(sp_smem(bar)) : "memory");
}

// ---- lane-interleaved layout ------------------------------------------------------------------------------------------
// The row-per-thread walk above spends ~24 instructions per entry (divergent trip counts, 64-bit bookkeeping) and is bound
// by instruction issue, not by memory.  Here the entries of 32 consecutive rows are stored interleaved, [k][lane], padded to
// the longest row of the group with the index of a zero slot: a warp walks its group with one coalesced 64-byte index
// load, one gather and one add per 32 entries, no branches and no per-row offsets (1.5x the index bytes at 0.5 % missing).
constexpr int kEllRows = 1024;                    // rows per work item = 32 groups
constexpr int kEllThreads = 512;                  // 16 warps x 2 groups
constexpr int kEllCap = 36864;                    // index entries staged per piece (72 KB); an item holds ~1024 x 30 at 0.5 % missing
constexpr int kEllSv = kSpTile + 8;               // vector tile + the zero slot (padded to 64 bytes)
constexpr int kEllSmem = kEllSv * 8 + 2 * kEllCap * 2 + 2 * 40 * 8;   // vector tile + two index stages + two sets of group starts

// len32[t * G + g] = 32 * (longest row of group g in tile t); pos = the tile-major CSR starts
__global__ void ell_len_kernel(const int64_t *__restrict__ pos, int64_t R, int64_t G, int n_tiles, int64_t *__restrict__ len32) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= G * n_tiles) return;
    const int64_t t = i / G, g = i % G;
    int64_t mx = 0;
    for (int r = 0; r < 32; r++) {
        const int64_t row = g * 32 + r;
        if (row < R) mx = max(mx, pos[t * R + row + 1] - pos[t * R + row]);
    }
    len32[i] = mx * 32;
}
// one warp per (tile, group): lane <-> row.  The order of a row's entries is free, and the gather from the shared-memory vector
// tile is what bounds the kernel (a random 8-byte gather costs ~6 shared-memory wavefronts per warp: two half-warps x ~3
// lanes on the fullest of the 16 eight-byte banks).  Entries are therefore placed so that at step k lane l reads bank
// (k + l) mod 16 whenever the row has such an entry -- distinct banks within each half-warp; entries that find their
// preferred steps taken go to the last free step, the rest of the column is padding (one address: a broadcast).
__global__ void ell_fill_kernel(const int64_t *__restrict__ pos, const uint16_t *__restrict__ i16, int64_t R, int64_t G, int n_tiles,
                                const int64_t *__restrict__ gstart, uint16_t *__restrict__ ell) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= G * n_tiles) return;
    const int64_t t = i / G, g = i % G, row = g * 32 + lane;
    const int64_t s = gstart[i], len = (gstart[i + 1] - s) >> 5;
    int64_t p0 = 0, mylen = 0;
    if (row < R) { p0 = pos[t * R + row]; mylen = pos[t * R + row + 1] - p0; }
    uint16_t *col = ell + s + lane;                          // this row's column: col[k * 32]
    for (int64_t k = 0; k < len; k++) col[k * 32] = 0xFFFFu;  // free
    int64_t last_free = len - 1;
    for (int64_t e = 0; e < mylen; e++) {
        const uint16_t v = i16[p0 + e];
        int64_t k = ((int)(v & 15) - lane) & 15;
        while (k < len && col[k * 32] != 0xFFFFu) k += 16;
        if (k >= len) {
            while (col[last_free * 32] != 0xFFFFu) last_free--;
            k = last_free;
        }
        col[k * 32] = v;
    }
    for (int64_t k = 0; k < len; k++)
        if (col[k * 32] == 0xFFFFu) col[k * 32] = (uint16_t)kSpTile;
}

// Persistent, one CTA per SM, two index stages: while the warps gather item i out of one stage, the bulk copy of item i + 1
// is already in flight into the other (an item = 32 groups = 1024 rows of one tile; a CTA walks a contiguous range of the
// tile-major item list, so the vector tile is reloaded only when the tile changes).  G is even and every block start a multiple
// of 32 entries, so the group starts and the index block are 16-byte aligned bulk copies counted on one mbarrier per stage.
__global__ void __launch_bounds__(kEllThreads) sparse_ell_sum_kernel(const int64_t *__restrict__ gstart, const uint16_t *__restrict__ ell,
                                                                     const double *__restrict__ vec, int64_t R, int64_t C, int64_t G,
                                                                     int n_tiles, double *__restrict__ part) {
    extern __shared__ __align__(16) uint8_t smem_sp[];
    double *sv = reinterpret_cast<double *>(smem_sp);
    uint16_t *sidx0 = reinterpret_cast<uint16_t *>(smem_sp + kEllSv * 8);
    int64_t *sgs0 = reinterpret_cast<int64_t *>(smem_sp + kEllSv * 8 + 2 * kEllCap * 2);          // 2 x 40 group starts
    __shared__ unsigned long long s_bar[2];
    unsigned ph[2] = {0, 0};
    if (threadIdx.x == 0) {
        sp_mbar_init(&s_bar[0], 1);
        sp_mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_chunks = (G + 31) / 32;
    const int64_t n_work = n_chunks * n_tiles;
    const int64_t w_begin = n_work * blockIdx.x / gridDim.x, w_end = n_work * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x < 8) sv[kSpTile + threadIdx.x] = 0.0;                                 // the zero slot
    // thread 0 only: index range of an item (two global loads; fetched one item before they are needed)
    auto bounds = [&](int64_t w, int64_t &e0, int64_t &e1) {
        const int64_t t = w / n_chunks, chunk = w % n_chunks;
        e0 = gstart[t * G + chunk * 32];
        e1 = gstart[t * G + min(G, (chunk + 1) * 32)];
    };
    // thread 0 only: request item w into `stage` (group starts + first piece of the index block [+ the vector tile])
    auto issue = [&](int64_t w, int stage, int64_t e0, int64_t e1, bool with_sv) {
        const int64_t t = w / n_chunks, chunk = w % n_chunks, g0 = chunk * 32, c0 = t * kSpTile;
        const int ng = (int)(min(G, g0 + 32) - g0);
        const unsigned nb_gs = (unsigned)(((ng + 2) & ~1) * 8);
        const unsigned nb_idx = (unsigned)(min(e1 - e0, (int64_t)kEllCap) * 2);
        const int ncol = (int)min((int64_t)kSpTile, C - c0);
        const unsigned nb_sv = with_sv ? (unsigned)ncol * 8u : 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // earlier generic accesses of the stage
        sp_mbar_expect_tx(&s_bar[stage], nb_gs + nb_idx + nb_sv);
        sp_bulk_g2s(sgs0 + stage * 40, gstart + t * G + g0, nb_gs, &s_bar[stage]);
        if (nb_idx) sp_bulk_g2s(sidx0 + (size_t)stage * kEllCap, ell + e0, nb_idx, &s_bar[stage]);
        if (nb_sv) sp_bulk_g2s(sv, vec + c0, nb_sv, &s_bar[stage]);
    };
    int64_t nx0 = 0, nx1 = 0;            // thread 0: bounds of the item that will be requested next
    if (threadIdx.x == 0 && w_begin < w_end) bounds(w_begin, nx0, nx1);
    int cur_tile = -1;
    int it = 0;
    for (int64_t w = w_begin; w < w_end; w++, it++) {
        const int stage = it & 1;
        const int t = (int)(w / n_chunks);
        const int64_t chunk = w % n_chunks, g0 = chunk * 32, c0 = (int64_t)t * kSpTile;
        const int ng = (int)(min(G, g0 + 32) - g0);
        const bool new_tile = t != cur_tile;
        cur_tile = t;
        __syncthreads();                 // item w - 1 fully consumed: its stage and (on a tile change) the vector tile are free
        if (new_tile) {
            // not requested ahead: the vector tile has to change first
            const int ncol = (int)min((int64_t)kSpTile, C - c0);
            const bool sv_bulk = ((ncol & 1) == 0) && ((reinterpret_cast<uintptr_t>(vec + c0) & 15) == 0);
            if (!sv_bulk)
                for (int i = threadIdx.x; i < ncol; i += kEllThreads) cp_async8(sv + i, vec + c0 + i);
            for (int i = ncol + threadIdx.x; i < kSpTile; i += kEllThreads) sv[i] = 0.0;
            cp_async_commit();
            if (threadIdx.x == 0) {
                issue(w, stage, nx0, nx1, sv_bulk);
                if (w + 1 < w_end) bounds(w + 1, nx0, nx1);
            }
            cp_async_wait<0>();
        }
        // request item w + 1 (same tile) into the other stage before touching item w
        if (threadIdx.x == 0 && w + 1 < w_end && (int)((w + 1) / n_chunks) == t) {
            issue(w + 1, stage ^ 1, nx0, nx1, false);
            if (w + 2 < w_end) bounds(w + 2, nx0, nx1);
        }
        sp_mbar_wait(&s_bar[stage], ph[stage]);
        ph[stage] ^= 1;
        if (new_tile) __syncthreads();   // cp.async / plain stores of the vector tile by other threads
        const uint16_t *sidx = sidx0 + (size_t)stage * kEllCap;
        const int64_t *sgs = sgs0 + stage * 40;
        const int64_t e0 = sgs[0], e1 = sgs[ng];
        double acc[2] = {0, 0};
        for (int64_t pc = e0;; pc += kEllCap) {                           // almost always a single piece
            const int64_t pe = min(e1, pc + kEllCap);
            if (pc != e0) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    sp_mbar_expect_tx(&s_bar[stage], (unsigned)((pe - pc) * 2));
                    sp_bulk_g2s(sidx0 + (size_t)stage * kEllCap, ell + pc, (unsigned)((pe - pc) * 2), &s_bar[stage]);
                }
                sp_mbar_wait(&s_bar[stage], ph[stage]);
                ph[stage] ^= 1;
            }
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const int gi = warp + 16 * k;
                if (gi < ng) {
                    const int lo = (int)(max(sgs[gi], pc) - pc), hi = (int)(min(sgs[gi + 1], pe) - pc);
                    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                    int i = lo + lane;
                    for (; i + 96 < hi; i += 128) {
                        s0 += sv[sidx[i]]; s1 += sv[sidx[i + 32]]; s2 += sv[sidx[i + 64]]; s3 += sv[sidx[i + 96]];
                    }
                    for (; i < hi; i += 32) s0 += sv[sidx[i]];
                    acc[k] += (s0 + s1) + (s2 + s3);
                }
            }
            if (pe >= e1) break;
        }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int gi = warp + 16 * k;
            const int64_t r = (g0 + gi) * 32 + lane;
            if (gi < ng && r < R) part[(size_t)t * R + r] = acc[k];
        }
    }
}

// cnt[t * R + r] = number of entries of row r (sorted list idx[ptr[r] .. ptr[r+1])) with column in tile t
__global__ void sparse_tile_count_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, int64_t R,
                                         int64_t *__restrict__ cnt) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (r >= R) return;
    const int64_t lo0 = ptr[r], hi0 = ptr[r + 1];
    int64_t b[2];
    for (int e = 0; e < 2; e++) {
        int64_t lo = lo0, hi = hi0;
        const int64_t key = (int64_t)(t + e) * kSpTile;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (idx[mid] < key) lo = mid + 1; else hi = mid;
        }
        b[e] = lo;
    }
    cnt[(size_t)t * R + r] = b[1] - b[0];
}

// scatter the row-major lists into the tile-major 16-bit layout
__global__ void sparse_tile_fill_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, int64_t R,
                                        const int64_t *__restrict__ pos, uint16_t *__restrict__ idx16) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (r >= R) return;
    int64_t lo = ptr[r], hi = ptr[r + 1];
    const int64_t key = (int64_t)t * kSpTile;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    const int64_t o = pos[(size_t)t * R + r], n = pos[(size_t)t * R + r + 1] - o;
    for (int64_t i = 0; i < n; i++) idx16[o + i] = (uint16_t)(idx[lo + i] - key);
}

// dot_j, e_j, hm_j = h_j + 3 e_j;  max|e| and H = sum h_j (deterministic)
__global__ void __launch_bounds__(256) finalize_dots_kernel(const double *__restrict__ tq, int split, const double *__restrict__ upart,
                                                            int n_utiles, const double *__restrict__ lut, int64_t M, double inv_mtotal,
                                                            double *__restrict__ e, double *__restrict__ hm, double *partial,
                                                            unsigned int *counter, double *scal) {
    __shared__ double sm[8];
    __shared__ bool last;
    const double unit_b = scal[S_UNITB], sumb = scal[S_SUMB];
    double mx = 0, hs = 0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < M; j += (int64_t)gridDim.x * 256) {
        double t = 0;
        for (int s = 0; s < split; s++) t += tq[(size_t)s * M + j];
        const double l0 = lut[4 * j], inv = lut[4 * j + 1] - l0;
        double uj = 0;
        for (int t = 0; t < n_utiles; t++) uj += upart[(size_t)t * M + j];
        const double T = (unit_b == 0 ? 0.0 : unit_b * t) - 3.0 * uj;
        const double dot = inv * T + l0 * (sumb - uj);
        const double ej = dot * inv * inv_mtotal, hj = dot * l0 * inv_mtotal;
        e[j] = ej;
        hm[j] = hj + 3.0 * ej;
        mx = fmax(mx, fabs(ej));
        hs += hj;
        if (!isfinite(ej)) mx = ej;
    }
    // NaN-safe max: propagate non-finite values through the sum channel
    mx = block_reduce_max<256>(mx, sm);
    hs = block_reduce_sum<256>(hs, sm);
    const int G = gridDim.x;
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = mx;
        partial[G + blockIdx.x] = hs;
        __threadfence();
        last = atomicInc(counter, G - 1) == (unsigned)(G - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const volatile double *p = partial;
        double m = 0, t = 0;
        for (int i = 0; i < G; i++) { m = fmax(m, p[i]); t += p[G + i]; }
        if (!isfinite(t)) m = t;
        scal[S_MAXE] = m;
        scal[S_H] = t;
        double unit;
        quant_shift(m, &unit);
        scal[S_UNITE] = unit;
    }
}

// ------------------------------------------------------------------------------------------------
// Phase B: R_n = sum_j c'_nj * E_j for 1024 samples x a K-block range.
__global__ void __launch_bounds__(kBThreads, 2) imma_apply_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M,
                                                                  int64_t N, int64_t kblocks, int split,
                                                                  const int8_t *__restrict__ efrag, double *__restrict__ rpart) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const size_t byte0 = (size_t)blockIdx.x * kBRowBytes;
    const int sp = blockIdx.y;
    const int64_t kb_per = (kblocks + split - 1) / split;
    const int64_t kb_begin = (int64_t)sp * kb_per, kb_end = min(kblocks, kb_begin + kb_per);
    constexpr int kKbPerStage = kBStageVar / 32;
    const int64_t n_stage = (kb_end > kb_begin) ? (kb_end - kb_begin + kKbPerStage - 1) / kKbPerStage : 0;

    auto issue = [&](int64_t st) {
        if (st < n_stage) {
            uint8_t *buf = smem + (size_t)(st % kBStages) * kBStageBytes;
            const int64_t v0 = (kb_begin + st * kKbPerStage) * 32;
            for (int i = tid; i < kBStageVar * (kBRowBytes / 16); i += kBThreads) {
                const int r = i >> 4, gran = i & 15;
                const int64_t j = min(v0 + r, M - 1);
                cp_async16(buf + r * kBPitch + gran * 16, packed + (size_t)j * pitch + byte0 + (size_t)gran * 16);
            }
            uint8_t *ebuf = buf + kBStageVar * kBPitch;
            for (int i = tid; i < 16 * kKbPerStage; i += kBThreads) {
                const int64_t kb = min(kb_begin + st * kKbPerStage + (i >> 4), kblocks - 1);
                cp_async16(ebuf + i * 16, efrag + kb * 256 + (size_t)(i & 15) * 16);
            }
        }
        cp_async_commit();
    };

    int acc[kBRB][4][4];
#pragma unroll
    for (int rb = 0; rb < kBRB; rb++)
#pragma unroll
        for (int t = 0; t < 4; t++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[rb][t][q] = 0;
    // The int32 accumulators are turned into FP64 and written (first time) or added (later) to the CTA's own
    // slice of rpart; no FP64 running totals are kept in registers (keeps the kernel at <= 168 registers so that a
    // sparse-correction CTA fits beside two of these CTAs on an SM).
    bool first_flush = true;
    auto flush = [&]() {
        const double w0 = scalbn(1.0, 14 * tq), w1 = scalbn(1.0, 14 * tq + 7);
#pragma unroll
        for (int rb = 0; rb < kBRB; rb++)
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    // t = 0 holds the UNMASKED bytes (c0 + 4 c1 + 16 c2 + 64 c3 meet the same multiplier): plane 0 by subtraction
                    int s0, s1;
                    if (t == 0) {
                        s0 = acc[rb][0][2 * hh] - acc[rb][1][2 * hh] - acc[rb][2][2 * hh] - acc[rb][3][2 * hh];
                        s1 = acc[rb][0][2 * hh + 1] - acc[rb][1][2 * hh + 1] - acc[rb][2][2 * hh + 1] - acc[rb][3][2 * hh + 1];
                    } else {
                        s0 = acc[rb][t][2 * hh] >> (2 * t);
                        s1 = acc[rb][t][2 * hh + 1] >> (2 * t);
                    }
                    double v = w0 * (double)s0 + w1 * (double)s1;
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    const int64_t n = ((int64_t)byte0 + (warp * kBRB + rb) * 16 + g + hh * 8) * 4 + t;
                    if (tq == 0 && n < N) {
                        double *dst = rpart + (size_t)sp * N + n;
                        *dst = first_flush ? v : (*dst + v);
                    }
                }
#pragma unroll
        for (int rb = 0; rb < kBRB; rb++)
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[rb][t][q] = 0;
        first_flush = false;
    };

    // ldmatrix row addresses: lanes 0-15 -> variants 0..15 of the K-block, lanes 16-31 -> variants 16..31
    const int ld_row = lane;   // 0..31
    for (int s = 0; s < kBStages - 1; s++) issue(s);
    int since_flush = 0;
    for (int64_t st = 0; st < n_stage; st++) {
        cp_async_wait<kBStages - 2>();
        __syncthreads();
        issue(st + kBStages - 1);
        const uint8_t *buf = smem + (size_t)(st % kBStages) * kBStageBytes;
        const uint8_t *ebuf = buf + kBStageVar * kBPitch;
#pragma unroll
        for (int kk = 0; kk < kKbPerStage; kk++) {
            if (kb_begin + st * kKbPerStage + kk >= kb_end) break;
            const uint2 bf = *reinterpret_cast<const uint2 *>(ebuf + kk * 256 + lane * 8);
#pragma unroll
            for (int rb = 0; rb < kBRB; rb++) {
                // 16 byte positions (64 samples) of this row-block, 32 variants of this K-block
                const uint8_t *src = buf + (kk * 32 + ld_row) * kBPitch + (warp * kBRB + rb) * 16;
                const unsigned addr = (unsigned)__cvta_generic_to_shared(src);
                uint32_t x0, x1, x2, x3;   // x0/x1: variants 0-15 at byte g / g+8; x2/x3: variants 16-31
                asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
                             : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(addr));
                imma_u8s8(acc[rb][0], x0, x1, x2, x3, bf.x, bf.y);
#pragma unroll
                for (int t = 1; t < 4; t++) {
                    const uint32_t m = 0x03030303u << (2 * t);
                    imma_u8s8(acc[rb][t], x0 & m, x1 & m, x2 & m, x3 & m, bf.x, bf.y);
                }
            }
        }
        // one accumulator sees 32 terms of at most 255*64 per K-block (the unmasked plane): flush every 2048 K-blocks (< 2^31)
        since_flush += kKbPerStage;
        if (since_flush >= 2048) { flush(); since_flush = 0; }
    }
    cp_async_wait<0>();
    flush();
}

// out_n = unit_e * sum_s R_n[s] + H - corr_n
__global__ void combine_kernel(const double *__restrict__ rpart, int split, int64_t N, const double *__restrict__ cpart,
                               int n_ctiles, const double *__restrict__ scal, double *__restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double r = 0;
    for (int s = 0; s < split; s++) r += rpart[(size_t)s * N + n];
    double corr = 0;
    for (int t = 0; t < n_ctiles; t++) corr += cpart[(size_t)t * N + n];
    const double unit_e = scal[S_UNITE];
    out[n] = (unit_e == 0 ? 0.0 : unit_e * r) + scal[S_H] - corr;
}

// ------------------------------------------------------------------------------------------------
// sparse lists of missing genotypes
// by variant: one warp per variant, ascending sample order
__global__ void fill_mv_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M, int64_t N,
                               const int64_t *__restrict__ mv_ptr, int32_t *__restrict__ mv_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= M) return;
    const uint32_t *row = reinterpret_cast<const uint32_t *>(packed + (size_t)j * pitch);
    const int64_t n_words = (N + 15) / 16;
    int64_t pos = mv_ptr[j];
    for (int64_t w0 = 0; w0 < n_words; w0 += 32) {
        const int64_t wi = w0 + lane;
        uint32_t m = 0;
        if (wi < n_words) {
            const uint32_t w = row[wi];
            m = w & (w >> 1) & 0x55555555u;
            const int64_t rem = N - wi * 16;               // samples of this word that are < N
            if (rem < 16) m &= (rem <= 0) ? 0u : ((1u << (2 * rem)) - 1u);
        }
        const int cnt = __popc(m);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int64_t p = pos + incl - cnt;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            mv_idx[p++] = (int32_t)(wi * 16 + (bit >> 1));
        }
        pos += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// by sample: thread <-> one packed byte (4 samples); pass 0 counts, pass 1 fills in ascending variant order
__global__ void miss_by_sample_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M, int64_t N, int64_t NB,
                                      int32_t *__restrict__ counts, const int64_t *__restrict__ ms_ptr,
                                      int32_t *__restrict__ ms_idx, int fill) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= NB) return;
    int cnt[4] = {0, 0, 0, 0};
    int64_t pos[4] = {0, 0, 0, 0};
    if (fill)
        for (int k = 0; k < 4; k++) if (q * 4 + k < N) pos[k] = ms_ptr[q * 4 + k];
    for (int64_t j = 0; j < M; j++) {
        const uint32_t v = packed[(size_t)j * pitch + q];
        uint32_t m = v & (v >> 1) & 0x55u;
        while (m) {
            const int k = (__ffs(m) - 1) >> 1;
            m &= m - 1;
            if (q * 4 + k < N) {
                if (fill) ms_idx[pos[k]++] = (int32_t)j; else cnt[k]++;
            }
        }
    }
    if (!fill)
        for (int k = 0; k < 4; k++) if (q * 4 + k < N) counts[q * 4 + k] = cnt[k];
}

#include "grm_fused.cuh"
#include "grm_umma.cuh"

int pick_split(int64_t tiles, int slots, int max_split) {
    int best = 1;
    double best_eff = 0;
    for (int s = 1; s <= max_split; s++) {
        const int64_t t = tiles * s;
        const double eff = (double)t / (double)(((t + slots - 1) / slots) * slots);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }
    }
    return best;
}

// missing-genotype sums: rows = variants gathering b (U_j) or rows = samples gathering hm (corr_n)
void launch_sparse(Context &c, ImmaPlan *p, bool by_variant, const double *vec, cudaStream_t st, int grid, bool beside_imma) {
    const int64_t R = by_variant ? c.M : c.N, C = by_variant ? c.N : c.M;
    const int nt = by_variant ? p->n_stiles : p->n_vtiles;
    double *part = by_variant ? p->upart.get() : p->cpart.get();
    // beside the two-pass tensor-core kernels the row-per-thread kernel is used: its CTAs (88 KB) fit next to two IMMA CTAs on an
    // SM, the lane-interleaved kernel (one 180 KB CTA per SM) would serialise with them; on its own the latter is 1.6x faster
    if (p->use_csr || beside_imma) {
        sparse_tile_sum_kernel<<<grid, kSpThreads, kSpSmem, st>>>((by_variant ? p->mv_pos : p->ms_pos).get(),
                                                                  (by_variant ? p->mv_i16 : p->ms_i16).get(), vec, R, C, nt, part);
    } else {
        sparse_ell_sum_kernel<<<c.sm_count, kEllThreads, kEllSmem, st>>>((by_variant ? p->mv_gstart : p->ms_gstart).get(),
                                                                  (by_variant ? p->mv_ell : p->ms_ell).get(), vec, R, C, 2 * ((R + 63) / 64), nt,
                                                                  part);
    }
    SGB_CHECK_LAUNCH();
}


// ---- batched K-RHS product: host side (grm_umma.cuh) -----------------------------------------------------------------------
CUresult encode_tmap_2d(CUtensorMap *tm, const void *base, uint64_t inner_bytes, uint64_t rows, uint64_t stride_bytes, uint32_t box_inner,
                        uint32_t box_rows, CUtensorMapSwizzle swz) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
    static EncodeFn fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SGB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
        if (!f || qres != cudaDriverEntryPointSuccess) throw Error(SGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
        fn = (EncodeFn)f;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)inner_bytes, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)stride_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

// out[c][r] = sum over tiles of part[c][t][r] (fixed order).  grid (ceil(R / 256), ncols)
__global__ void sum_tiles_multi_kernel(const double *__restrict__ part, int n_tiles, int64_t R, double *__restrict__ out, int64_t ldo) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const int c = blockIdx.y;
    double s = 0;
    for (int t = 0; t < n_tiles; t++) s += part[((size_t)c * n_tiles + t) * R + r];
    out[(size_t)c * ldo + r] = s;
}

// split of the contraction range: fewest estimated "box times" (a CTA costs ~2 boxes of fixed overhead), int32 accumulators safe
int umma_pick_split(int64_t row_blocks, int boxes, int sms) {
    int best = 1;
    double best_t = 1e300;
    for (int s = 1; s <= 32 && s <= boxes; s++) {
        const int per = (boxes + s - 1) / s;
        const int64_t units = row_blocks * ((boxes + per - 1) / per);
        const double t = (double)((units + sms - 1) / sms) * (per + 2.0);
        if (t < best_t * 0.97) { best_t = t; best = s; }
    }
    while ((int64_t)((boxes + best - 1) / best) * kUBoxElems * 192 >= (1ll << 31)) best++;
    return best;
}

void umma_prepare(Context &c, ImmaPlan *p) {
    ImmaPlan::Umma &u = p->um;
    if (u.ready || u.failed) return;
    const int64_t M = c.M, N = c.N;
    try {
        u.pitch_t = (size_t)(((M + 3) / 4 + 255) / 256) * 256;
        u.cpad_a = (int64_t)c.pitch * 4;
        u.cpad_b = (int64_t)u.pitch_t * 4;
        u.boxes_a = (int)((N + kUBoxElems - 1) / kUBoxElems);
        u.boxes_b = (int)((M + kUBoxElems - 1) / kUBoxElems);
        u.pt.ensure((size_t)N * u.pitch_t);
        transpose_2bit_kernel<<<dim3((unsigned)(c.pitch / 128), (unsigned)((M + 127) / 128)), 512, 0, c.stream>>>(c.packed.get(), c.pitch, M, N,
                                                                                                        u.pt.get(), u.pitch_t);
        SGB_CHECK_LAUNCH();
        c.stats.n_kernel_launches++;
        u.db.ensure((size_t)kUMaxN * u.cpad_a);
        u.de.ensure((size_t)kUMaxN * u.cpad_b);
        SGB_CUDA(cudaMemsetAsync(u.db.get(), 0, (size_t)kUMaxN * u.cpad_a, c.stream));
        SGB_CUDA(cudaMemsetAsync(u.de.get(), 0, (size_t)kUMaxN * u.cpad_b, c.stream));
        u.t_lo.ensure((size_t)kUMaxCols * M); u.t_hi.ensure((size_t)kUMaxCols * M);
        u.r_lo.ensure((size_t)kUMaxCols * N); u.r_hi.ensure((size_t)kUMaxCols * N);
        u.e.ensure((size_t)kUMaxCols * M); u.hm.ensure((size_t)kUMaxCols * M); u.u.ensure((size_t)kUMaxCols * M);
        u.corr.ensure((size_t)kUMaxCols * N);
        u.part.ensure((size_t)4 * std::max((size_t)p->n_stiles * M, (size_t)p->n_vtiles * N));
        u.vt.ensure((size_t)32 * (std::max(M, N) + 1));
        u.scal.ensure((size_t)kUMaxCols * kUScal);
        u.red.ensure((size_t)kUMaxCols * 2 * 1024);
        u.counter.ensure(kUMaxCols);
        u.err.ensure(1);
        u.herr.ensure(1);
        *u.herr.p = 0;
        SGB_CUDA(cudaMemsetAsync(u.counter.get(), 0, sizeof(unsigned int) * kUMaxCols, c.stream));
        SGB_CUDA(cudaMemsetAsync(u.err.get(), 0, sizeof(int), c.stream));
        CUresult r = encode_tmap_2d(&u.tmap_p, c.packed.get(), c.pitch, (uint64_t)M, c.pitch, kUBoxBytes, kURows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r == CUDA_SUCCESS)
            r = encode_tmap_2d(&u.tmap_pt, u.pt.get(), u.pitch_t, (uint64_t)N, u.pitch_t, kUBoxBytes, kURows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r == CUDA_SUCCESS) r = encode_tmap_2d(&u.tmap_p128, c.packed.get(), c.pitch, (uint64_t)M, c.pitch, kUBoxBytes, kPRows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r == CUDA_SUCCESS) r = encode_tmap_2d(&u.tmap_pt128, u.pt.get(), u.pitch_t, (uint64_t)N, u.pitch_t, kUBoxBytes, kPRows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != CUDA_SUCCESS) throw Error(SGB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        SGB_CUDA(cudaFuncSetAttribute(umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(sparse_ell_multi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, em_smem<4>()));
        SGB_CUDA(cudaFuncSetAttribute(sparse_ell_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, em_smem<2>()));
        u.split_a = umma_pick_split((M + kURows - 1) / kURows, u.boxes_a, c.sm_count);
        u.split_b = umma_pick_split((N + kURows - 1) / kURows, u.boxes_b, c.sm_count);
        u.split_a2 = umma_pick_split((M + kURows - 1) / kURows, u.boxes_a, c.sm_count / 2);
        u.split_b2 = umma_pick_split((N + kURows - 1) / kURows, u.boxes_b, c.sm_count / 2);
        if (const char *e = getenv("SGB_UMMA_SPLIT_A")) u.split_a = u.split_a2 = std::max(1, atoi(e));
        if (const char *e = getenv("SGB_UMMA_SPLIT_B")) u.split_b = u.split_b2 = std::max(1, atoi(e));
        c.sync();
        u.ready = true;
    } catch (const Error &e) {
        // e.g. no room for the second orientation of the packed matrix: multi-column products stay on the single-RHS kernels
        cudaGetLastError();
        u.failed = true;
        u.pt.release(); u.db.release(); u.de.release(); u.part.release();
        c.printf("note: batched tensor-core product unavailable (%s); using the single-RHS kernels per column\n", e.what());
    }
}

// multi-column missing-genotype sums into out[c][row] (rows = variants gathering b, or samples gathering hm)
void launch_sparse_multi(Context &c, ImmaPlan *p, bool by_variant, const double *vec, int64_t ldv, int ncols, double *out, int64_t ldo,
                         cudaStream_t st) {
    ImmaPlan::Umma &u = p->um;
    const int64_t R = by_variant ? c.M : c.N, Cn = by_variant ? c.N : c.M;
    if (ncols > p->um_gather_cols) {
        // many columns: transpose the block once, then one coalesced W * 8 byte load per entry out of L2
        const int64_t *ptr = (by_variant ? p->mv_ptr : p->ms_ptr).get();
        const int32_t *idx = (by_variant ? p->mv_idx : p->ms_idx).get();
        const int grid = c.sm_count * 8;
        c.prof_begin();
        // passes of W columns.  Measured at N = 430K, M = 100K (profiles/r02_gather_variants.txt), both gathers of one product:
        // 30 columns: 2 passes of 16 columns 9.1 ms, 1 pass of 32 columns 10.6 ms, 1 pass of 32 columns with 16-byte loads 7.8 ms;
        // 12 columns: 16 columns 4.35 ms (16-byte loads 4.45), 32 columns 6.8-9.2 ms
        const int W = p->um_gather_w ? p->um_gather_w : (ncols > 16 ? 32 : 16);
        const bool v2 = p->um_gather_v2 >= 0 ? p->um_gather_v2 != 0 : W == 32;
        for (int c0 = 0; c0 < ncols; c0 += W) {
            const int nc = std::min(W, ncols - c0);
            const double *v0 = vec + (size_t)c0 * ldv;
            double *o0 = out + (size_t)c0 * ldo;
            if (W == 32) {
                transpose_cols_kernel<32><<<(unsigned)((Cn + 256) / 256), 256, 0, st>>>(v0, ldv, nc, Cn, u.vt.get());
                // four loads in flight per lane: the gather is bound by the L2's random-row throughput (~14.5 TB/s), not by latency
                // (4 / 8 / 16 in flight: 7.43 / 7.59 / 7.72 ms for both gathers), and fewer registers keep more warps resident
                if (v2) sparse_rows_gather2_kernel<32, 4><<<grid, 256, 0, st>>>(ptr, idx, u.vt.get(), Cn, nc, R, o0, ldo);
                else sparse_rows_gather_kernel<32><<<grid, 256, 0, st>>>(ptr, idx, u.vt.get(), Cn, nc, R, o0, ldo);
            } else if (W == 16) {
                transpose_cols_kernel<16><<<(unsigned)((Cn + 256) / 256), 256, 0, st>>>(v0, ldv, nc, Cn, u.vt.get());
                if (v2) sparse_rows_gather2_kernel<16><<<grid, 256, 0, st>>>(ptr, idx, u.vt.get(), Cn, nc, R, o0, ldo);
                else sparse_rows_gather_kernel<16><<<grid, 256, 0, st>>>(ptr, idx, u.vt.get(), Cn, nc, R, o0, ldo);
            } else {
                transpose_cols_kernel<8><<<(unsigned)((Cn + 256) / 256), 256, 0, st>>>(v0, ldv, nc, Cn, u.vt.get());
                if (v2) sparse_rows_gather2_kernel<8><<<grid, 256, 0, st>>>(ptr, idx, u.vt.get(), Cn, nc, R, o0, ldo);
                else sparse_rows_gather_kernel<8><<<grid, 256, 0, st>>>(ptr, idx, u.vt.get(), Cn, nc, R, o0, ldo);
            }
            c.stats.n_kernel_launches += 2;
        }
        SGB_CHECK_LAUNCH();
        c.prof_end(by_variant ? "sparse_rows_gather_kernel (U)" : "sparse_rows_gather_kernel (corr)");
        return;
    }
    const int nt = by_variant ? p->n_stiles : p->n_vtiles;
    const int64_t G = 2 * ((R + 63) / 64);
    const int64_t *gs = (by_variant ? p->mv_gstart : p->ms_gstart).get();
    const uint16_t *ell = (by_variant ? p->mv_ell : p->ms_ell).get();
    for (int c0 = 0; c0 < ncols; c0 += 4) {
        const int nc = std::min(4, ncols - c0);
        c.prof_begin();
        if (nc > 2)
            sparse_ell_multi_kernel<4><<<c.sm_count, kEmThreads, em_smem<4>(), st>>>(gs, ell, vec + (size_t)c0 * ldv, ldv, nc, R, Cn, G, nt,
                                                                                           u.part.get());
        else
            sparse_ell_multi_kernel<2><<<c.sm_count, kEmThreads, em_smem<2>(), st>>>(gs, ell, vec + (size_t)c0 * ldv, ldv, nc, R, Cn, G, nt,
                                                                                           u.part.get());
        SGB_CHECK_LAUNCH();
        c.prof_end(by_variant ? "sparse_ell_multi_kernel (U)" : "sparse_ell_multi_kernel (corr)");
        c.prof_begin();
        sum_tiles_multi_kernel<<<dim3((unsigned)((R + 255) / 256), nc), 256, 0, st>>>(u.part.get(), nt, R, out + (size_t)c0 * ldo, ldo);
        SGB_CHECK_LAUNCH();
        c.prof_end("sum_tiles_multi_kernel");
        c.stats.n_kernel_launches += 2;
    }
}

// digit-tile maps by box height (ng rows for the single-CTA kernel, ng / 2 for the pair kernel: a multiple of 16 either way)
const CUtensorMap *umma_digit_map(ImmaPlan::Umma &u, bool phase_b, int box_rows) {
    const int i = box_rows / 16;
    CUtensorMap *tm = phase_b ? &u.tmap_de[i] : &u.tmap_db[i];
    bool &have = u.have_d[i];
    if (!have) {
        CUresult r = encode_tmap_2d(&u.tmap_db[i], u.db.get(), (uint64_t)u.cpad_a, kUMaxN, (uint64_t)u.cpad_a, 128, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r == CUDA_SUCCESS)
            r = encode_tmap_2d(&u.tmap_de[i], u.de.get(), (uint64_t)u.cpad_b, kUMaxN, (uint64_t)u.cpad_b, 128, (uint32_t)box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r != CUDA_SUCCESS) throw Error(SGB_ERR_CUDA, "cuTensorMapEncodeTiled (digit tiles) failed (" + std::to_string((int)r) + ")");
        have = true;
    }
    return tm;
}

// launch of one GEMM phase: clusters of two CTAs along x (cta_group::2 MMAs) or plain CTAs
void umma_launch(Context &c, ImmaPlan *p, bool pair, bool prof, int64_t rows, int splits, const CUtensorMap &tp, const CUtensorMap &td, const UmmaArgs &a) {
    if (p->um_pair == 2) {          // pair kernel: 128 rows per CTA, clusters of two (static __cluster_dims__)
        const unsigned gx2 = ((unsigned)((rows + kPRows - 1) / kPRows) + 1) & ~1u;
        umma_pair_kernel<<<dim3(gx2, (unsigned)splits), kPThreads, kPSmemBytes, c.stream>>>(tp, td, a);
        SGB_CHECK_LAUNCH();
        return;
    }
    unsigned gx = (unsigned)((rows + kURows - 1) / kURows);
    if (pair) gx = (gx + 1) & ~1u;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, (unsigned)splits);
    cfg.blockDim = dim3(kUThreads);
    cfg.dynamicSmemBytes = kUSmemBytes;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pair ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (pair) {
        if (prof) SGB_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<true, true>, tp, td, a));
        else SGB_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<false, true>, tp, td, a));
    } else {
        if (prof) SGB_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<true, false>, tp, td, a));
        else SGB_CUDA(cudaLaunchKernelEx(&cfg, umma_gemm_kernel<false, false>, tp, td, a));
    }
    (void)p;
}

// few right-hand sides: the integer GEMM of one phase on mma.sync (imma_small_gemm_kernel); same operands and limb outputs as umma_launch
void small_gemm_launch(Context &c, ImmaPlan *p, const uint8_t *P, size_t pitch, int64_t R, int64_t C, const int8_t *D, int64_t cpad, int ncols,
                       unsigned long long *out_lo, unsigned long long *out_hi, int64_t ldo) {
    bool &attr_set = p->sg_attr_set;          // per plan = per context / device (function attributes are per device)
    int &per_sm = p->sg_per_sm;
    if (!attr_set) {
        SGB_CUDA(cudaFuncSetAttribute(imma_small_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(imma_small_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgSmemBytes));
        SGB_CUDA(cudaFuncSetAttribute(imma_small_gemm_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgSmemBytes));
        attr_set = true;
    }
    const int64_t ksteps = (C + 255) / 256;
    const int64_t row_ctas = (R + kSgRows - 1) / kSgRows;
    // split of the contraction range: all CTAs cost the same, so what counts is the number of waves (two CTAs per SM) times the length
    // of a CTA (its K-steps + ~8 K-steps of pipeline fill and epilogue); at least what keeps the int32 accumulators safe
    if (!per_sm) {
        SGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, imma_small_gemm_kernel<3>, kSgThreads, kSgSmemBytes));
        per_sm = std::max(per_sm, 1);
    }
    const int64_t slots = (int64_t)per_sm * c.sm_count;
    const int s_min = (int)((ksteps + kSgMaxSteps - 1) / kSgMaxSteps);
    int split = s_min;
    double best = 1e300;
    for (int sp = s_min; sp <= 32 && sp <= std::max<int64_t>(s_min, ksteps / 8); sp++) {
        const int64_t ctas = row_ctas * sp, waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * ((double)((ksteps + sp - 1) / sp) + 8.0);
        if (cost < best * 0.99) { best = cost; split = sp; }
    }
    const dim3 grid((unsigned)row_ctas, (unsigned)split);
    const int nt = (kUND * ncols + 7) / 8;
    if (nt == 1) imma_small_gemm_kernel<1><<<grid, kSgThreads, kSgSmemBytes, c.stream>>>(P, pitch, R, ksteps, split, D, cpad, ncols, out_lo, out_hi, ldo);
    else if (nt == 2) imma_small_gemm_kernel<2><<<grid, kSgThreads, kSgSmemBytes, c.stream>>>(P, pitch, R, ksteps, split, D, cpad, ncols, out_lo, out_hi, ldo);
    else imma_small_gemm_kernel<3><<<grid, kSgThreads, kSgSmemBytes, c.stream>>>(P, pitch, R, ksteps, split, D, cpad, ncols, out_lo, out_hi, ldo);
    SGB_CHECK_LAUNCH();
}

// one pass: ncols <= 32 columns
void umma_grm_mv_pass(Context &c, ImmaPlan *p, const double *b, double *out, int ncols) {
    ImmaPlan::Umma &u = p->um;
    const int64_t M = c.M, N = c.N;
    const bool pair = p->um_pair != 0;
    const int ng = pair ? ((kUND * ncols + 31) / 32) * 32 : ((kUND * ncols + 15) / 16) * 16;     // pair: each CTA holds ng / 2 rows of a digit tile
    // The sparse corrections can run on the side stream beside the tensor-core GEMMs (env SGB_UMMA_FORK=1): U beside phase A, corr
    // beside phase B.  Measured on the B200: no gain (19.6 vs 19.2 ms at K = 30) -- next to a GEMM CTA (640 threads, 46 K registers)
    // only one 8-warp gather block fits per SM, and the gather lives on having ~64 warps x 16 loads in flight -- so the default is
    // serial on the main stream.
    cudaStream_t side = (c.profiling || !p->um_fork) ? c.stream : p->side;
    const bool fork = side != c.stream;
    if (fork) {
        SGB_CUDA(cudaEventRecord(p->ev_in, c.stream));
        SGB_CUDA(cudaStreamWaitEvent(side, p->ev_in, 0));
    }
    launch_sparse_multi(c, p, true, b, N, ncols, u.u.get(), M, side);
    if (fork) SGB_CUDA(cudaEventRecord(p->ev_u, side));
    c.prof_begin();
    umma_colstats_kernel<<<ncols, 1024, 0, c.stream>>>(b, N, u.scal.get());
    SGB_CHECK_LAUNCH();
    umma_digits_kernel<<<dim3((unsigned)((u.cpad_a + 255) / 256), ncols), 256, 0, c.stream>>>(b, N, N, u.cpad_a, u.scal.get(), 0, u.db.get());
    SGB_CHECK_LAUNCH();
    SGB_CUDA(cudaMemsetAsync(u.t_lo.get(), 0, sizeof(unsigned long long) * (size_t)ncols * M, c.stream));
    SGB_CUDA(cudaMemsetAsync(u.t_hi.get(), 0, sizeof(unsigned long long) * (size_t)ncols * M, c.stream));
    SGB_CUDA(cudaMemsetAsync(u.r_lo.get(), 0, sizeof(unsigned long long) * (size_t)ncols * N, c.stream));
    SGB_CUDA(cudaMemsetAsync(u.r_hi.get(), 0, sizeof(unsigned long long) * (size_t)ncols * N, c.stream));
    c.prof_end("umma_prep_b (colstats+digits+memsets)");
    UmmaArgs a;
    a.ncols = ncols; a.ng = ng; a.err = u.err.get();
    // phase A: rows = variants, contraction over samples
    const int sa = p->um_pair == 2 ? u.split_a2 : u.split_a, sb = p->um_pair == 2 ? u.split_b2 : u.split_b;
    a.R = M; a.boxes_total = u.boxes_a; a.boxes_per_split = (u.boxes_a + sa - 1) / sa;
    a.out_lo = u.t_lo.get(); a.out_hi = u.t_hi.get(); a.ldo = M;
    const int ns_a = (u.boxes_a + a.boxes_per_split - 1) / a.boxes_per_split;
    c.prof_begin();
    const bool prof = getenv("SGB_UMMA_PROF") != nullptr;
    a.prof = nullptr;
    if (prof) {
        u.prof.ensure((size_t)16 * 65536);
        SGB_CUDA(cudaMemsetAsync(u.prof.get(), 0, sizeof(long long) * 16 * 65536, c.stream));
        a.prof = u.prof.get();
    }
    const bool small = ncols <= p->um_small_cols && kUND * ncols <= 8 * kSgNT;      // few columns: mma.sync GEMM (a tcgen05.mma costs its dispatch whatever N)
    if (small) small_gemm_launch(c, p, c.packed.get(), c.pitch, M, N, u.db.get(), u.cpad_a, ncols, u.t_lo.get(), u.t_hi.get(), M);
    else umma_launch(c, p, pair, prof, M, ns_a, p->um_pair == 2 ? u.tmap_p128 : u.tmap_p, *umma_digit_map(u, false, pair ? ng / 2 : ng), a);
    if (prof) {
        // debugging aid: cycle counters of one issuer thread and of one expander warp of the leader CTAs, per OWN stage
        std::vector<long long> h((size_t)16 * 65536);
        c.d2h(h.data(), u.prof.get(), sizeof(long long) * h.size());
        c.sync();
        double s[16] = {0};
        for (size_t i = 0; i < 65536; i++) for (int k = 0; k < 16; k++) s[k] += (double)h[i * 16 + k];
        const double ks = s[5] > 0 ? s[5] : 1, es = s[14] > 0 ? s[14] : ks;
        c.printf("umma prof (clk per own stage): issuer total %.0f = wait_a %.0f + mma %.0f + commit %.0f | expander total %.0f = "
                 "wait_p %.0f + load/expand %.0f + wait_done %.0f + st/wait::st %.0f + wait_b/arrive %.0f\n", s[0] / ks, s[2] / ks, s[3] / ks,
                 s[4] / ks, s[8] / es, s[9] / es, s[10] / es, s[11] / es, s[12] / es, s[13] / es);
    }
    SGB_CHECK_LAUNCH();
    c.prof_end(small ? "imma_small_gemm_kernel (phase A)" : "umma_gemm_kernel (phase A)");
    if (fork) SGB_CUDA(cudaStreamWaitEvent(c.stream, p->ev_u, 0));
    c.prof_begin();
    const int Gm = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (M + 255) / 256));
    umma_finalize_kernel<<<dim3(Gm, ncols), 256, 0, c.stream>>>(u.t_lo.get(), u.t_hi.get(), u.u.get(), 1, c.lut.get(), M, 1.0 / (double)c.M_total,
                                                                u.e.get(), u.hm.get(), u.red.get(), u.counter.get(), u.scal.get());
    SGB_CHECK_LAUNCH();
    if (fork) {
        SGB_CUDA(cudaEventRecord(p->ev_hm, c.stream));
        SGB_CUDA(cudaStreamWaitEvent(side, p->ev_hm, 0));
    }
    umma_digits_kernel<<<dim3((unsigned)((u.cpad_b + 255) / 256), ncols), 256, 0, c.stream>>>(u.e.get(), M, M, u.cpad_b, u.scal.get(), 1, u.de.get());
    SGB_CHECK_LAUNCH();
    c.prof_end("umma_finalize+digits_e");
    launch_sparse_multi(c, p, false, u.hm.get(), M, ncols, u.corr.get(), N, side);
    if (fork) SGB_CUDA(cudaEventRecord(p->ev_corr, side));
    // phase B: rows = samples, contraction over variants (sample-major copy)
    a.R = N; a.boxes_total = u.boxes_b; a.boxes_per_split = (u.boxes_b + sb - 1) / sb;
    a.out_lo = u.r_lo.get(); a.out_hi = u.r_hi.get(); a.ldo = N;
    const int ns_b = (u.boxes_b + a.boxes_per_split - 1) / a.boxes_per_split;
    c.prof_begin();
    if (small) small_gemm_launch(c, p, u.pt.get(), u.pitch_t, N, M, u.de.get(), u.cpad_b, ncols, u.r_lo.get(), u.r_hi.get(), N);
    else umma_launch(c, p, pair, false, N, ns_b, p->um_pair == 2 ? u.tmap_pt128 : u.tmap_pt, *umma_digit_map(u, true, pair ? ng / 2 : ng), a);
    SGB_CHECK_LAUNCH();
    c.prof_end(small ? "imma_small_gemm_kernel (phase B)" : "umma_gemm_kernel (phase B)");
    if (fork) SGB_CUDA(cudaStreamWaitEvent(c.stream, p->ev_corr, 0));
    c.prof_begin();
    umma_combine_kernel<<<dim3((unsigned)((N + 255) / 256), ncols), 256, 0, c.stream>>>(u.r_lo.get(), u.r_hi.get(), N, u.corr.get(), 1, u.scal.get(), out);
    SGB_CHECK_LAUNCH();
    c.prof_end("umma_combine_kernel");
    c.stats.n_kernel_launches += 7;
    c.stats.n_product_launches += 1;
}

}  // namespace

// ---- class sums of a packed genotype block against fixed model columns: the dense part of the score scan (score.cu) -----------------
// The score statistics of a variant are sums of model columns over the samples of each genotype class, so the scan over a block of
// variants is three integer GEMMs on the pair kernel above -- A = low bit, high bit and (low & high) of the 2-bit codes, B = the digit
// planes of the model columns, quantised once per model.
static_assert(kClassMaxCols == kUMaxCols && kClassDigitRows == kUMaxN && kClassScal == kUScal, "ctx.h mirrors grm_umma.cuh");
void umma_class_digits(Context &c, const double *cols, int64_t n, int ncols, int64_t cpad, int8_t *digits, double *scal, long long *tot) {
    if (ncols < 1 || ncols > kUMaxCols) throw Error(SGB_ERR_INVALID, "umma_class_digits: 1..32 columns per group");
    umma_colstats_kernel<<<ncols, 1024, 0, c.stream>>>(cols, n, scal);
    SGB_CHECK_LAUNCH();
    SGB_CUDA(cudaMemsetAsync(digits, 0, (size_t)kUMaxN * cpad, c.stream));
    umma_digits_kernel<<<dim3((unsigned)((cpad + 255) / 256), ncols), 256, 0, c.stream>>>(cols, n, n, cpad, scal, 0, digits);
    SGB_CHECK_LAUNCH();
    umma_digit_totals_kernel<<<ncols, 1024, 0, c.stream>>>(cols, n, n, scal, tot);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches += 3;
}

// out_lo / out_hi [ncols][rows] (zeroed here) += limbs of sum_i A(code(row, i)) * digits(col, i); amode as in UmmaArgs
void umma_class_sums(Context &c, const uint8_t *packed, size_t pitch, int64_t rows, int64_t n, const int8_t *digits, int64_t cpad, int ncols,
                     int amode, unsigned long long *out_lo, unsigned long long *out_hi, int *err_dev) {
    if (pitch % 128 != 0 || (int64_t)pitch * 4 > cpad || ((uintptr_t)packed & 127) != 0)
        throw Error(SGB_ERR_INVALID, "umma_class_sums: the packed block needs a 128-byte aligned base and pitch");
    // per call: function attributes are per device, and a process may hold contexts on several devices (microseconds next to the kernel)
    SGB_CUDA(cudaFuncSetAttribute(umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPSmemBytes));
    const int ng = ((kUND * ncols + 31) / 32) * 32;
    CUtensorMap tp, td;
    CUresult r = encode_tmap_2d(&tp, packed, pitch, (uint64_t)rows, pitch, kUBoxBytes, kPRows, CU_TENSOR_MAP_SWIZZLE_128B);
    if (r == CUDA_SUCCESS) r = encode_tmap_2d(&td, digits, (uint64_t)cpad, kUMaxN, (uint64_t)cpad, 128, (uint32_t)(ng / 2), CU_TENSOR_MAP_SWIZZLE_128B);
    if (r != CUDA_SUCCESS) throw Error(SGB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    const int boxes = (int)((n + kUBoxElems - 1) / kUBoxElems);
    const int splits = umma_pick_split((rows + kURows - 1) / kURows, boxes, c.sm_count / 2);
    UmmaArgs a;
    a.ncols = ncols; a.ng = ng; a.R = rows; a.boxes_total = boxes; a.boxes_per_split = (boxes + splits - 1) / splits;
    a.out_lo = out_lo; a.out_hi = out_hi; a.ldo = rows; a.err = err_dev; a.prof = nullptr; a.amode = amode;
    const int ns = (boxes + a.boxes_per_split - 1) / a.boxes_per_split;
    SGB_CUDA(cudaMemsetAsync(out_lo, 0, sizeof(unsigned long long) * (size_t)ncols * rows, c.stream));
    SGB_CUDA(cudaMemsetAsync(out_hi, 0, sizeof(unsigned long long) * (size_t)ncols * rows, c.stream));
    const unsigned gx2 = ((unsigned)((rows + kPRows - 1) / kPRows) + 1) & ~1u;
    umma_pair_kernel<<<dim3(gx2, (unsigned)ns), kPThreads, kPSmemBytes, c.stream>>>(tp, td, a);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

bool imma_available(const Context &c) { return c.imma != nullptr; }

void imma_release(Context &c) {
    c.async_err = nullptr;   // points into the plan's pinned flag
    c.async_err_dev = nullptr;
    delete c.imma;
    c.imma = nullptr;
}

void imma_prepare(Context &c) {
    imma_release(c);
    ImmaPlan *p = new ImmaPlan();
    try {
        const int64_t M = c.M, N = c.N;
        p->ksteps = (N + 255) / 256;
        p->kblocks = (M + 31) / 32;
        if ((size_t)p->ksteps * 64 > c.pitch) throw Error(SGB_ERR_STATE, "row pitch too small for the IMMA kernels");
        // ---- missing lists by variant (counts are known from the store pass: N - cnt_num) ----
        std::vector<int64_t> ptr(M + 1, 0);
        for (int64_t j = 0; j < M; j++) ptr[j + 1] = ptr[j] + (N - c.h_cnt_num[j]);
        p->nnz = ptr[M];
        p->mv_ptr.ensure(M + 1);
        p->mv_idx.ensure(std::max<int64_t>(p->nnz, 1));
        c.h2d(p->mv_ptr.get(), ptr.data(), sizeof(int64_t) * (M + 1));
        fill_mv_kernel<<<(unsigned)((M + 7) / 8), 256, 0, c.stream>>>(c.packed.get(), c.pitch, M, N, p->mv_ptr.get(),
                                                                     p->mv_idx.get());
        SGB_CHECK_LAUNCH();
        c.sync();
        // ---- by sample ----
        DevBuf<int32_t> counts;
        counts.ensure(N);
        p->ms_ptr.ensure(N + 1);
        p->ms_idx.ensure(std::max<int64_t>(p->nnz, 1));
        miss_by_sample_kernel<<<(unsigned)((c.NB + 127) / 128), 128, 0, c.stream>>>(c.packed.get(), c.pitch, M, N, c.NB,
                                                                                   counts.get(), nullptr, nullptr, 0);
        SGB_CHECK_LAUNCH();
        std::vector<int32_t> hc(N);
        c.d2h(hc.data(), counts.get(), sizeof(int32_t) * N);
        c.sync();
        std::vector<int64_t> sp(N + 1, 0);
        for (int64_t n = 0; n < N; n++) sp[n + 1] = sp[n] + hc[n];
        if (sp[N] != p->nnz) throw Error(SGB_ERR_STATE, "missing-genotype lists disagree");
        c.h2d(p->ms_ptr.get(), sp.data(), sizeof(int64_t) * (N + 1));
        miss_by_sample_kernel<<<(unsigned)((c.NB + 127) / 128), 128, 0, c.stream>>>(c.packed.get(), c.pitch, M, N, c.NB,
                                                                                   nullptr, p->ms_ptr.get(), p->ms_idx.get(), 1);
        SGB_CHECK_LAUNCH();
        c.sync();
        c.stats.n_kernel_launches += 3;
        // ---- work buffers ----
        const int slots = 2 * c.sm_count;
        p->split_a = pick_split((M + kAVar - 1) / kAVar, slots, 8);
        p->split_b = pick_split((N + kBSamp - 1) / kBSamp, slots, 8);
        p->dfrag.ensure((size_t)p->ksteps * 2048);
        p->efrag.ensure((size_t)p->kblocks * 256);
        p->tq.ensure((size_t)p->split_a * M);
        p->e.ensure(M); p->hm.ensure(M);
        p->n_stiles = (int)((N + kSpTile - 1) / kSpTile);
        p->n_vtiles = (int)((M + kSpTile - 1) / kSpTile);
        p->upart.ensure((size_t)p->n_stiles * M);
        p->cpart.ensure((size_t)p->n_vtiles * N);
        auto tile_major = [&](const DevBuf<int64_t> &ptr, const DevBuf<int32_t> &idx, int64_t R, int nt, DevBuf<int64_t> &pos,
                              DevBuf<uint16_t> &i16) {
            const size_t cells = (size_t)nt * R;
            pos.ensure(cells + 1);
            i16.ensure((size_t)p->nnz + 16);   // +16: the staged copy rounds the range up to 16-byte granules
            SGB_CUDA(cudaMemsetAsync(pos.get() + cells, 0, sizeof(int64_t), c.stream));
            sparse_tile_count_kernel<<<dim3((unsigned)((R + 255) / 256), nt), 256, 0, c.stream>>>(ptr.get(), idx.get(), R, pos.get());
            SGB_CHECK_LAUNCH();
            thrust::exclusive_scan(thrust::cuda::par.on(c.stream), pos.get(), pos.get() + cells + 1, pos.get());
            sparse_tile_fill_kernel<<<dim3((unsigned)((R + 255) / 256), nt), 256, 0, c.stream>>>(ptr.get(), idx.get(), R, pos.get(),
                                                                                               i16.get());
            SGB_CHECK_LAUNCH();
        };
        tile_major(p->mv_ptr, p->mv_idx, M, p->n_stiles, p->mv_pos, p->mv_i16);
        tile_major(p->ms_ptr, p->ms_idx, N, p->n_vtiles, p->ms_pos, p->ms_i16);
        c.sync();
        auto to_ell = [&](const DevBuf<int64_t> &pos, const DevBuf<uint16_t> &i16, int64_t R, int nt, DevBuf<int64_t> &gstart,
                          DevBuf<uint16_t> &ell) {
            const int64_t G = 2 * ((R + 63) / 64), cells = G * nt;   // even: 16-byte aligned group-start rows
            gstart.ensure((size_t)cells + 4);
            SGB_CUDA(cudaMemsetAsync(gstart.get() + cells, 0, sizeof(int64_t), c.stream));
            ell_len_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, c.stream>>>(pos.get(), R, G, nt, gstart.get());
            SGB_CHECK_LAUNCH();
            thrust::exclusive_scan(thrust::cuda::par.on(c.stream), gstart.get(), gstart.get() + cells + 1, gstart.get());
            int64_t total = 0;
            c.d2h(&total, gstart.get() + cells, sizeof(int64_t));
            c.sync();
            ell.ensure((size_t)total + 64);
            ell_fill_kernel<<<(unsigned)((cells + 7) / 8), 256, 0, c.stream>>>(pos.get(), i16.get(), R, G, nt, gstart.get(), ell.get());
            SGB_CHECK_LAUNCH();
        };
        to_ell(p->mv_pos, p->mv_i16, M, p->n_stiles, p->mv_gstart, p->mv_ell);
        to_ell(p->ms_pos, p->ms_i16, N, p->n_vtiles, p->ms_gstart, p->ms_ell);
        c.sync();
        p->use_csr = getenv("SGB_SPARSE_CSR") != nullptr;   // comparison switch: the row-per-thread kernel everywhere
        SGB_CUDA(cudaFuncSetAttribute(sparse_ell_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kEllSmem));
        // the row-major 32-bit lists are only needed to build the tile-major ones
        // (kept: the row-gather kernel of the batched product walks them)
        SGB_CUDA(cudaFuncSetAttribute(sparse_tile_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpSmem));
        p->rpart.ensure((size_t)p->split_b * N);
        {
            int lo_prio = 0, hi_prio = 0;
            SGB_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
            SGB_CUDA(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, hi_prio));
        }
        for (cudaEvent_t *e : {&p->ev_in, &p->ev_u, &p->ev_hm, &p->ev_corr}) SGB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        p->scal.ensure(16);
        p->red.ensure(8 * 1024);
        p->counter.ensure(8);
        SGB_CUDA(cudaMemsetAsync(p->counter.get(), 0, sizeof(unsigned int) * 8, c.stream));
        SGB_CUDA(cudaMemsetAsync(p->scal.get(), 0, sizeof(double) * 16, c.stream));
        SGB_CUDA(cudaFuncSetAttribute(imma_dots_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * kAStageBytes));
        if (4 * kAStageBytes <= 227 * 1024)
            SGB_CUDA(cudaFuncSetAttribute(imma_dots_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * kAStageBytes));
        {
            const int64_t hsteps = 2 * p->ksteps;
            const int ks_per = (int)((hsteps + c.sm_count - 1) / c.sm_count);      // half-steps (128 samples) per CTA slice
            int coop = 0;
            SGB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c.dev));
            const char *fe = getenv("SGB_FUSED");
            if (coop && ks_per >= 1 && ks_per <= 2 * kFMaxKs && !(fe && atoi(fe) == 0)) {
                SGB_CUDA(cudaFuncSetAttribute(imma_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFSmemBytes));
                int per_sm = 0;
                SGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, imma_fused_kernel, kFThreads, kFSmemBytes));
                p->f_ks_per_cta = ks_per;
                p->f_grid = (int)((hsteps + ks_per - 1) / ks_per);
                p->f_tiles = (M + kFV - 1) / kFV;
                if (per_sm >= 1 && p->f_grid <= per_sm * c.sm_count) {
                    p->dfrag128.ensure((size_t)p->ksteps * 2048);
                    p->f_acc_stride = ((p->f_tiles + 15) / 16) * 16 + 16;   // odd multiple of 128 B between limbs
                    p->f_acc.ensure((size_t)p->f_acc_stride * kFV * 2);
                    p->f_rout.ensure(N);
                    p->f_u.ensure(M);
                    p->f_htotal.ensure(1);
                    p->f_hpart.ensure((size_t)p->f_grid);
                    p->f_edig.ensure((size_t)p->f_tiles * 32 + 128);
                    p->f_err.ensure(1);
                    p->f_herr.ensure(1);
                    *p->f_herr.p = 0;
                    SGB_CUDA(cudaMemsetAsync(p->f_err.get(), 0, sizeof(int), c.stream));
                    // bound on |e_j| / |b|_2 (Cauchy-Schwarz).  sum_n lut_j[c_nj]^2 = n0 l0^2 + n1 l1^2 + n2 l2^2 grows with n2
                    // at fixed (n_valid, allele sum), so n2 = floor(sum / 2) gives a rigorous upper bound (<= 2x the HWE value).
                    std::vector<double> hl((size_t)4 * M);
                    c.d2h(hl.data(), c.lut.get(), sizeof(double) * 4 * M);
                    c.sync();
                    double ef = 0;
                    for (int64_t j = 0; j < M; j++) {
                        const double num = c.h_cnt_num[j], sum = c.h_cnt_sum[j];
                        const double n2 = std::floor(sum / 2), n1 = sum - 2 * n2, n0 = std::max(0.0, num - n1 - n2);
                        const double l0 = hl[4 * j], l1 = hl[4 * j + 1], l2 = hl[4 * j + 2];
                        const double ss = n0 * l0 * l0 + n1 * l1 * l1 + n2 * l2 * l2;
                        ef = std::max(ef, std::sqrt(ss) * std::fabs(l1 - l0));
                    }
                    p->f_efactor = ef / (double)c.M_total * (1.0 + 1e-9);
                    if (!(p->f_efactor > 0) || !std::isfinite(p->f_efactor)) p->f_efactor = 1e-300;   // all-monomorphic shard: e == 0
                    if (const char *e = getenv("SGB_FUSED_POLL_NS")) p->f_poll_ns = std::max(0, atoi(e));
                    if (const char *e = getenv("SGB_FUSED_LAG")) p->f_lag = atoi(e);
                    p->f_lag = std::min(kFNBuf - 2, std::max(1, p->f_lag));
                    {
                        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
                        void *fn = nullptr;
                        cudaDriverEntryPointQueryResult qres;
                        SGB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
                        if (!fn || qres != cudaDriverEntryPointSuccess) throw Error(SGB_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
                        const cuuint64_t gdim[2] = {(cuuint64_t)c.pitch, (cuuint64_t)M};
                        const cuuint64_t gstride[1] = {(cuuint64_t)c.pitch};
                        const cuuint32_t box[2] = {128, (cuuint32_t)kFV};
                        const cuuint32_t estr[2] = {1, 1};
                        CUresult r = ((EncodeFn)fn)(&p->f_tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)c.packed.get(), gdim, gstride, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                        if (r != CUDA_SUCCESS) throw Error(SGB_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
                    }
                    p->fused_ok = true;
                }
            }
        }
        if (const char *e = getenv("SGB_WAIT_TIMEOUT_MS")) {
            const unsigned long long ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
            SGB_CUDA(cudaMemcpyToSymbolAsync(g_wait_timeout_ns, &ns, sizeof(ns), 0, cudaMemcpyHostToDevice, c.stream));
        }
        if (const char *e = getenv("SGB_UMMA_MIN_COLS")) { p->um_min_cols = atoi(e); if (p->um_min_cols <= 0) p->um_min_cols = INT_MAX; }
        if (const char *e = getenv("SGB_UMMA_FORK")) p->um_fork = atoi(e);
        if (const char *e = getenv("SGB_UMMA_PAIR")) p->um_pair = atoi(e);
        if (const char *e = getenv("SGB_UMMA_SMALL_COLS")) p->um_small_cols = atoi(e);
        if (const char *e = getenv("SGB_UMMA_GATHER_COLS")) p->um_gather_cols = atoi(e);
        if (const char *e = getenv("SGB_UMMA_GATHER_V2")) p->um_gather_v2 = atoi(e) != 0 ? 1 : 0;
        if (const char *e = getenv("SGB_UMMA_GATHER_W")) { const int w = atoi(e); if (w == 8 || w == 16 || w == 32) p->um_gather_w = w; }
        if (const char *e = getenv("SGB_SPARSE_FORK")) p->opt_fork = atoi(e);
        if (const char *e = getenv("SGB_FUSED_FORK")) p->opt_fork_fused = atoi(e);
        if (const char *e = getenv("SGB_SPARSE_GRID_MULT")) p->opt_grid_mult = std::max(1, atoi(e));
        if (const char *e = getenv("SGB_DOTS_STAGES")) p->opt_stages = (atoi(e) == 4) ? 4 : 3;
        SGB_CUDA(cudaFuncSetAttribute(imma_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBStages * kBStageBytes));
        c.sync();
    } catch (...) {
        delete p;
        throw;
    }
    c.imma = p;
}

void imma_grm_mv(Context &c, const double *b_all, double *out_all, int k) {
    ImmaPlan *p = c.imma;
    if (!p) throw Error(SGB_ERR_STATE, "the IMMA product kernel is not prepared");
    const int64_t M = c.M, N = c.N;
    const int G = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (N + 2047) / 2048));
    const int Gm = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (M + 255) / 256));
    // In profiling mode everything runs serially on the main stream so that each kernel can be timed.
    cudaStream_t side = (c.profiling || !p->opt_fork) ? c.stream : p->side;
    const int sp_grid = c.sm_count * p->opt_grid_mult;
    // The fused kernel pays ~1 us of cross-CTA latency per 32-variant tile whatever the slice width, so it only beats the two
    // HBM passes when every CTA's slice is (nearly) full: SGB_KERNEL_AUTO takes it from 10 of the 12 K-steps per CTA upwards
    // (N >= ~380K on 148 SMs) and the two-pass kernels otherwise; SGB_KERNEL_IMMA forces it whenever the shape allows.
    const bool use_fused = p->fused_ok && !c.fused_disabled && (c.kernel == SGB_KERNEL_IMMA || (c.kernel == SGB_KERNEL_AUTO && p->f_ks_per_cta >= 20));
    // several right-hand sides: one pass over the packed matrix for all of them on tcgen05 (grm_umma.cuh)
    if ((c.kernel == SGB_KERNEL_UMMA || (c.kernel == SGB_KERNEL_AUTO && k >= p->um_min_cols)) && !p->um.failed && !c.fused_disabled) {
        umma_prepare(c, p);
        if (p->um.ready) {
            c.async_err = p->um.herr.p;
            c.async_err_dev = p->um.err.get();
            // balanced passes of at most 32 columns
            const int n_pass = (k + kUMaxCols - 1) / kUMaxCols;
            int done = 0;
            for (int ps = 0; ps < n_pass; ps++) {
                const int nc = (k - done + (n_pass - ps) - 1) / (n_pass - ps);
                umma_grm_mv_pass(c, p, b_all + (size_t)done * N, out_all + (size_t)done * N, nc);
                done += nc;
            }
            SGB_CUDA(cudaMemcpyAsync(p->um.herr.p, p->um.err.get(), sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            return;
        }
    }
    if (use_fused) {
        c.async_err = p->f_herr.p;
        c.async_err_dev = p->f_err.get();
        for (int col = 0; col < k; col++) {
            const double *b = b_all + (size_t)col * N;
            double *out = out_all + (size_t)col * N;
            // the small preparation kernels (max|b|, |b|_2, digits of b, zeroing of the limbs) need no shared memory and run on
            // the side stream beside the sparse U_j kernel
            // (single right-hand side only: with many columns per call -- the fits -- the cross-stream dependencies were measured
            //  to cost more than the ~0.03 ms they hide; env SGB_FUSED_FORK = 0 / 1 forces it off / on)
            const bool fork = side != c.stream && (p->opt_fork_fused == 1 || (p->opt_fork_fused < 0 && k == 1));
            if (fork) {
                SGB_CUDA(cudaEventRecord(p->ev_in, c.stream));
                SGB_CUDA(cudaStreamWaitEvent(side, p->ev_in, 0));
            }
            c.prof_begin();
            cudaStream_t pst = fork ? side : c.stream;
            absmax_sum_kernel<<<G, 256, 0, pst>>>(b, N, p->red.get(), p->counter.get(), p->scal.get(), p->f_efactor);
            SGB_CHECK_LAUNCH();
            digits_b128_kernel<<<(unsigned)((p->ksteps * 256 + 255) / 256), 256, 0, pst>>>(b, N, p->ksteps * 256, p->scal.get(),
                                                                                          p->dfrag128.get());
            SGB_CHECK_LAUNCH();
            SGB_CUDA(cudaMemsetAsync(p->f_acc.get(), 0, sizeof(unsigned long long) * p->f_acc_stride * kFV * 2, pst));
            SGB_CUDA(cudaMemsetAsync(p->f_edig.get(), 0, sizeof(unsigned long long) * (p->f_tiles * 32 + 128), pst));
            c.prof_end("imma_prep_b (absmax+digits+memset)");
            if (fork) SGB_CUDA(cudaEventRecord(p->ev_u, side));
            c.prof_begin();
            launch_sparse(c, p, true, b, c.stream, sp_grid, false);
            c.prof_end("sparse_tile_sum_kernel (U_j)");
            c.prof_begin();
            sum_tiles_kernel<<<(unsigned)((M + 255) / 256), 256, 0, c.stream>>>(p->upart.get(), p->n_stiles, M, p->f_u.get());
            SGB_CHECK_LAUNCH();
            c.prof_end("sum_tiles_kernel");
            if (fork) SGB_CUDA(cudaStreamWaitEvent(c.stream, p->ev_u, 0));
            FusedArgs fa;
            fa.packed = c.packed.get(); fa.pitch = c.pitch; fa.M = M; fa.N = N; fa.hsteps = 2 * p->ksteps;
            fa.hs_per_cta = p->f_ks_per_cta; fa.n_tiles = p->f_tiles; fa.dfrag128 = p->dfrag128.get();
            fa.acc_t = p->f_acc.get(); fa.acc_stride = p->f_acc_stride; fa.u = p->f_u.get(); fa.poll_ns = p->f_poll_ns; fa.lag = p->f_lag;
            fa.lut = c.lut.get(); fa.inv_mtotal = 1.0 / (double)c.M_total; fa.scal = p->scal.get(); fa.hm = p->hm.get();
            fa.h_part = p->f_hpart.get(); fa.edig = p->f_edig.get(); fa.rout = p->f_rout.get(); fa.err = p->f_err.get();
            void *kargs[] = {&p->f_tmap, &fa};
            c.prof_begin();
            SGB_CUDA(cudaLaunchCooperativeKernel((const void *)imma_fused_kernel, dim3(p->f_grid), dim3(kFThreads), kargs,
                                                 (size_t)kFSmemBytes, c.stream));
            c.prof_end("imma_fused_kernel");
            SGB_CUDA(cudaMemcpyAsync(p->f_herr.p, p->f_err.get(), sizeof(int), cudaMemcpyDeviceToHost, c.stream));
            c.prof_begin();
            launch_sparse(c, p, false, p->hm.get(), c.stream, sp_grid, false);
            c.prof_end("sparse_tile_sum_kernel (corr_n)");
            c.prof_begin();
            CombineSrc cs;
            cs.rout = p->f_rout.get(); cs.cpart = p->cpart.get(); cs.n_ctiles = p->n_vtiles; cs.h_part = p->f_hpart.get(); cs.n_hpart = p->f_grid;
            if (k == 1 && comm_combine_allreduce(c, cs, out)) {
                // several ranks on one node: the last addition and the sum over the ranks are one kernel over peer memory (comm.cu)
                c.product_reduced = true;
                c.prof_end("peer_allreduce_kernel (combine + sum over ranks)");
            } else {
                combine_fused_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c.stream>>>(p->f_rout.get(), N, p->cpart.get(), p->n_vtiles,
                                                                                      p->f_hpart.get(), p->f_grid, out);
                SGB_CHECK_LAUNCH();
                c.prof_end("combine_kernel");
            }
            c.stats.n_kernel_launches += 7;
            c.stats.n_product_launches += 1;
        }
        return;
    }
    const bool fork = side != c.stream;
    for (int col = 0; col < k; col++) {
        const double *b = b_all + (size_t)col * N;
        double *out = out_all + (size_t)col * N;
        if (fork) {
            SGB_CUDA(cudaEventRecord(p->ev_in, c.stream));
            SGB_CUDA(cudaStreamWaitEvent(side, p->ev_in, 0));
        }
        c.prof_begin();
        launch_sparse(c, p, true, b, side, sp_grid, fork);
        c.prof_end("sparse_tile_sum_kernel (U_j)");
        if (fork) SGB_CUDA(cudaEventRecord(p->ev_u, side));
        c.prof_begin();
        absmax_sum_kernel<<<G, 256, 0, c.stream>>>(b, N, p->red.get(), p->counter.get(), p->scal.get(), 0.0);
        SGB_CHECK_LAUNCH();
        digits_b_kernel<<<(unsigned)((p->ksteps * 256 + 255) / 256), 256, 0, c.stream>>>(b, N, p->ksteps * 256, p->scal.get(),
                                                                                       p->dfrag.get());
        SGB_CHECK_LAUNCH();
        c.prof_end("imma_prep_b (absmax+digits)");
        c.prof_begin();
        if (p->opt_stages == 3)
            imma_dots_kernel<3><<<dim3((unsigned)((M + kAVar - 1) / kAVar), p->split_a), kAThreads, 3 * kAStageBytes, c.stream>>>(
                c.packed.get(), c.pitch, M, p->ksteps, p->split_a, p->dfrag.get(), p->tq.get());
        else
            imma_dots_kernel<4><<<dim3((unsigned)((M + kAVar - 1) / kAVar), p->split_a), kAThreads, 4 * kAStageBytes, c.stream>>>(
                c.packed.get(), c.pitch, M, p->ksteps, p->split_a, p->dfrag.get(), p->tq.get());
        SGB_CHECK_LAUNCH();
        c.prof_end("imma_dots_kernel");
        if (fork) SGB_CUDA(cudaStreamWaitEvent(c.stream, p->ev_u, 0));
        c.prof_begin();
        finalize_dots_kernel<<<Gm, 256, 0, c.stream>>>(p->tq.get(), p->split_a, p->upart.get(), p->n_stiles, c.lut.get(), M,
                                                       1.0 / (double)c.M_total, p->e.get(), p->hm.get(), p->red.get() + 4096,
                                                       p->counter.get() + 1, p->scal.get());
        SGB_CHECK_LAUNCH();
        if (fork) {
            SGB_CUDA(cudaEventRecord(p->ev_hm, c.stream));
            SGB_CUDA(cudaStreamWaitEvent(side, p->ev_hm, 0));
        }
        digits_e_kernel<<<(unsigned)((p->kblocks * 32 + 255) / 256), 256, 0, c.stream>>>(p->e.get(), M, p->kblocks * 32,
                                                                                      p->scal.get(), p->efrag.get());
        SGB_CHECK_LAUNCH();
        c.prof_end("imma_finalize+digits_e");
        c.prof_begin();
        launch_sparse(c, p, false, p->hm.get(), side, sp_grid, fork);
        c.prof_end("sparse_tile_sum_kernel (corr_n)");
        if (fork) SGB_CUDA(cudaEventRecord(p->ev_corr, side));
        c.prof_begin();
        imma_apply_kernel<<<dim3((unsigned)((N + kBSamp - 1) / kBSamp), p->split_b), kBThreads, kBStages * kBStageBytes,
                            c.stream>>>(c.packed.get(), c.pitch, M, N, p->kblocks, p->split_b, p->efrag.get(), p->rpart.get());
        SGB_CHECK_LAUNCH();
        c.prof_end("imma_apply_kernel");
        if (fork) SGB_CUDA(cudaStreamWaitEvent(c.stream, p->ev_corr, 0));
        c.prof_begin();
        combine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, c.stream>>>(p->rpart.get(), p->split_b, N, p->cpart.get(),
                                                                        p->n_vtiles, p->scal.get(), out);
        SGB_CHECK_LAUNCH();
        c.prof_end("combine_kernel");
        c.stats.n_kernel_launches += 9;
        c.stats.n_product_launches += 1;
    }
}

}  // namespace sgb
