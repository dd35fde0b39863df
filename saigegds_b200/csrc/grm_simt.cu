// FP64 CUDA-core implementation of get_crossprod_b_grm (saige_fitnull.cpp:435-536, dense branch).
//
// Two sample-tiled kernels per product, both deterministic (fixed summation order, no FP atomics):
//   dots : thread <-> variant.  A CTA stages a [128 variants x 2048 samples] packed tile and the
//          matching slice of b in shared memory; each thread sweeps its variant's row with b broadcast
//          from shared memory.  Per genotype: two predicated FP64 adds (bit 0 -> S_lo, bit 1 -> S_hi);
//          dot = l0*S0 + l1*S1 + l2*S2 with S1 = S_lo - S_both, S2 = S_hi - S_both, S0 = sum(b) - S1 - S2 - S_both.
//          (code 3 = missing has standardised value 0, saige_fitnull.cpp:199.)
//   apply: thread <-> 8 samples, accumulators in registers for the whole variant sweep; per variant a
//          4-entry table tab_j[code] = dot_j * lut_j[code] is read from shared memory.
// The same apply kernel with tab = lut^2 gives diag(GRM) (saige_fitnull.cpp:205-227).
// This is the reference-shaped GPU path; the int8 tensor-core path in grm_imma.cu is the fast one.
#include "ctx.h"

namespace sgb {

namespace {

constexpr int kDotVar = 128;     // variants per CTA (== threads)
constexpr int kDotWords = 128;   // 32-bit words per row chunk = 2048 samples
constexpr int kDotSamples = kDotWords * 16;
constexpr int kTilePitch = kDotWords + 1;  // odd word pitch: conflict-free row sweeps

__global__ void __launch_bounds__(kDotVar) simt_dots_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M,
                                                            int64_t N, const double *__restrict__ lut,
                                                            const double *__restrict__ b, double *__restrict__ partial) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    double *sb = reinterpret_cast<double *>(smem_raw);
    uint32_t *tile = reinterpret_cast<uint32_t *>(smem_raw + sizeof(double) * kDotSamples);
    __shared__ double s_warp_sum[kDotVar / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t chunk = blockIdx.y, j0 = (int64_t)blockIdx.x * kDotVar;
    const int64_t n0 = chunk * kDotSamples;
    double local = 0;
    for (int i = tid; i < kDotSamples; i += kDotVar) {
        double v = (n0 + i < N) ? b[n0 + i] : 0.0;
        sb[i] = v;
        local += v;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (lane == 0) s_warp_sum[warp] = local;
    const size_t byte0 = (size_t)chunk * (kDotWords * 4);
    for (int r = warp; r < kDotVar; r += kDotVar / 32) {
        const int64_t j = j0 + r;
        uint4 v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        const size_t off = byte0 + (size_t)lane * 16;
        if (j < M && off + 16 <= pitch) v = *reinterpret_cast<const uint4 *>(packed + (size_t)j * pitch + off);
        uint32_t *dst = tile + r * kTilePitch + lane * 4;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    const int64_t j = j0 + tid;
    if (j >= M) return;
    double sb_total = 0;
#pragma unroll
    for (int w = 0; w < kDotVar / 32; w++) sb_total += s_warp_sum[w];
    double s_lo = 0, s_hi = 0, s_both = 0;
    const uint32_t *row = tile + tid * kTilePitch;
    for (int i = 0; i < kDotWords; i++) {
        const uint32_t w = row[i];
        const double *bp = sb + i * 16;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const double bv = bp[s];
            if (w & (1u << (2 * s))) s_lo += bv;
            if (w & (2u << (2 * s))) s_hi += bv;
        }
        uint32_t m = w & (w >> 1) & 0x55555555u;
        while (m) {
            int pos = __ffs(m) - 1;
            m &= m - 1;
            s_both += bp[pos >> 1];
        }
    }
    const double S1 = s_lo - s_both, S2 = s_hi - s_both, S0 = sb_total - S1 - S2 - s_both;
    const double *l = lut + 4 * j;
    partial[(size_t)chunk * M + j] = l[0] * S0 + l[1] * S1 + l[2] * S2;
}

// dot_j = sum over sample chunks (fixed order); tab_j[k] = dot_j * lut_j[k]
__global__ void simt_dots_finalize_kernel(const double *__restrict__ partial, int n_chunk, int64_t M,
                                          const double *__restrict__ lut, double *__restrict__ tab) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    double d = 0;
    for (int c = 0; c < n_chunk; c++) d += partial[(size_t)c * M + j];
    const double *l = lut + 4 * j;
    double *t = tab + 4 * j;
    t[0] = d * l[0]; t[1] = d * l[1]; t[2] = d * l[2]; t[3] = 0.0;
}

constexpr int kApplyThreads = 128;
constexpr int kApplyVarTile = 64;

// out[n] = scale * sum_j tab_j[code(n, j)];  thread <-> 2 packed bytes (8 samples)
__global__ void __launch_bounds__(kApplyThreads) simt_apply_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M,
                                                                   int64_t N, const double *__restrict__ tab,
                                                                   double *__restrict__ out, double scale) {
    __shared__ double stab[kApplyVarTile * 4];
    const int tid = threadIdx.x;
    const size_t byte0 = ((size_t)blockIdx.x * kApplyThreads + tid) * 2;
    const bool in_range = byte0 + 2 <= pitch;
    double acc[8];
#pragma unroll
    for (int s = 0; s < 8; s++) acc[s] = 0;
    for (int64_t j0 = 0; j0 < M; j0 += kApplyVarTile) {
        __syncthreads();
        for (int i = tid; i < kApplyVarTile * 4; i += kApplyThreads) {
            int64_t idx = j0 * 4 + i;
            stab[i] = (idx < 4 * M) ? tab[idx] : 0.0;
        }
        __syncthreads();
        const int nv = (int)min((int64_t)kApplyVarTile, M - j0);
        const uint8_t *base = packed + (size_t)j0 * pitch + byte0;
        for (int jj = 0; jj < nv; jj += 8) {
            uint32_t w[8];
#pragma unroll
            for (int u = 0; u < 8; u++)
                w[u] = (in_range && jj + u < nv) ? *reinterpret_cast<const uint16_t *>(base + (size_t)(jj + u) * pitch) : 0xFFFFu;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const double *t = stab + (jj + u) * 4;   // rows past nv hold zeros or are masked by code 3 -> t[3]
                if (jj + u < nv) {
#pragma unroll
                    for (int s = 0; s < 8; s++) acc[s] += t[(w[u] >> (2 * s)) & 3u];
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < 8; s++) {
        int64_t n = (int64_t)byte0 * 4 + s;
        if (n < N) out[n] = acc[s] * scale;
    }
}

}  // namespace

void simt_table_apply(Context &c, const double *tab_device, double *out_device, double scale) {
    const unsigned grid = (unsigned)((c.pitch / 2 + kApplyThreads - 1) / kApplyThreads);
    c.prof_begin();
    simt_apply_kernel<<<grid, kApplyThreads, 0, c.stream>>>(c.packed.get(), c.pitch, c.M, c.N, tab_device, out_device, scale);
    SGB_CHECK_LAUNCH();
    c.prof_end("simt_apply_kernel");
    c.stats.n_kernel_launches++;
}

void simt_grm_mv(Context &c, const double *b_device, double *out_device) {
    const int n_chunk = (int)((c.N + kDotSamples - 1) / kDotSamples);
    c.ws_partial.ensure((size_t)n_chunk * c.M);
    c.ws_tab.ensure((size_t)4 * c.M);
    const size_t smem = sizeof(double) * kDotSamples + sizeof(uint32_t) * kDotVar * kTilePitch;
    SGB_CUDA(cudaFuncSetAttribute(simt_dots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((c.M + kDotVar - 1) / kDotVar), (unsigned)n_chunk);
    c.prof_begin();
    simt_dots_kernel<<<grid, kDotVar, smem, c.stream>>>(c.packed.get(), c.pitch, c.M, c.N, c.lut.get(), b_device,
                                                        c.ws_partial.get());
    SGB_CHECK_LAUNCH();
    c.prof_end("simt_dots_kernel");
    c.prof_begin();
    simt_dots_finalize_kernel<<<(unsigned)((c.M + 255) / 256), 256, 0, c.stream>>>(c.ws_partial.get(), n_chunk, c.M,
                                                                                 c.lut.get(), c.ws_tab.get());
    SGB_CHECK_LAUNCH();
    c.prof_end("simt_dots_finalize_kernel");
    c.stats.n_kernel_launches += 2;
    simt_table_apply(c, c.ws_tab.get(), out_device, 1.0 / (double)c.M_total);
}

}  // namespace sgb
