// Genotype store: device layout, allele counts, standardised look-up table, diag(GRM), decode.
// Replaces saige_store_2b_geno (saige_fitnull.cpp:159-230) and get_geno_ds (:394-427).
//
// HBM layout: one row per variant, `pitch` = ceil(N/4) rounded up to 256 bytes so that every row
// starts 256-byte aligned (uint4 / cp.async / TMA friendly, whole sample tiles).  Sample 4j+k sits in bits 2k..2k+1 of byte j
// (the reference's format, unchanged).  Samples >= N of the last byte and the pitch padding are
// rewritten to code 3 (missing), whose standardised value is 0, so product kernels never need a
// tail branch.  The reference's allele counts sweep the raw pad bits (:188-192), so the counts are
// taken from the raw bytes BEFORE that rewrite -- they are bit-exact with the reference.
#include "ctx.h"

namespace sgb {

namespace {

__device__ __forceinline__ void count_word(uint32_t w, int &n_valid, int &sum) {
    uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
    uint32_t miss = lo & hi;
    n_valid += 16 - __popc(miss);
    sum += __popc(lo & ~miss) + 2 * __popc(hi & ~miss);
}

// One warp per variant.  n_valid/sum: reference semantics (all NB bytes).  cnt_num/cnt_sum: samples < N only.
__global__ void count_lut_kernel(const uint8_t *__restrict__ src, size_t src_pitch, int64_t M, int64_t N, int64_t NB,
                                 int32_t *__restrict__ n_valid_out, int32_t *__restrict__ sum_out,
                                 int32_t *__restrict__ cnt_num, int32_t *__restrict__ cnt_sum,
                                 double *__restrict__ lut) {
    const int lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= M) return;
    const uint8_t *row = src + (size_t)j * src_pitch;
    int n_valid = 0, sum = 0;
    // head bytes up to 4-byte alignment of the address, then words, then tail bytes
    const uintptr_t addr = (uintptr_t)row;
    int64_t head = (int64_t)((4 - (addr & 3)) & 3);
    if (head > NB) head = NB;
    const int64_t n_words = (NB - head) / 4;
    const int64_t tail0 = head + n_words * 4;
    if (lane < head) count_word(0xFFFFFF00u | row[lane], n_valid, sum);
    const uint32_t *wp = (const uint32_t *)(row + head);
    for (int64_t i = lane; i < n_words; i += 32) count_word(wp[i], n_valid, sum);
    if (lane < NB - tail0) count_word(0xFFFFFF00u | row[tail0 + lane], n_valid, sum);
    // pad correction for the "samples < N" counts: codes of samples >= N in the last byte
    int pad_valid = 0, pad_sum = 0;
    const int n_pad = (int)(NB * 4 - N);
    if (lane == 0 && n_pad > 0) {
        uint32_t last = row[NB - 1];
        for (int k = 4 - n_pad; k < 4; k++) {
            uint32_t c = (last >> (2 * k)) & 3u;
            if (c < 3) { pad_valid++; pad_sum += c; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    if (lane == 0) {
        n_valid_out[j] = n_valid;
        sum_out[j] = sum;
        cnt_num[j] = n_valid - pad_valid;
        cnt_sum[j] = sum - pad_sum;
        // saige_fitnull.cpp:193-199
        double af = double(sum) / (2 * n_valid);
        double inv = 1 / sqrt(2 * af * (1 - af));
        if (!isfinite(af) || !isfinite(inv)) af = inv = 0;
        double *p = lut + 4 * j;
        p[0] = (0 - 2 * af) * inv;
        p[1] = (1 - 2 * af) * inv;
        p[2] = (2 - 2 * af) * inv;
        p[3] = 0;
    }
}

// Copy raw rows into the pitched layout; pad samples / pitch padding -> code 3.
__global__ void relayout_kernel(const uint8_t *__restrict__ src, size_t src_pitch, uint8_t *__restrict__ dst,
                                size_t pitch, int64_t M, int64_t N, int64_t NB) {
    const int64_t j = blockIdx.y;
    const int n_pad = (int)(NB * 4 - N);
    const uint8_t pad_mask = n_pad > 0 ? (uint8_t)(0xFFu << (2 * (4 - n_pad))) : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)pitch; i += (int64_t)gridDim.x * blockDim.x) {
        uint8_t v = 0xFF;
        if (i < NB) {
            v = src[(size_t)j * src_pitch + i];
            if (i == NB - 1) v |= pad_mask;
        }
        dst[(size_t)j * pitch + i] = v;
    }
}

__global__ void square_table_kernel(const double *__restrict__ lut, double *__restrict__ tab, int64_t n4) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) { double v = lut[i]; tab[i] = v * v; }
}

__global__ void decode_kernel(const uint8_t *__restrict__ row, int64_t N, double *__restrict__ out) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    uint32_t c = (row[n >> 2] >> (2 * (n & 3))) & 3u;
    out[n] = (c < 3) ? (double)c : __longlong_as_double(0x7ff8000000000000LL);
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// One thread per output byte (4 samples).  Counter-based: depends only on (seed, global variant, sample).
__global__ void synth_kernel(uint8_t *__restrict__ out, int64_t N, int64_t NB, int64_t M, int64_t var_offset, uint64_t seed,
                             double miss) {
    const int64_t j = blockIdx.y;
    const int64_t gj = j + var_offset;
    const double maf = 0.005 + 0.495 * ((mix64(seed ^ (0xA5A5A5A5ULL + (uint64_t)gj * 0x632BE59BD9B4E019ULL)) >> 11) * (1.0 / 9007199254740992.0));
    const double q0 = (1 - maf) * (1 - maf), q1 = q0 + 2 * maf * (1 - maf);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < NB; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t byte = 0;
        for (int k = 0; k < 4; k++) {
            int64_t n = i * 4 + k;
            uint32_t code = 3;
            if (n < N) {
                uint64_t r = mix64(seed + (uint64_t)gj * 0x9E3779B97F4A7C15ULL + (uint64_t)n * 0xD1B54A32D192ED03ULL);
                double u = (double)(r >> 40) * (1.0 / 16777216.0);
                double um = (double)((r >> 8) & 0xFFFFFF) * (1.0 / 16777216.0);
                code = (u < q0) ? 0u : (u < q1 ? 1u : 2u);
                if (um < miss) code = 3u;
            }
            byte |= code << (2 * k);
        }
        out[(size_t)j * NB + i] = (uint8_t)byte;
    }
}

}  // namespace

void store_device_layout(Context &c, const uint8_t *src, size_t src_pitch) {
    const int64_t M = c.M, N = c.N, NB = c.NB;
    c.pitch = (size_t)((NB + 255) / 256) * 256;   // 256-byte multiple: whole 1024-sample tiles for the IMMA kernels
    c.packed.ensure((size_t)M * c.pitch);
    c.lut.ensure((size_t)4 * M);
    c.diag.ensure((size_t)N);
    c.n_valid.ensure(M); c.sum.ensure(M); c.cnt_num.ensure(M); c.cnt_sum.ensure(M);
    {
        const int warps = 8;
        count_lut_kernel<<<(unsigned)((M + warps - 1) / warps), warps * 32, 0, c.stream>>>(
            src, src_pitch, M, N, NB, c.n_valid.get(), c.sum.get(), c.cnt_num.get(), c.cnt_sum.get(), c.lut.get());
        SGB_CHECK_LAUNCH();
    }
    {
        // grid.y is limited to 65535: loop over slabs of variants
        const int64_t slab = 32768;
        for (int64_t j0 = 0; j0 < M; j0 += slab) {
            int64_t mj = std::min<int64_t>(slab, M - j0);
            dim3 grid((unsigned)std::min<int64_t>((c.pitch + 255) / 256, 64), (unsigned)mj);
            relayout_kernel<<<grid, 256, 0, c.stream>>>(src + (size_t)j0 * src_pitch, src_pitch,
                                                        c.packed.get() + (size_t)j0 * c.pitch, c.pitch, mj, N, NB);
            SGB_CHECK_LAUNCH();
        }
    }
    c.stats.n_kernel_launches += 2;
    // diag(GRM): (1/M_total) sum_j lut_j[g_ij]^2  (saige_fitnull.cpp:205-227) == table apply with lut^2
    c.ws_tab.ensure((size_t)4 * M);
    square_table_kernel<<<(unsigned)((4 * M + 255) / 256), 256, 0, c.stream>>>(c.lut.get(), c.ws_tab.get(), 4 * M);
    SGB_CHECK_LAUNCH();
    simt_table_apply(c, c.ws_tab.get(), c.diag.get(), 1.0 / (double)c.M_total);
    if (c.world > 1) comm_allreduce_sum(c, c.diag.get(), (size_t)N);
    c.h_cnt_num.resize(M); c.h_cnt_sum.resize(M);
    c.d2h(c.h_cnt_num.data(), c.cnt_num.get(), sizeof(int32_t) * M);
    c.d2h(c.h_cnt_sum.data(), c.cnt_sum.get(), sizeof(int32_t) * M);
    c.sync();
}

void decode_variant(Context &c, int64_t local_idx, double *out_device) {
    decode_kernel<<<(unsigned)((c.N + 255) / 256), 256, 0, c.stream>>>(c.packed.get() + (size_t)local_idx * c.pitch, c.N,
                                                                        out_device);
    SGB_CHECK_LAUNCH();
    c.stats.n_kernel_launches++;
}

void synth_geno(Context &c, int64_t n_samp, int64_t m_local, int64_t var_offset, uint64_t seed, double miss,
                uint8_t *out_device) {
    const int64_t NB = (n_samp + 3) / 4;
    const int64_t slab = 32768;
    for (int64_t j0 = 0; j0 < m_local; j0 += slab) {
        int64_t mj = std::min<int64_t>(slab, m_local - j0);
        dim3 grid((unsigned)std::min<int64_t>((NB + 255) / 256, 128), (unsigned)mj);
        synth_kernel<<<grid, 256, 0, c.stream>>>(out_device + (size_t)j0 * NB, n_samp, NB, mj, var_offset + j0, seed, miss);
        SGB_CHECK_LAUNCH();
    }
    c.sync();
}

}  // namespace sgb
