// Genotype store: device layout, allele counts, standardised look-up table, diag(GRM), decode.
// Replaces saige_store_2b_geno (saige_fitnull.cpp:159-230) and get_geno_ds (:394-427).
//
// HBM layout: one row per variant, `pitch` = ceil(N/4) rounded up to 256 bytes so that every row
// starts 256-byte aligned (uint4 / cp.async / TMA friendly, whole sample tiles).  Sample 4j+k sits in bits 2k..2k+1 of byte j
// (the reference's format, unchanged).  Samples >= N of the last byte and the pitch padding are
// rewritten to code 3 (missing), whose standardised value is 0, so product kernels never need a
// tail branch.  The reference's allele counts sweep the raw pad bits (:188-192), so the counts are
// taken from the raw bytes BEFORE that rewrite -- they are bit-exact with the reference.
#include <algorithm>

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

// GDS genotype/data (bit2 allele indices, [variant][sample][ploidy 2], no row padding) -> 2-bit alt-allele dosage rows.
// One nibble per sample: allele 1 in bits 0-1, allele 2 in bits 2-3; 0 = reference, 1/2 = an alternative allele,
// 3 = missing.  nib0: nibble index of (first variant of the slab, sample 0) relative to `bits`.  One thread per output
// byte; per-variant allele counts (all integer) by warp reduction + integer atomics, so they do not depend on the order.
__global__ void gds_to_dosage_kernel(const uint8_t *__restrict__ bits, int64_t nib0, int64_t n_file,
                                     const int32_t *__restrict__ sample_sel, int64_t N, int64_t NB, uint8_t *__restrict__ out,
                                     int32_t *__restrict__ n_valid_alleles, int32_t *__restrict__ n_alt_alleles) {
    const int64_t v = blockIdx.y;
    const int64_t base = nib0 + v * n_file;
    int valid = 0, alt = 0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < NB; j += (int64_t)gridDim.x * blockDim.x) {
        uint32_t byte = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int64_t smp = 4 * j + k;
            uint32_t code = 3;
            if (smp < N) {
                const int64_t idx = base + (sample_sel ? (int64_t)sample_sel[smp] : smp);
                const uint32_t nib = (bits[idx >> 1] >> (4 * (int)(idx & 1))) & 15u;
                const uint32_t a0 = nib & 3u, a1 = nib >> 2;
                valid += (a0 != 3u) + (a1 != 3u);
                alt += (a0 == 1u || a0 == 2u) + (a1 == 1u || a1 == 2u);
                code = (a0 == 3u || a1 == 3u) ? 3u : (uint32_t)(a0 != 0u) + (uint32_t)(a1 != 0u);
            }
            byte |= code << (2 * k);
        }
        out[(size_t)v * NB + j] = (uint8_t)byte;
    }
    for (int o = 16; o > 0; o >>= 1) {
        valid += __shfl_xor_sync(0xffffffffu, valid, o);
        alt += __shfl_xor_sync(0xffffffffu, alt, o);
    }
    if ((threadIdx.x & 31) == 0 && (valid | alt)) {
        atomicAdd(&n_valid_alleles[v], valid);
        atomicAdd(&n_alt_alleles[v], alt);
    }
}

// dst row r = src row rows[r]
__global__ void gather_rows_kernel(const uint8_t *__restrict__ src, const int64_t *__restrict__ rows, int64_t NB,
                                   uint8_t *__restrict__ dst) {
    const int64_t r = blockIdx.y, sr = rows[r];
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < NB; j += (int64_t)gridDim.x * blockDim.x)
        dst[(size_t)r * NB + j] = src[(size_t)sr * NB + j];
}

}  // namespace

void gds_to_dosage(Context &c, const uint8_t *bits_host, int64_t n_file, int64_t m_file, const int32_t *sample_sel_host,
                   int64_t N, uint8_t *packed_device, int32_t *n_valid_alleles_device, int32_t *n_alt_alleles_device) {
    const int64_t NB = (N + 3) / 4;
    DevBuf<int32_t> sel;
    if (sample_sel_host) {
        sel.ensure((size_t)N);
        c.h2d(sel.get(), sample_sel_host, sizeof(int32_t) * N);
    }
    SGB_CUDA(cudaMemsetAsync(n_valid_alleles_device, 0, sizeof(int32_t) * m_file, c.stream));
    SGB_CUDA(cudaMemsetAsync(n_alt_alleles_device, 0, sizeof(int32_t) * m_file, c.stream));
    // slabs of an even number of variants (a slab then starts on a byte boundary even when n_file is odd), <= ~256 MB
    int64_t slab = std::max<int64_t>(2, (((int64_t)512 << 20) / std::max<int64_t>(1, n_file)) & ~(int64_t)1);
    slab = std::min<int64_t>(slab, 32768);            // grid.y limit
    DevBuf<uint8_t> stage;
    stage.ensure((size_t)((slab * n_file + 1) / 2 + 1));
    for (int64_t v0 = 0; v0 < m_file; v0 += slab) {
        const int64_t mv = std::min(slab, m_file - v0);
        const int64_t nib_first = v0 * n_file, nib_last = (v0 + mv) * n_file;     // nib_first is even
        const int64_t byte0 = nib_first >> 1, bytes = ((nib_last + 1) >> 1) - byte0;
        c.h2d(stage.get(), bits_host + byte0, (size_t)bytes);
        dim3 grid((unsigned)std::min<int64_t>((NB + 255) / 256, 64), (unsigned)mv);
        gds_to_dosage_kernel<<<grid, 256, 0, c.stream>>>(stage.get(), 0, n_file, sample_sel_host ? sel.get() : nullptr, N, NB,
                                                         packed_device + (size_t)v0 * NB, n_valid_alleles_device + v0,
                                                         n_alt_alleles_device + v0);
        SGB_CHECK_LAUNCH();
        c.stats.n_kernel_launches++;
        c.sync();   // the staging buffer is reused by the next slab
    }
}

void gather_rows(Context &c, const uint8_t *src_device, const int64_t *rows_host, int64_t n_rows, int64_t NB,
                 uint8_t *dst_device) {
    if (n_rows == 0) return;
    DevBuf<int64_t> rows;
    rows.ensure((size_t)n_rows);
    c.h2d(rows.get(), rows_host, sizeof(int64_t) * n_rows);
    const int64_t slab = 32768;
    for (int64_t r0 = 0; r0 < n_rows; r0 += slab) {
        const int64_t mr = std::min(slab, n_rows - r0);
        dim3 grid((unsigned)std::min<int64_t>((NB + 255) / 256, 64), (unsigned)mr);
        gather_rows_kernel<<<grid, 256, 0, c.stream>>>(src_device, rows.get() + r0, NB, dst_device + (size_t)r0 * NB);
        SGB_CHECK_LAUNCH();
        c.stats.n_kernel_launches++;
    }
    c.sync();
}

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
