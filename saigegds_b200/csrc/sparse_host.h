// Host side of the sparse-genotype entry points (the reference's default, geno.sparse=TRUE):
//   saige_get_sparse     saige_fitnull.cpp:252-320   one variant -> (n1, n2, n3, indices of 1s, of 2s, of missing)
//   saige_store_sp_geno  saige_fitnull.cpp:324-388   list of such vectors -> look-up table, diag(GRM), product state
// The GPU keeps ONE genotype layout (2-bit packed): the index lists are turned into packed rows on
// the host, slab by slab, and go through the same device store as saige_store_2b_geno.  A variant's packed row is 16x
// smaller than its index list at MAF 0.25, so packing before the PCIe copy is also the cheaper order.  Pad samples of the
// last byte are written as code 3 (missing): then the dense allele counts over whole bytes equal the sparse ones
// (n_valid = N - n3, sum = n1 + 2 n2, :349-350) and the look-up table is bit-identical.
#pragma once
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace sgb {

// geno_type: 0 = bytes (RAWSXP), 1 = int32 (INTSXP), 2 = double (REALSXP); out holds n_samp + 3 ints.
inline int64_t get_sparse_host(const void *geno, int geno_type, int64_t n_samp, int32_t *out) {
    std::vector<uint8_t> gs((size_t)n_samp);
    if (geno_type == 0) {
        memcpy(gs.data(), geno, (size_t)n_samp);
    } else if (geno_type == 1) {
        const int32_t *s = (const int32_t *)geno;
        for (int64_t i = 0; i < n_samp; i++) gs[i] = (0 <= s[i] && s[i] <= 2) ? (uint8_t)s[i] : 3;
    } else {
        const double *s = (const double *)geno;
        for (int64_t i = 0; i < n_samp; i++) {
            uint8_t c = 3;
            if (std::isfinite(s[i])) {
                const double r = std::round(s[i]);
                if (r >= 0 && r <= 2) c = (uint8_t)r;
            }
            gs[i] = c;
        }
    }
    int64_t n = 0, sum = 0;
    for (int64_t i = 0; i < n_samp; i++)
        if (gs[i] < 3) { sum += gs[i]; n++; }
    const bool flip = sum > n;   // alternative allele is the major one: count the other allele (:298-303)
    // raw bytes above 3 are neither genotypes nor missing in the reference (:297, :310-316): they stay out of all three
    // runs and so read as genotype 0
    int32_t cnt[4] = {0, 0, 0, 0};
    for (int64_t i = 0; i < n_samp; i++) {
        if (flip && gs[i] < 3) gs[i] = 2 - gs[i];
        if (gs[i] <= 3) cnt[gs[i]]++;
    }
    int32_t *p1 = out + 3, *p2 = p1 + cnt[1], *p3 = p2 + cnt[2];
    for (int64_t i = 0; i < n_samp; i++) {
        const uint8_t c = gs[i];
        if (c == 1) *p1++ = (int32_t)i;
        else if (c == 2) *p2++ = (int32_t)i;
        else if (c == 3) *p3++ = (int32_t)i;
    }
    out[0] = cnt[1]; out[1] = cnt[2]; out[2] = cnt[3];
    return 3 + (int64_t)cnt[1] + cnt[2] + cnt[3];
}

// One variant's index vector -> one packed row of nb = ceil(n_samp/4) bytes.  Returns an error text or NULL.
inline const char *sparse_row_to_packed(const int32_t *pg, int64_t len, int64_t n_samp, int64_t nb, uint8_t *row) {
    if (len < 3) return "a sparse genotype vector needs at least the three counts";
    const int64_t n1 = pg[0], n2 = pg[1], n3 = pg[2];
    if (n1 < 0 || n2 < 0 || n3 < 0 || 3 + n1 + n2 + n3 != len) return "sparse genotype vector: counts do not match its length";
    memset(row, 0, (size_t)nb);
    const int n_pad = (int)(nb * 4 - n_samp);
    if (n_pad > 0) row[nb - 1] = (uint8_t)(0xFFu << (2 * (4 - n_pad)));
    const int32_t *ii = pg + 3;
    const int64_t runs[3] = {n1, n2, n3};
    for (int code = 1; code <= 3; code++) {
        for (int64_t k = 0; k < runs[code - 1]; k++) {
            const int64_t i = *ii++;
            if (i < 0 || i >= n_samp) return "sparse genotype vector: sample index out of range";
            row[i >> 2] |= (uint8_t)(code << (2 * (i & 3)));
        }
    }
    return nullptr;
}

// Variants [j0, j1) -> packed rows at dst (row pitch nb), split over host threads.  Throws std::string on bad input.
inline void sparse_to_packed_host(const int32_t *data, const int64_t *offsets, int64_t j0, int64_t j1, int64_t n_samp,
                                  uint8_t *dst, int n_thread = 0) {
    const int64_t nb = (n_samp + 3) / 4, m = j1 - j0;
    if (m <= 0) return;
    if (n_thread <= 0) n_thread = (int)std::thread::hardware_concurrency();
    if (n_thread < 1) n_thread = 1;
    if ((int64_t)n_thread > m) n_thread = (int)m;
    std::vector<const char *> err((size_t)n_thread, nullptr);
    auto work = [&](int t) {
        const int64_t a = j0 + m * t / n_thread, b = j0 + m * (t + 1) / n_thread;
        for (int64_t j = a; j < b && !err[t]; j++)
            err[t] = sparse_row_to_packed(data + offsets[j], offsets[j + 1] - offsets[j], n_samp, nb, dst + (size_t)(j - j0) * nb);
    };
    if (n_thread == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_thread; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (const char *e : err)
        if (e) throw std::string(e);
}

}  // namespace sgb
