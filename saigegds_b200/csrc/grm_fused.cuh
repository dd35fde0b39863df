// Fused single-HBM-pass GRM product (included by grm_imma.cu inside its anonymous namespace).
//
// One persistent CTA per SM owns a slice of <= 3072 samples (12 K-steps of 256) for the whole product and walks the
// variants in tiles of 32.  A packed tile [32 variants x 768 B] is fetched ONCE from HBM (cp.async.bulk + mbarrier),
// used for phase A (partial dots over the CTA's samples), kept in shared memory while the partial dots of all CTAs
// are summed through exact 64-bit integer atomics in L2, and then used again for phase B (apply) -- the int32
// accumulators of phase B live in registers for the whole product.
//
//   compute warps 0..7 : step s: wait full[s] -> phase A(tile s) -> shared int32 atomics -> arrive acc_done[s]
//                                wait ready[s-3] -> phase B(tile s-3) -> arrive empty[s-3]
//   loader warp 8      : wait empty -> arm full with expect_tx -> one bulk copy per row (lane <-> variant)
//   service warp 9     : wait acc_done[s] -> 64 global red.add.u64 (two limbs per variant) -> fence -> counter[s]++
//                        spin until counter[s-2] == #CTAs -> dot, e, hm for the tile (lane <-> variant) -> base-128
//                        digits of e in MMA fragment order -> ready[s-2]
//
// All cross-warp synchronisation is mbarrier based; there is no CTA-wide barrier in the main loop.  The digit
// exponent of e is adaptive (the global max |e_j| is not known in advance): every CTA sees the same e_j and takes the
// same decision, so all CTAs switch exponent at the same tile; a switch flushes the int32 accumulators to FP64.
// Everything that crosses CTAs is integer, so the result is independent of timing and bit-reproducible.
// Spin loops are bounded: on time-out an error flag is raised and the kernel drains instead of hanging the GPU.
#pragma once

constexpr int kFV = 32;                 // variants per tile
constexpr int kFComputeWarps = 8;
constexpr int kFThreads = (kFComputeWarps + 2) * 32;
constexpr int kFMaxKs = 12;             // K-steps (256 samples) per CTA slice
constexpr int kFRowBytes = kFMaxKs * 64;          // 768
constexpr int kFPitch = kFRowBytes + 16;          // 784 = 16 (mod 128): conflict-free for both ldmatrix patterns
constexpr int kFTileBytes = kFV * kFPitch;        // 25,088
constexpr int kFNBuf = 7;               // tile ring: 3 ahead of phase A + 3 waiting for phase B + 1
constexpr int kFLag = 3;                // phase B runs 3 tiles behind phase A
constexpr int kFFinLag = 2;             // the service warp finalises tile s-2 at step s
constexpr int kFNE = 4;                 // e-digit ring
constexpr int kFNAcc = 4;               // shared partial-dot ring
constexpr int kFHpw = 3;                // half-steps (128 samples) per compute warp: 24 / 8
constexpr int kFRbw = 6;                // 64-sample row-blocks per compute warp: 48 / 8
constexpr int kFHead = 6;               // head-room bits of the adaptive exponent
constexpr long long kFSpinMax = 1ll << 24;

struct FusedSmem {
    unsigned long long full[kFNBuf], empty[kFNBuf], acc_done[kFNAcc], ready[kFNE];
    int esh[kFNE];                      // digit shift of the tile in this slot (INT_MIN: all e are zero)
    int eflag[kFNE];
    int accum[kFNAcc][kFV * 8];         // partial dots of this CTA: [variant][digit plane]
    unsigned char efrag[kFNE][256];
};
constexpr int kFSmemBytes = kFNBuf * kFTileBytes + (int)sizeof(FusedSmem) + 128;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded wait; returns false on time-out
__device__ __forceinline__ bool mbar_wait(unsigned long long *b, unsigned parity, volatile int *err) {
    const unsigned a = smem_u32(b);
    for (long long it = 0; it < kFSpinMax; it++) {
        unsigned ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) return true;
        if ((it & 1023) == 1023 && *err) return false;
    }
    *err = 1;
    return false;
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

struct FusedArgs {
    const uint8_t *packed; size_t pitch; int64_t M, N, ksteps;
    int ks_per_cta;                     // K-steps per CTA (<= kFMaxKs)
    int64_t n_tiles;
    const int8_t *dfrag128;             // [half-step][lane][t0][h][beta]: digits of b, 1 KB per 128 samples
    unsigned long long *acc_t;          // [n_tiles * 32][2] exact integer T'_j limbs (zeroed before the launch)
    unsigned int *counter;              // [n_tiles] arrivals (zeroed before the launch)
    const double *u;                    // [M] U_j = sum of b over the missing samples of variant j
    const double *lut;
    double inv_mtotal;
    const double *scal;                 // S_UNITB, S_SUMB
    double *hm;                         // [M] out: h_j + 3 e_j for the sparse output correction
    double *h_total;                    // out: H = sum_j h_j
    double *rout;                       // [N] out: sum_j e_j c'_nj in real units (this CTA's slice)
    int *err;
};

__global__ void __launch_bounds__(kFThreads, 1) imma_fused_kernel(FusedArgs A) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t *tiles = smem_raw;
    FusedSmem &S = *reinterpret_cast<FusedSmem *>(smem_raw + kFNBuf * kFTileBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int n_cta = gridDim.x;
    const int64_t ks0 = (int64_t)blockIdx.x * A.ks_per_cta;
    const int ks_n = (int)min((int64_t)A.ks_per_cta, A.ksteps - ks0);        // >= 1 by construction of the grid
    const unsigned row_bytes = (unsigned)min((size_t)kFRowBytes, A.pitch - (size_t)ks0 * 64);
    const int64_t T = A.n_tiles;
    volatile int *err = A.err;

    if (tid == 0) {
        for (int i = 0; i < kFNBuf; i++) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], kFComputeWarps); }
        for (int i = 0; i < kFNAcc; i++) mbar_init(&S.acc_done[i], kFComputeWarps);
        for (int i = 0; i < kFNE; i++) mbar_init(&S.ready[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kFNAcc * kFV * 8; i += kFThreads) (&S.accum[0][0])[i] = 0;
    // the tile buffers may be read before every byte was ever written (short rows of the last CTA): clear once
    for (int i = tid; i < kFNBuf * kFTileBytes / 16; i += kFThreads) reinterpret_cast<uint4 *>(tiles)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    if (warp == kFComputeWarps) {
        // ------------------------------------------------------------------ loader: lane <-> variant row
        for (int64_t t = 0; t < T; t++) {
            const int b = (int)(t % kFNBuf);
            if (t >= kFNBuf && !mbar_wait(&S.empty[b], (unsigned)(((t / kFNBuf) - 1) & 1), err)) break;
            if (lane == 0) mbar_expect_tx(&S.full[b], row_bytes * kFV);
            __syncwarp();
            const int64_t j = min(t * kFV + lane, A.M - 1);
            bulk_copy_g2s(tiles + (size_t)b * kFTileBytes + lane * kFPitch, A.packed + (size_t)j * A.pitch + (size_t)ks0 * 64,
                          row_bytes, &S.full[b]);
        }
    } else if (warp == kFComputeWarps + 1) {
        // ------------------------------------------------------------------ service: publish partial dots, finalise e
        const double unit_b = A.scal[S_UNITB], sumb = A.scal[S_SUMB];
        int sh_cur = INT_MIN;            // adaptive exponent state, identical on every CTA
        double hsum = 0;                 // lane-wise partial of H (fixed order: tile order, then a butterfly)
        for (int64_t s = 0; s < T + kFFinLag; s++) {
            if (s < T) {
                const int a = (int)(s % kFNAcc);
                if (!mbar_wait(&S.acc_done[a], (unsigned)((s / kFNAcc) & 1), err)) break;
                // lane <-> variant: two exact limbs  L = sum_{l<4} I_l 128^l,  Hh = sum_{l>=4} I_l 128^(l-4)
                int *acc = &S.accum[a][lane * 8];
                long long lo = 0, hi = 0;
#pragma unroll
                for (int l = 3; l >= 0; l--) { lo = lo * 128 + acc[l]; hi = hi * 128 + acc[l + 4]; }
#pragma unroll
                for (int l = 0; l < 8; l++) acc[l] = 0;
                unsigned long long *dst = A.acc_t + ((size_t)s * kFV + lane) * 2;
                atomicAdd(dst, (unsigned long long)lo);
                atomicAdd(dst + 1, (unsigned long long)hi);
                __threadfence();
                __syncwarp();
                if (lane == 0) atomicAdd(A.counter + s, 1u);
            }
            const int64_t tb = s - kFFinLag;
            if (tb >= 0) {
                // wait until every CTA has published tile tb
                bool ok = true;
                if (lane == 0) {
                    long long it = 0;
                    while (*((volatile unsigned int *)(A.counter + tb)) < (unsigned)n_cta) {
                        if (++it > kFSpinMax || ((it & 1023) == 0 && *err)) { *err = 1; ok = false; break; }
                    }
                }
                ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
                if (!ok) break;
                __threadfence();
                const int64_t j = tb * kFV + lane;
                double ej = 0, hmj = 0, hj = 0;
                if (j < A.M) {
                    const long long lo = (long long)__ldcg(A.acc_t + (size_t)j * 2), hi = (long long)__ldcg(A.acc_t + (size_t)j * 2 + 1);
                    const double uj = __ldg(A.u + j);
                    const double tsum = (double)lo + 268435456.0 * (double)hi;          // 128^4 = 2^28
                    const double Tj = (unit_b == 0 ? 0.0 : unit_b * tsum) - 3.0 * uj;
                    const double l0 = A.lut[4 * j], inv = A.lut[4 * j + 1] - l0;
                    const double dot = inv * Tj + l0 * (sumb - uj);
                    ej = dot * inv * A.inv_mtotal;
                    hj = dot * l0 * A.inv_mtotal;
                    hmj = hj + 3.0 * ej;
                    if ((int)(tb % n_cta) == (int)blockIdx.x) A.hm[j] = hmj;
                }
                hsum += hj;
                double mx = fabs(ej);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                int flag = 0;
                if (mx > 0 && isfinite(mx)) {
                    const int need = ilogb(mx);
                    if (sh_cur == INT_MIN || need + sh_cur > 54) { sh_cur = 54 - need - kFHead; flag = 1; }
                }
                const int e = (int)(tb % kFNE);
                // (the slot is free: its previous tile tb-4 finished phase B before the compute warps could arrive at
                //  acc_done for step tb+2, which this warp has already passed)
                int8_t d[8];
                to_digits((sh_cur == INT_MIN || !isfinite(ej)) ? 0.0 : ej, sh_cur == INT_MIN ? 0 : sh_cur, d);
                const int hh = lane >> 4, vq = (lane >> 2) & 3, beta = lane & 3;
#pragma unroll
                for (int l = 0; l < 8; l++) S.efrag[e][(l * 4 + vq) * 8 + hh * 4 + beta] = (unsigned char)d[l];
                if (lane == 0) { S.esh[e] = sh_cur; S.eflag[e] = flag; }
                __syncwarp();
                __threadfence_block();
                if (lane == 0) mbar_arrive(&S.ready[e]);
            }
        }
        // H: every CTA holds the same lane-wise partials; CTA 0 publishes the butterfly sum
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
        if (blockIdx.x == 0 && lane == 0) *A.h_total = hsum;
    } else {
        // ------------------------------------------------------------------ compute warps
        // B fragments of phase A (digits of b for this warp's half-steps) stay in registers for the whole product
        uint32_t bfr[kFHpw][8];
#pragma unroll
        for (int i = 0; i < kFHpw; i++) {
            const int hs = warp + i * kFComputeWarps;                 // half-step inside the slice
            const bool on = hs < 2 * ks_n;
            const uint4 *src = reinterpret_cast<const uint4 *>(A.dfrag128 + ((size_t)(ks0 * 2 + hs)) * 1024 + lane * 32);
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
            if (on) { v0 = src[0]; v1 = src[1]; }
            bfr[i][0] = v0.x; bfr[i][1] = v0.y; bfr[i][2] = v0.z; bfr[i][3] = v0.w;
            bfr[i][4] = v1.x; bfr[i][5] = v1.y; bfr[i][6] = v1.z; bfr[i][7] = v1.w;
        }
        int accB[kFRbw][4][4];
#pragma unroll
        for (int r = 0; r < kFRbw; r++)
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int q = 0; q < 4; q++) accB[r][t][q] = 0;
        int sh_mine = INT_MIN;
        bool first_flush = true;
        int tiles_since_flush = 0;
        auto flush = [&]() {
            if (sh_mine != INT_MIN) {
                const double unit = scalbn(1.0, -sh_mine);
                const double w0 = scalbn(unit, 14 * tq), w1 = scalbn(unit, 14 * tq + 7);
#pragma unroll
                for (int r = 0; r < kFRbw; r++)
#pragma unroll
                    for (int t = 0; t < 4; t++)
#pragma unroll
                        for (int hh = 0; hh < 2; hh++) {
                            double v = w0 * (double)(accB[r][t][2 * hh] >> (2 * t)) + w1 * (double)(accB[r][t][2 * hh + 1] >> (2 * t));
                            accB[r][t][2 * hh] = accB[r][t][2 * hh + 1] = 0;
                            v += __shfl_xor_sync(0xffffffffu, v, 1);
                            v += __shfl_xor_sync(0xffffffffu, v, 2);
                            const int64_t n = (ks0 * 64 + (warp * kFRbw + r) * 16 + g + hh * 8) * 4 + t;
                            if (tq == 0 && n < A.N && (warp * kFRbw + r) < 4 * ks_n) {
                                double *dst = A.rout + n;
                                *dst = first_flush ? v : (*dst + v);
                            }
                        }
                first_flush = false;
            }
            tiles_since_flush = 0;
        };
        for (int64_t s = 0; s < T + kFLag; s++) {
            if (s < T) {
                // ---------------- phase A on tile s
                const int b = (int)(s % kFNBuf);
                if (!mbar_wait(&S.full[b], (unsigned)((s / kFNBuf) & 1), err)) break;
                const uint8_t *tile = tiles + (size_t)b * kFTileBytes;
                int accA[2][4][4];
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int t = 0; t < 4; t++)
#pragma unroll
                        for (int q = 0; q < 4; q++) accA[r][t][q] = 0;
#pragma unroll
                for (int i = 0; i < kFHpw; i++) {
                    const int hs = warp + i * kFComputeWarps;
                    if (hs < 2 * ks_n) {
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            // A fragment of [16 variants x 32 bytes]: matrices (rows 0-7 | 8-15) x (bytes 0-15 | 16-31)
                            const uint8_t *src = tile + (r * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kFPitch + hs * 32 + (lane >> 4) * 16;
                            uint32_t a0, a1, a2, a3;
                            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(smem_u32(src)));
#pragma unroll
                            for (int t = 0; t < 4; t++) {
                                const uint32_t m = 0x03030303u << (2 * t);
                                imma_u8s8(accA[r][t], a0 & m, a1 & m, a2 & m, a3 & m, bfr[i][t * 2], bfr[i][t * 2 + 1]);
                            }
                        }
                    }
                }
                // partial dots -> shared int32 atomics: rows g / g+8 of each row-block, digit planes 2tq, 2tq+1
                int *accS = S.accum[s % kFNAcc];
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int v = (accA[r][0][q]) + (accA[r][1][q] >> 2) + (accA[r][2][q] >> 4) + (accA[r][3][q] >> 6);
                        const int row = r * 16 + g + (q >> 1) * 8, plane = 2 * tq + (q & 1);
                        atomicAdd(accS + row * 8 + plane, v);
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.acc_done[s % kFNAcc]);
            }
            const int64_t tb = s - kFLag;
            if (tb >= 0) {
                // ---------------- phase B on tile tb
                const int e = (int)(tb % kFNE);
                if (!mbar_wait(&S.ready[e], (unsigned)((tb / kFNE) & 1), err)) break;
                const int sh_tile = S.esh[e];
                if (sh_tile != sh_mine || tiles_since_flush >= 2048) {
                    if (sh_tile != sh_mine) { flush(); sh_mine = sh_tile; } else flush();
                }
                const uint2 bf = *reinterpret_cast<const uint2 *>(&S.efrag[e][lane * 8]);
                const int b = (int)(tb % kFNBuf);
                const uint8_t *tile = tiles + (size_t)b * kFTileBytes;
#pragma unroll
                for (int r = 0; r < kFRbw; r++) {
                    if ((warp * kFRbw + r) < 4 * ks_n) {
                        const uint8_t *src = tile + lane * kFPitch + (warp * kFRbw + r) * 16;
                        uint32_t x0, x1, x2, x3;
                        asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(smem_u32(src)));
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            const uint32_t m = 0x03030303u << (2 * t);
                            imma_u8s8(accB[r][t], x0 & m, x1 & m, x2 & m, x3 & m, bf.x, bf.y);
                        }
                    }
                }
                tiles_since_flush++;
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.empty[b]);
            }
        }
        flush();
        // slices that never saw a non-zero e still have to define their output
        if (first_flush) {
#pragma unroll
            for (int r = 0; r < kFRbw; r++)
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        const int64_t n = (ks0 * 64 + (warp * kFRbw + r) * 16 + g + hh * 8) * 4 + t;
                        if (tq == 0 && n < A.N && (warp * kFRbw + r) < 4 * ks_n) A.rout[n] = 0.0;
                    }
        }
    }
}

// digits of b for the fused kernel: [half-step (128 samples)][lane = l*4+tq][t0][h][beta]
__global__ void digits_b128_kernel(const double *__restrict__ b, int64_t N, int64_t Npad, const double *__restrict__ scal,
                                   int8_t *__restrict__ dfrag) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Npad) return;
    double unit;
    const int sh = quant_shift(scal[S_MAXB], &unit);
    int8_t d[8];
    const double v = (n < N && unit > 0) ? b[n] : 0.0;
    to_digits(v, sh, d);
    const int64_t hs = n >> 7;
    const int r = (int)(n & 127), h = r >> 6, r2 = r & 63, tq = r2 >> 4, beta = (r2 >> 2) & 3, t0 = r2 & 3;
    int8_t *base = dfrag + hs * 1024 + tq * 32 + (t0 * 2 + h) * 4 + beta;
#pragma unroll
    for (int l = 0; l < 8; l++) base[l * 128] = d[l];
}

// u_j = sum over sample tiles of the sparse partial sums (fixed order)
__global__ void sum_tiles_kernel(const double *__restrict__ part, int n_tiles, int64_t R, double *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    double s = 0;
    for (int t = 0; t < n_tiles; t++) s += part[(size_t)t * R + r];
    out[r] = s;
}

// out_n = R_n (real units) + H - corr_n
__global__ void combine_fused_kernel(const double *__restrict__ rout, int64_t N, const double *__restrict__ cpart, int n_ctiles,
                                     const double *__restrict__ h_total, double *__restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double corr = 0;
    for (int t = 0; t < n_ctiles; t++) corr += cpart[(size_t)t * N + n];
    out[n] = rout[n] + *h_total - corr;
}
