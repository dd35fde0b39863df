// Fused single-HBM-pass GRM product (included by grm_imma.cu inside its anonymous namespace).
//
// One persistent CTA per SM owns a slice of <= 3072 samples (12 K-steps of 256) for the whole product and walks the
// variants in tiles of 32.  A packed tile [32 variants x 768 B] is fetched ONCE from HBM (cp.async.bulk + mbarrier,
// L2 evict-first), used for phase A (partial dots over the CTA's samples), kept in shared memory while the partial
// dots of all CTAs are summed through exact 64-bit integer reductions in L2, and then used again for phase B
// (apply); the int32 accumulators of phase B live in registers for the whole product.
//
//   compute warps 0..7 : step s: wait full[s] and ready[s-lag]; phase A(tile s) and phase B(tile s-lag) issued interleaved;
//                        arrive empty[s-lag]; per-warp partial dots of tile s into shared memory -> part_full
//   loader warp 8      : wait empty -> arm full with expect_tx -> six 2-D TMA copies [32 variants x 128 B], SWIZZLE_128B
//   publisher warp 9   : wait part_full -> add the 8 warps' partials (lane <-> variant) -> two red.add.u64 per
//                        variant.  Every CTA adds 2^52 on top of its value, so a limb carries its own arrival count
//                        in bits 52.. and needs no fence, flag or second round trip.
//   owner warp 10      : for the tiles this CTA owns (tile mod #CTAs): read the limbs once all CTAs have arrived -> dot, e, hm
//                        in FP64 -> base-128 digits of e in MMA fragment order -> publish the self-validating 256-byte block
//   copier warp 11     : fetch the digit blocks of ALL tiles in order (four in flight) into the e-digit ring -> ready[tile]
//
// All cross-warp synchronisation is mbarrier based; there is no CTA-wide barrier in the main loop.  The fixed-point
// exponent of e is chosen BEFORE the launch from a rigorous bound: |e_j| <= |inv_j| sqrt(sum_n lut_j[c_nj]^2) |b|_2 / M
// (Cauchy-Schwarz; the maximum over j of the first two factors is computed once at prepare time from the allele
// counts).  For random b the bound is ~sqrt(N) above the largest |e_j|, i.e. ~12 of the 56 fixed-point bits are head
// room and e is still resolved to ~2^-43 of its largest element; in exchange every tile is independent of all
// others.  Everything that crosses warps or CTAs is integer, so the result is independent of timing and
// bit-reproducible.  Spin loops are bounded: on time-out an error flag is raised and the kernel drains.
#pragma once
#include <cuda.h>   // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

constexpr int kFV = 32;                 // variants per tile
constexpr int kFComputeWarps = 8;
#ifndef SGB_FUSED_NFIN
#define SGB_FUSED_NFIN 2
#endif
constexpr int kFNFin = SGB_FUSED_NFIN;  // finaliser warps: the owner warp and the copier warp
static_assert(kFNFin == 2, "one owner warp and one copier warp");
constexpr int kFWarps = (kFNFin <= 2) ? 12 : 16;   // compute + loader + publisher + finalisers (+ idle: setmaxnreg works on groups of four)
static_assert(kFComputeWarps + 2 + kFNFin <= kFWarps, "warp roles");
constexpr int kFThreads = kFWarps * 32;
constexpr int kFMaxKs = 12;             // K-steps (256 samples) per CTA slice
constexpr int kFRowBytes = kFMaxKs * 64;          // 768
constexpr int kFPanels = kFRowBytes / 128;        // a tile is 6 panels of [32 variants x 128 B], each written by one 2-D TMA copy
constexpr int kFPanelBytes = kFV * 128;           // 4,096; SWIZZLE_128B: 16-byte chunk index ^= (row & 7) -> both ldmatrix patterns conflict-free
constexpr int kFTileBytes = kFPanels * kFPanelBytes;   // 24,576
constexpr int kFNBuf = 9;               // tile ring: loads in flight + tiles waiting for their e digits
constexpr int kFNE = 10;                // e-digit ring (>= kFNBuf: phase B may trail phase A by a full tile ring)
constexpr int kFNPart = 1;              // partial-dot slots (the publisher drains a slot ~10x faster than a tile takes)
constexpr int kFHpw = 3;                // half-steps (128 samples) per compute warp: 24 / 8
constexpr int kFRbw = 6;                // 64-sample row-blocks per compute warp: 48 / 8
// Cross-CTA / cross-warp waits are bounded by WALL CLOCK (%globaltimer), not by iteration counts: a time-sliced, throttled or
// instrumented GPU (MPS, compute-sanitizer, cuda-gdb) slows a valid run down by orders of magnitude and must not turn it into an
// error.  On expiry the kernel raises its error flag and drains; the host then redoes the call with the two-pass kernels, which
// have no inter-CTA waits (api.cu, guarded()).  Default 2 s, env SGB_WAIT_TIMEOUT_MS.
__device__ unsigned long long g_wait_timeout_ns = 2000000000ull;
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kFArrive = 1ull << 52;   // arrival count lives above bit 52 of a limb

struct FusedSmem {
    unsigned long long full[kFNBuf], empty[kFNBuf], part_full[kFNPart], part_free[kFNPart], ready[kFNE];
    unsigned long long deadline;        // %globaltimer value after which every wait of this CTA gives up
    alignas(16) int part[kFNPart][kFComputeWarps][kFV * 8];   // per-warp partial dots x 64: [variant][digit plane]
    alignas(16) unsigned char efrag[kFNE][256];
};
constexpr int kFSmemBytes = kFNBuf * kFTileBytes + (int)sizeof(FusedSmem);   // dynamic shared memory starts 1 KB aligned (checked)
static_assert(kFSmemBytes <= 227 * 1024, "fused kernel shared memory");

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
// bounded wait; returns false on time-out (or when another role has already raised the error flag).  The hot path is a bare
// try_wait loop (a try_wait already suspends the thread for a while); every 1,024 misses the error flag and the kernel's
// wall-clock deadline (a shared-memory word written at kernel start: nothing extra stays live in registers) are looked at.
__device__ __forceinline__ bool mbar_wait_a(unsigned a, unsigned parity, volatile int *err, const volatile unsigned long long *deadline) {
    for (;;) {
#pragma unroll 1
        for (int it = 0; it < 1024; it++) {
            unsigned ok;
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(a), "r"(parity) : "memory");
            if (ok) return true;
        }
        // warp-uniform decision: the callers go on to warp-collective instructions
        if (__any_sync(__activemask(), global_ns() > *deadline)) *err = 1;
        if (__any_sync(__activemask(), *err != 0)) return false;
    }
}
__device__ __forceinline__ bool mbar_wait(unsigned long long *b, unsigned parity, volatile int *err, const volatile unsigned long long *deadline) {
    return mbar_wait_a(smem_u32(b), parity, err, deadline);
}
__device__ __forceinline__ void mbar_arrive_a(unsigned a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ bool mbar_test_a(unsigned a, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar,
                                            unsigned long long policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// register re-allocation between the warp groups (sm_90a+): the compute warps take what the service warps give up
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
// 12 warps: 256 x 216 + 128 x 72 = 384 x 168;  16 warps: 256 x 192 + 256 x 64 = 512 x 128
constexpr int kFRegsCompute = (kFWarps == 12) ? 216 : 192, kFRegsService = (kFWarps == 12) ? 72 : 64;
// same instruction as imma_u8s8 but not volatile: a pure register operation the scheduler may interleave freely
__device__ __forceinline__ void imma_nv(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], unsigned addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_t8(uint32_t (&r)[4], unsigned addr) {
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}

struct FusedArgs {
    const uint8_t *packed; size_t pitch; int64_t M, N, hsteps;        // hsteps = half-steps of 128 samples = 2 ceil(N / 256)
    int hs_per_cta;                     // half-steps per CTA slice (<= 2 kFMaxKs); slicing by half-steps puts 147 of 148 SMs to work at N = 430K
    int lag;                            // phase B runs `lag` tiles behind phase A (1 .. kFNBuf - 2)
    int poll_ns;                        // back-off between two polls of a tile's limbs
    int64_t n_tiles;
    const int8_t *dfrag128;             // [half-step][lane][t0][h][beta]: digits of b, 1 KB per 128 samples
    unsigned long long *acc_t;          // [64 = variant-in-tile x limb][acc_stride] exact integer T'_j limbs + arrival counts (zeroed before
                                        // the launch).  Limb-major: the 64 limbs of one tile sit in 64 different L2 lines / slices, because every
                                        // CTA hits them at about the same time and the L2 atomic unit serialises per line
    int64_t acc_stride;
    const double *u;                    // [M] U_j = sum of b over the missing samples of variant j
    const double *lut;
    double inv_mtotal;
    const double *scal;                 // S_UNITB, S_SUMB, S_FESH, S_FUNITE, S_FEBOUND
    double *hm;                         // [M] out: h_j + 3 e_j for the sparse output correction
    double *h_part;                     // out: [#CTAs] partial sums of H = sum_j h_j over the tiles a CTA owned (added in fixed order afterwards)
    unsigned long long *edig;           // [n_tiles][32] self-validating digit blocks of e published by the tile owners (zeroed before the launch)
    double *rout;                       // [N] out: sum_j e_j c'_nj in real units (this CTA's slice)
    int *err;
};

__global__ void __launch_bounds__(kFThreads, 1) imma_fused_kernel(const __grid_constant__ CUtensorMap tmap, FusedArgs A) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *smem_raw = smem_dyn;
    uint8_t *tiles = smem_raw;
    FusedSmem &S = *reinterpret_cast<FusedSmem *>(smem_raw + kFNBuf * kFTileBytes);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int n_cta = gridDim.x;
    const int64_t hs0 = (int64_t)blockIdx.x * A.hs_per_cta;
    const int hs_n = (int)min((int64_t)A.hs_per_cta, A.hsteps - hs0);        // >= 1 by construction of the grid
    const int64_t T = A.n_tiles;
    const int lag = A.lag;
    volatile int *err = A.err;

    const volatile unsigned long long *dl = &S.deadline;
    if (tid == 0) {
        S.deadline = global_ns() + g_wait_timeout_ns;
        if (smem_u32(smem_raw) & 1023u) *A.err = 3;     // SWIZZLE_128B needs 1 KB aligned tiles
        for (int i = 0; i < kFNBuf; i++) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], kFComputeWarps); }
        for (int i = 0; i < kFNPart; i++) { mbar_init(&S.part_full[i], kFComputeWarps); mbar_init(&S.part_free[i], 1); }
        for (int i = 0; i < kFNE; i++) mbar_init(&S.ready[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= kFComputeWarps) {
      reg_dec<kFRegsService>();
      if (warp == kFComputeWarps) {
        // ------------------------------------------------------------------ loader: lane <-> variant row
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
        for (int64_t t = 0; t < T; t++) {
            const int b = (int)(t % kFNBuf);
            if (t >= kFNBuf && !mbar_wait(&S.empty[b], (unsigned)(((t / kFNBuf) - 1) & 1), err, dl)) break;
            if (lane == 0) mbar_expect_tx(&S.full[b], kFTileBytes);     // out-of-bounds rows / columns are zero-filled and counted
            __syncwarp();
            if (lane < kFPanels)
                tma_load_2d(tiles + (size_t)b * kFTileBytes + lane * kFPanelBytes, &tmap, (int)(hs0 * 32) + lane * 128, (int)(t * kFV),
                            &S.full[b], policy);
        }
      } else if (warp == kFComputeWarps + 1) {
        // ------------------------------------------------------------------ publisher: lane <-> variant
        for (int64_t s = 0; s < T; s++) {
            const int a = (int)(s % kFNPart);
            if (!mbar_wait(&S.part_full[a], (unsigned)((s / kFNPart) & 1), err, dl)) break;
            int x[8];
#pragma unroll
            for (int l = 0; l < 8; l++) x[l] = 0;
            // a lane owns 32 contiguous bytes; lanes 4..7 of every eight read their second 16 bytes first, so that the eight lanes of
            // a quarter-warp hit eight different bank groups (in lane order these loads were 2-way conflicts: 29 M excess wavefronts
            // per product, profiles/r02_ncu_fused_summary.json)
            const int sel = (lane >> 2) & 1;
#pragma unroll
            for (int w = 0; w < kFComputeWarps; w++) {
                const int4 va = *reinterpret_cast<const int4 *>(&S.part[a][w][lane * 8 + 4 * sel]);
                const int4 vb = *reinterpret_cast<const int4 *>(&S.part[a][w][lane * 8 + 4 * (sel ^ 1)]);
                const int4 v0 = sel ? vb : va, v1 = sel ? va : vb;
                x[0] += v0.x; x[1] += v0.y; x[2] += v0.z; x[3] += v0.w;
                x[4] += v1.x; x[5] += v1.y; x[6] += v1.z; x[7] += v1.w;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.part_free[a]);
            // two exact limbs  L = sum_{l<4} I_l 128^l,  Hh = sum_{l>=4} I_l 128^(l-4)   (the partials carry a factor 64)
            long long lo = 0, hi = 0;
#pragma unroll
            for (int l = 3; l >= 0; l--) { lo = lo * 128 + (x[l] >> 6); hi = hi * 128 + (x[l + 4] >> 6); }
            unsigned long long *dst = A.acc_t + (size_t)(2 * lane) * A.acc_stride + s;
            red_add_u64(dst, (unsigned long long)lo + kFArrive);
            red_add_u64(dst + A.acc_stride, (unsigned long long)hi + kFArrive);
        }
      } else if (warp < kFComputeWarps + 2 + kFNFin) {
        // ------------------------------------------------------------------ finalisers: tiles f, f + kFNFin, ...
        const int f = warp - (kFComputeWarps + 2);
        const double unit_b = A.scal[S_UNITB], sumb = A.scal[S_SUMB], ebound = A.scal[S_FEBOUND], eunit = A.scal[S_FUNITE];
        const int esh = (int)A.scal[S_FESH];
        const bool e_on = eunit > 0 && isfinite(eunit);
        const unsigned long long want = (unsigned long long)n_cta;
        const unsigned poll_ns = (unsigned)A.poll_ns;
        double hsum = 0;                 // lane-wise partial of H (fixed order: tile order, then a butterfly)
        // Tile tb is finalised by ONE CTA, its owner (tb mod #CTAs): the owner warp (first finaliser) reads the 64 limbs once
        // they carry all arrivals, computes dot / e / hm in FP64, cuts e into digits and publishes the 256-byte digit block; the
        // copier warp (second finaliser) of every CTA fetches the blocks in tile order.  (With every CTA finalising every tile,
        // 140 x 64 limb reads per tile hit the very L2 lines the reductions go to, and a finaliser spent ~2 us per tile.)
        // A published 8-byte word validates itself: every digit fits in 7 bits, so bit 7 of a byte repeats bit 6, and bit 7 of
        // byte 0 is set instead -- a word that is still zero has not been written (the buffer is zeroed before the launch).
        // Everything published is integer: the result does not depend on who computes it.
        if (f == 0) {
            for (int64_t tb = blockIdx.x; tb < T; tb += n_cta) {
                const int64_t j = tb * kFV + lane;
                const int64_t jc = min(j, A.M - 1);
                const double uj = __ldg(A.u + jc);
                const double l0 = __ldg(A.lut + 4 * jc), l1 = __ldg(A.lut + 4 * jc + 1);
                const unsigned long long *src = A.acc_t + (size_t)(2 * lane) * A.acc_stride + tb;
                unsigned long long x0 = 0, x1 = 0;
                bool ok = true;
                // optimistic full read; while the tile is incomplete only lane 0 probes (its two limbs), backing off in between
                for (int it = 0;;) {
                    x0 = ld_relaxed_u64(src);
                    x1 = ld_relaxed_u64(src + A.acc_stride);
                    const bool done = ((x0 + (kFArrive >> 1)) >> 52) == want && ((x1 + (kFArrive >> 1)) >> 52) == want;
                    if (__all_sync(0xffffffffu, done)) break;
                    bool probe = false;
                    while (!probe) {
                        __nanosleep(poll_ns);
                        if (lane == 0) {
                            const unsigned long long p0 = ld_relaxed_u64(src), p1 = ld_relaxed_u64(src + A.acc_stride);
                            probe = ((p0 + (kFArrive >> 1)) >> 52) == want && ((p1 + (kFArrive >> 1)) >> 52) == want;
                        }
                        if ((++it & 255) == 0 && (*err || global_ns() > *dl)) { *err = 1; ok = false; probe = true; }
                        probe = __shfl_sync(0xffffffffu, probe ? 1 : 0, 0) != 0;
                        ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
                    }
                    if (!ok) break;
                }
                if (!ok) break;
                double ej = 0, hj = 0;
                if (j < A.M) {
                    const long long lo = (long long)(x0 - want * kFArrive), hi = (long long)(x1 - want * kFArrive);
                    const double tsum = (double)lo + 268435456.0 * (double)hi;          // 128^4 = 2^28
                    const double Tj = (unit_b == 0 ? 0.0 : unit_b * tsum) - 3.0 * uj;
                    const double inv = l1 - l0;
                    const double dot = inv * Tj + l0 * (sumb - uj);
                    ej = dot * inv * A.inv_mtotal;
                    hj = dot * l0 * A.inv_mtotal;
                }
                // digits d_l in [-64, 63] of the 56-bit fixed-point value, all at once: adding 64 to every digit position
                // turns them into the plain base-128 digits of a non-negative number.  Lane <-> variant here; the 8-byte word
                // w of the block holds digit plane w >> 2 of the variants 16 hh + 4 (w & 3) + beta (bytes hh * 4 + beta), so
                // the block is transposed through the warp with shuffles.
                const long long Bq = __double2ll_rn(scalbn((e_on && isfinite(ej)) ? ej : 0.0, esh)) + 0x1020408102040LL;   // sum_{l<7} 64 * 128^l
                unsigned long long word = 0;
                const int pl = lane >> 2, vq = lane & 3;                               // this lane assembles word `lane`
#pragma unroll
                for (int by = 0; by < 8; by++) {
                    const int srcl = (by >> 2) * 16 + vq * 4 + (by & 3);               // variant whose digit goes into byte `by`
                    const long long Bs = __shfl_sync(0xffffffffu, Bq, srcl);
                    const int dl = (pl < 7) ? ((int)((Bs >> (7 * pl)) & 127) - 64) : (int)(Bs >> 49);
                    word |= (unsigned long long)(unsigned)(dl & 0xFF) << (8 * by);
                }
                word |= 0x80ull;                                                       // valid
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(A.edig + (size_t)tb * 32 + lane), "l"(word) : "memory");
                // off the critical path: H, the output-correction weights, the sanity check of the bound
                hsum += hj;
                if (j < A.M) {
                    A.hm[j] = hj + 3.0 * ej;
                    if (e_on && fabs(ej) > ebound) *err = 2;      // the a-priori bound must hold; fail loudly if it ever does not
                }
            }
        } else {
            // ---- copier: digit blocks in tile order, four tiles' words in flight
            unsigned long long w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            const unsigned long long *gd = A.edig + lane;
            if (0 < T) w0 = ld_relaxed_u64(gd);
            if (1 < T) w1 = ld_relaxed_u64(gd + 32);
            if (2 < T) w2 = ld_relaxed_u64(gd + 64);
            if (3 < T) w3 = ld_relaxed_u64(gd + 96);
            bool ok = true;
            for (int64_t tb = 0; tb < T; tb++) {
                for (int it = 0; !__all_sync(0xffffffffu, w0 != 0); it++) {
                    __nanosleep(poll_ns);
                    if ((it & 255) == 255 && __any_sync(0xffffffffu, *err || global_ns() > *dl)) { *err = 1; ok = false; break; }
                    if (w0 == 0) w0 = ld_relaxed_u64(gd + tb * 32);
                }
                if (!ok) break;
                const int e = (int)(tb % kFNE);
                // (the e-digit slot is free: its previous tile finished phase B before this CTA's phase A of tile tb completed,
                //  which the arrival count of tile tb includes)
                const unsigned long long dg = (w0 & ~0x80ull) | ((w0 & 0x40ull) << 1);     // bit 7 of byte 0 back to the sign
                *reinterpret_cast<unsigned long long *>(&S.efrag[e][lane * 8]) = dg;
                __syncwarp();
                __threadfence_block();
                if (lane == 0) mbar_arrive(&S.ready[e]);
                w0 = w1; w1 = w2; w2 = w3;
                w3 = (tb + 4 < T) ? ld_relaxed_u64(gd + (tb + 4) * 32) : 0;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) hsum += __shfl_xor_sync(0xffffffffu, hsum, o);
        if (lane == 0 && f == 0) A.h_part[blockIdx.x] = hsum;        // partial H of the tiles this CTA owned
      }
    } else {
        // ------------------------------------------------------------------ compute warps
        reg_inc<kFRegsCompute>();
        // B fragments of phase A (digits of b for this warp's half-steps) stay in registers for the whole product
        uint32_t bfr[kFHpw][8];
#pragma unroll
        for (int i = 0; i < kFHpw; i++) {
            const int hs = warp + i * kFComputeWarps;                 // half-step inside the slice
            const bool on = hs < hs_n;
            const uint4 *src = reinterpret_cast<const uint4 *>(A.dfrag128 + ((size_t)(hs0 + hs)) * 1024 + lane * 32);
            uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
            if (on) { v0 = src[0]; v1 = src[1]; }
            bfr[i][0] = v0.x; bfr[i][1] = v0.y; bfr[i][2] = v0.z; bfr[i][3] = v0.w;
            bfr[i][4] = v1.x; bfr[i][5] = v1.y; bfr[i][6] = v1.z; bfr[i][7] = v1.w;
        }
        // phase-B accumulators: [row-block][t][fragment]; t = 0 holds the UNMASKED bytes (c0 + 4 c1 + 16 c2 + 64 c3),
        // i.e. the sum of all four planes -- plane 0 is recovered exactly by subtraction when flushing
        int accB[kFRbw][4][4];
#pragma unroll
        for (int r = 0; r < kFRbw; r++)
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int q = 0; q < 4; q++) accB[r][t][q] = 0;
        // swizzled fragment addresses inside a tile.  Phase A, half-step hs = warp + 8 i, row-block r: byte x = 32 hs + 16 (lane >> 4)
        // of row 16 r + (lane & 15): panel x >> 7, chunk ((x >> 4) & 7) ^ (row & 7).  Phase B, row-block g = 6 warp + r: 16-byte
        // chunk g of row `lane`.
        const unsigned a_off = (unsigned)((warp >> 2) * kFPanelBytes + (lane & 15) * 128 + (((((warp & 3) << 1) | (lane >> 4)) ^ (lane & 7)) << 4));
        unsigned b_off[kFRbw];
#pragma unroll
        for (int r = 0; r < kFRbw; r++) {
            const int gch = warp * kFRbw + r;
            b_off[r] = (unsigned)((gch >> 3) * kFPanelBytes + lane * 128 + (((gch & 7) ^ (lane & 7)) << 4));
        }
        const double eunit = A.scal[S_FUNITE];
        const bool e_on = eunit > 0 && isfinite(eunit);
        bool first_flush = true;
        auto flush = [&]() {
            const double w0 = scalbn(e_on ? eunit : 0.0, 14 * tq), w1 = scalbn(e_on ? eunit : 0.0, 14 * tq + 7);
#pragma unroll
            for (int r = 0; r < kFRbw; r++)
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        int s0, s1;
                        if (t == 0) {
                            s0 = accB[r][0][2 * hh] - accB[r][1][2 * hh] - accB[r][2][2 * hh] - accB[r][3][2 * hh];
                            s1 = accB[r][0][2 * hh + 1] - accB[r][1][2 * hh + 1] - accB[r][2][2 * hh + 1] - accB[r][3][2 * hh + 1];
                        } else {
                            s0 = accB[r][t][2 * hh] >> (2 * t);
                            s1 = accB[r][t][2 * hh + 1] >> (2 * t);
                        }
                        double v = w0 * (double)s0 + w1 * (double)s1;
                        v += __shfl_xor_sync(0xffffffffu, v, 1);
                        v += __shfl_xor_sync(0xffffffffu, v, 2);
                        const int64_t n = (hs0 * 32 + (warp * kFRbw + r) * 16 + g + hh * 8) * 4 + t;
                        if (tq == 0 && n < A.N && (warp * kFRbw + r) < 2 * hs_n) {
                            double *dst = A.rout + n;
                            *dst = first_flush ? v : (*dst + v);
                        }
                    }
#pragma unroll
            for (int r = 0; r < kFRbw; r++)
#pragma unroll
                for (int t = 0; t < 4; t++)
#pragma unroll
                    for (int q = 0; q < 4; q++) accB[r][t][q] = 0;
            first_flush = false;
        };
        // everything inside the tile loop uses 32-bit shared-window addresses and incrementally maintained ring indices
        const unsigned sb = smem_u32(smem_raw), so = sb + kFNBuf * kFTileBytes;
        const unsigned full0 = so + (unsigned)offsetof(FusedSmem, full), empty0 = so + (unsigned)offsetof(FusedSmem, empty);
        const unsigned pfull0 = so + (unsigned)offsetof(FusedSmem, part_full), pfree0 = so + (unsigned)offsetof(FusedSmem, part_free);
        const unsigned ready0 = so + (unsigned)offsetof(FusedSmem, ready);
        unsigned part_w = so + (unsigned)offsetof(FusedSmem, part) + warp * (kFV * 8 * 4) + (g * 8 + 2 * tq) * 4;   // this thread's first int2
        unsigned efrag_l = so + (unsigned)offsetof(FusedSmem, efrag) + lane * 8;
        unsigned a_addr = sb + a_off;
        // keep the addresses in registers (the compiler would otherwise re-derive them from %tid in every tile)
        asm volatile("" : "+r"(part_w), "+r"(efrag_l), "+r"(a_addr));
#pragma unroll
        for (int r = 0; r < kFRbw; r++) { b_off[r] += sb; asm volatile("" : "+r"(b_off[r])); }
        const int Ti = (int)T;
        int bA = 0, phA = 0;            // tile ring slot / parity of tile s
        int bB = 0;                     // tile ring slot of tile tb
        int phP = 1;                    // parity of the part_free wait (first use of the slot: no wait)
        int eB = 0, phE = 0;            // e-digit slot / parity of tile tb
        bool alive = true;
        // One step = phase A on tile s and phase B on tile tb = s - lag.  Both waits first, then all twelve fragment loads, then
        // the 48 tensor-core instructions of the two phases interleaved (independent work for the scheduler), then the stores.
        auto step = [&](auto do_a, auto do_b, int s) -> bool {
            constexpr bool DA = decltype(do_a)::value, DB = decltype(do_b)::value;
            uint32_t fa[2][kFHpw][4], fb[kFRbw][4], bfx = 0, bfy = 0;
            int accA[2][4][4];
            if (DA) {
                if (!mbar_wait_a(full0 + bA * 8, (unsigned)phA, err, dl)) return false;
            }
            if (DB) {
                if (!mbar_wait_a(ready0 + eB * 8, (unsigned)phE, err, dl)) return false;
            }
            if (DA) {
                // A fragments ([16 variants x 32 bytes]: matrices (rows 0-7 | 8-15) x (bytes 0-15 | 16-31)); half-steps beyond
                // the slice see zeros (TMA zero fill, zero digits)
                const unsigned a_base = a_addr + bA * kFTileBytes;
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int i = 0; i < kFHpw; i++) ldsm_x4(fa[r][i], a_base + r * 16 * 128 + i * 2 * kFPanelBytes);
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int t = 0; t < 4; t++)
#pragma unroll
                        for (int q = 0; q < 4; q++) accA[r][t][q] = 0;
            }
            if (DB) {
                asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(bfx), "=r"(bfy) : "r"(efrag_l + eB * 256) : "memory");
                const unsigned tb_base = bB * kFTileBytes;
#pragma unroll
                for (int r = 0; r < kFRbw; r++) ldsm_t8(fb[r], b_off[r] + tb_base);
            }
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (DA) {
                    const int i = k >> 1, r = k & 1;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const uint32_t m = 0x03030303u << (2 * t);
                        imma_nv(accA[r][t], fa[r][i][0] & m, fa[r][i][1] & m, fa[r][i][2] & m, fa[r][i][3] & m, bfr[i][t * 2], bfr[i][t * 2 + 1]);
                    }
                }
                if (DB) {
                    imma_nv(accB[k][0], fb[k][0], fb[k][1], fb[k][2], fb[k][3], bfx, bfy);
#pragma unroll
                    for (int t = 1; t < 4; t++) {
                        const uint32_t m = 0x03030303u << (2 * t);
                        imma_nv(accB[k][t], fb[k][0] & m, fb[k][1] & m, fb[k][2] & m, fb[k][3] & m, bfx, bfy);
                    }
                }
            }
            if (DB) {
                __syncwarp();
                if (lane == 0) mbar_arrive_a(empty0 + bB * 8);
                if (++bB == kFNBuf) bB = 0;
                if (++eB == kFNE) { eB = 0; phE ^= 1; }
            }
            if (DA) {
                // partial dots -> this warp's slot: rows g / g+8 of each row-block, digit planes 2tq, 2tq+1.
                // ((a0*4 + a1)*4 + a2)*4 + a3 = 64 * (sum of the four planes with their 4^t factors removed)
                if (s >= 1 && !mbar_wait_a(pfree0, (unsigned)phP, err, dl)) return false;
#pragma unroll
                for (int r = 0; r < 2; r++)
#pragma unroll
                    for (int hh = 0; hh < 2; hh++) {
                        int v[2];
#pragma unroll
                        for (int c = 0; c < 2; c++) {
                            const int q = 2 * hh + c;
                            v[c] = ((accA[r][0][q] * 4 + accA[r][1][q]) * 4 + accA[r][2][q]) * 4 + accA[r][3][q];
                        }
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(part_w + (r * 16 + hh * 8) * 32), "r"(v[0]), "r"(v[1]) : "memory");
                    }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(pfull0);
                if (++bA == kFNBuf) { bA = 0; phA ^= 1; }
                phP ^= 1;
            }
            return true;
        };
        const std::true_type yes;
        const std::false_type no;
        const int ramp = min(lag, Ti);
        for (int s = 0; s < ramp && alive; s++) alive = step(yes, no, s);                     // tiles 0 .. lag-1: phase A only
        for (int s0 = ramp; s0 < Ti && alive; s0 += 2048) {                                    // steady state
            const int s1 = min(Ti, s0 + 2048);
            for (int s = s0; s < s1 && alive; s++) alive = step(yes, yes, s);
            // at most 2048 (+ lag) tiles went into the int32 accumulators (unmasked plane: 255 * 64 * 32 per tile < 2^31 / 4096)
            flush();
        }
        for (int s = Ti; s < Ti + ramp && alive; s++) alive = step(no, yes, s);                // drain: phase B only
        flush();
    }
    __syncthreads();
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
                                     const double *__restrict__ h_part, int n_hpart, double *__restrict__ out) {
    __shared__ double sh_h;
    if (threadIdx.x == 0) {              // H = the owners' partial sums in a fixed order
        double h = 0;
        for (int i = 0; i < n_hpart; i++) h += h_part[i];
        sh_h = h;
    }
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double corr = 0;
    for (int t = 0; t < n_ctiles; t++) corr += cpart[(size_t)t * N + n];
    out[n] = rout[n] + sh_h - corr;
}
