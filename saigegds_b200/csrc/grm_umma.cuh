// Batched K-right-hand-side GRM product on the 5th-generation tensor cores (included by grm_imma.cu inside its anonymous
// namespace, after grm_fused.cuh whose mbarrier / TMA helpers it reuses).
//
// Reference call sites that are K independent products run back to back: the 30 trace solves (saige_fitnull.cpp:646-654),
// the (1+p) solves of get_coeff_w (:744-752), the variance-ratio markers (:1321).  Here one pass over the packed matrix
// serves all K columns.  Every column is quantised to 46-bit fixed point relative to its own largest element and cut into
// six signed base-256 digits; the two halves of the product are then exact integer GEMMs
//     T'[j][(c,l)] = sum_n code[n][j] * digitB[n][(c,l)]          (phase A: rows = variants, contraction over samples)
//     R [n][(c,l)] = sum_j code[n][j] * digitE[j][(c,l)]          (phase B: rows = samples,  contraction over variants)
// with MMA-N = 6 K <= 192 columns, executed by tcgen05.mma kind::i8 (u8 x s8 -> s32) with the accumulators in tensor memory.
//
// One kernel does both phases: it multiplies a row-major 2-bit matrix P[R][pitch] (contraction index contiguous) by a digit
// matrix D[6K][C].  Phase A runs it on the packed genotype matrix as stored (variant rows), phase B on a sample-major copy
// built once per store (both orientations stay in HBM: 2 x 10.75 GB at N = 430K, M = 100K -- the transposed copy is what
// lets the 2-bit -> 8-bit expansion go straight from registers into tensor memory in both phases).
//
//   warp 0       producer : 2-D TMA boxes [256 rows x 128 B] of P (SWIZZLE_128B, L2 evict-first; 512 contraction elements = 4
//                           stages) and, per stage of 128 elements, the digit tile [6K x 128 B] (SWIZZLE_128B, K-major; the
//                           digit matrix stays in L2, which feeds all SMs at 22 TB/s: profiles/r02_l2_probe_b200.txt)
//   warps 4..19  expanders: two sets of 8 warps, thread <-> row; set s serves the stages st % 2 == s.  32 packed bytes -> 32
//                           registers of u8 codes (`(w >> 2t) & 0x03030303`) -> one tcgen05.st.32x32b.x32 into the A-operand
//                           slot of its M-tile in tensor memory.  A never touches shared memory (TS operands: 7,558 vs 6,028
//                           MAC/clk/SM for SS at N = 240, profiles/r02_umma_probe_b200.txt)
//   warps 1, 2   issuers  : one elected thread per M = 128 tile: per stage one mbarrier wait, four
//                           tcgen05.mma.cta_group::1.kind::i8 [acc], [A in TMEM], B descriptor, one tcgen05.commit that frees
//                           the A slot (expanders) and the digit tile (producer)
//   epilogue (expanders)  : tcgen05.ld the int32 accumulators, assemble two exact int64 limbs per (row, column)
//                           (sum_l acc_l 256^l for l < 3 and l >= 3) and add them to global limb planes with red.add.u64 --
//                           integer, hence independent of the split of the contraction range and of the order of arrival;
//                           limbs -> FP64 is a single rounding, so a column's result does not depend on its batch.
//
// Tensor memory (512 columns): accumulators 2 x 192, A-operand slots 2 M-tiles x 2 x 32 columns.
// Code 3 (missing) is multiplied like any other code; the sparse missing-genotype corrections remove 3 b_n / 3 e_j together
// with the mean-imputation terms, exactly as in the single-RHS path.  Contraction order inside a group of 16 elements is
// permuted (byte 4t + q of a unit <-> element 4q + t) because that is what the mask-and-shift expansion yields; the digit
// matrices are written in the same order.
#pragma once

constexpr int kUND = 6;                 // signed base-256 digits per value: 46..47-bit fixed point relative to the column's largest element
constexpr int kUMaxCols = 32;           // columns per pass
constexpr int kUMaxN = kUND * kUMaxCols;            // 192 MMA columns
constexpr int kURows = 256;             // rows per CTA: two M = 128 tiles
constexpr int kUBoxBytes = 128;         // packed bytes per row and TMA box = 512 contraction elements
constexpr int kUBoxElems = kUBoxBytes * 4;
constexpr int kUStage = 128;            // contraction elements per digit tile (128 B per MMA column = 4 k-steps of 32)
constexpr int kUNP = 3;                 // packed-box ring (its loads are issued by their own lane, up to three boxes = 12 stages ahead)
constexpr int kUNB = 4;                 // digit-tile ring (= number of mma_done barriers per M-tile)
constexpr int kUPackBytes = kURows * kUBoxBytes;    // 32 KB
constexpr int kUWarps = 20;             // 0: TMA producer (packed boxes), 1 / 2: MMA issuers of M-tile 0 / 1, 3: TMEM allocator + TMA producer (digit tiles), 4..19: expanders
constexpr int kUThreads = kUWarps * 32;
constexpr int kUExpWarps = 16;          // two sets of 8 (256 rows); set s serves the stages with st % 2 == s and owns A slot s
// Tensor memory (512 columns).  Up to 32 right-hand sides (N <= 192): accumulators at 192 mt, two A-operand slots of 32 columns per M-tile at
// 384 + 64 mt + 32 slot.  Up to 21 right-hand sides (N <= 128): accumulators at 128 mt, FOUR A slots per M-tile at 256 + 128 mt + 32 slot --
// the hand-over loop issuer <-> expanders is ~1,600 clk long, and with few columns a stage's MMAs are far shorter than that.
__host__ __device__ inline int umma_ns(int ng) { return ng <= 128 ? 4 : 2; }
// Measured on the B200 (SGB_UMMA_PROF, tools/umma_prof.py): for the single issuing thread an mbarrier try_wait costs ~190 clk, a
// tcgen05.mma ~65 clk and a tcgen05.commit ~130 clk of issue latency.  Hence one wait, four MMAs and one commit per stage and
// issuer (~580 clk) against the 768 clk the eight MMAs of a stage keep the tensor pipe busy at K = 32 columns; one issuer per
// M-tile; the digit-tile dependency is checked by the expanders before they hand the A slot over.

struct UmmaSmem {
    unsigned long long p_full[kUNP], p_empty[kUNP], b_full[kUNB], a_full[2][4], mma_done[2][kUNB], acc_full[2];
    unsigned long long deadline;        // %globaltimer value after which every wait of this CTA gives up
    unsigned tmem_base;
};
constexpr int kUSmemBytes = kUNP * kUPackBytes + kUNB * kUMaxN * 128 + 1024 /* alignment slack */ + (int)sizeof(UmmaSmem);
static_assert(kUSmemBytes <= 227 * 1024, "batched kernel shared memory");

__device__ __forceinline__ void tma_load_2d_nohint(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_arrive(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, A u8, B s8, D s32
__device__ __forceinline__ void umma_i8_ts(unsigned d_tmem, unsigned a_tmem, uint64_t bdesc, unsigned idesc, unsigned accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// cta_group::2: one instruction drives the tensor cores of both SMs of a CTA pair (M = 256: 128 rows per CTA, each CTA's A rows in its
// own tensor memory, B split by rows over the two CTAs' shared memory at the same offset); issued by the leader CTA only
__device__ __forceinline__ void umma_i8_ts_pair(unsigned d_tmem, unsigned a_tmem, uint64_t bdesc, unsigned idesc, unsigned accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
// the arrival lands on the barrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_arrive_pair(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(unsigned long long *bar, unsigned rank) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");     // (the form CUTLASS' ClusterBarrier::arrive uses)
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D TMA load of a CTA pair: the bytes land in this CTA's shared memory, the transaction count on the barrier of CTA `bar_rank`
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar, unsigned bar_rank) {
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(bar_rank));
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(ra) : "memory");
}
// K-major, 128-byte-swizzled operand tile (rows of 128 B, 8-row atoms of 1 KB): start >> 4 | LBO | SBO = 1024 >> 4 | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(unsigned addr) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::i8: D s32 (2 << 4), A u8 (0 << 7), B s8 (1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ inline unsigned umma_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_st32(unsigned taddr, const uint32_t (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
                 "%25,%26,%27,%28,%29,%30,%31,%32};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                   "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
                   "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(unsigned taddr, int *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

struct UmmaArgs {
    int ncols;                 // K right-hand sides of this pass (<= 32)
    int ng;                    // MMA N = 6 K rounded up to 16
    int64_t R;                 // rows of P (variants / samples)
    int boxes_per_split;       // contraction range of one split, in TMA boxes of 512 elements
    int boxes_total;
    unsigned long long *out_lo, *out_hi;   // [K][ldo] exact limbs (zeroed before the launch)
    int64_t ldo;
    int *err;
    long long *prof;           // PROF build only: per CTA 16 cycle counters (tools/umma_prof.py)
    int amode = 0;             // pair kernel: value of the A operand per 2-bit code -- 0: the code itself (0, 1, 2, 3), the GRM product;
                               // 1: its low bit, 2: its high bit, 3: low & high (code 3 = missing): the class sums of the score scan (score.cu)
};

template <bool PROF, bool PAIR>
__global__ void __launch_bounds__(kUThreads, 1) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                                                 const __grid_constant__ CUtensorMap tmap_d, UmmaArgs A) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t *pack = base;                                   // kUNP x 32 KB
    uint8_t *btile = base + kUNP * kUPackBytes;             // kUNB x (ng x 128 B)
    // PAIR: the kernel runs in clusters of two CTAs along x; CTA `crank` holds rows [ng / 2 crank, + ng / 2) of every digit tile
    unsigned crank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const unsigned btile_bytes = (unsigned)(PAIR ? A.ng / 2 : A.ng) * 128u;
    UmmaSmem &S = *reinterpret_cast<UmmaSmem *>(btile + kUNB * kUMaxN * 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    volatile int *err = A.err;
    const int box0 = blockIdx.y * A.boxes_per_split;
    const int n_box = min(A.boxes_per_split, A.boxes_total - box0);
    const int n_st = n_box * 4;                             // stages of 128 contraction elements
    const int row0 = blockIdx.x * kURows;

    const volatile unsigned long long *dl = &S.deadline;
    if (tid == 0) {
        S.deadline = global_ns() + g_wait_timeout_ns;
        for (int i = 0; i < kUNP; i++) { mbar_init(&S.p_full[i], 1); mbar_init(&S.p_empty[i], kUExpWarps); }
        for (int i = 0; i < kUNB; i++) { mbar_init(&S.b_full[i], 1); mbar_init(&S.mma_done[0][i], 1); mbar_init(&S.mma_done[1][i], 1); }
        for (int i = 0; i < 4; i++) { mbar_init(&S.a_full[0][i], PAIR ? 8 : 4); mbar_init(&S.a_full[1][i], PAIR ? 8 : 4); }   // leader: both CTAs' expanders
        for (int i = 0; i < 2; i++) mbar_init(&S.acc_full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all(); else __syncthreads();     // barriers initialised and tensor memory allocated in both CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = S.tmem_base;
    const int ns = umma_ns(A.ng);                          // A slots per M-tile
    const unsigned acc_stride = ns == 4 ? 128u : 192u, stage_col = ns == 4 ? 256u : 384u, stage_mt = (unsigned)ns * 32u;

    if (n_st > 0) {
    if (warp == 0) {
        // ------------------------------------------------------------------ producer
        // packed boxes: they depend on the expanders only.  (Issued from one loop together with the digit tiles, the box of the next
        // four stages would be requested just four stages ahead and arrive late; two lanes of ONE warp do not work either -- a lane
        // spinning in try_wait starves the other, 6.2 instead of 4.2 ms per phase -- so the digit tiles have their own warp, 3.)
        if (lane == 0) {
            unsigned long long policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            for (int bx = 0; bx < n_box; bx++) {
                const int ps = bx % kUNP;
                if (bx >= kUNP && !mbar_wait(&S.p_empty[ps], (unsigned)(((bx / kUNP) - 1) & 1), err, dl)) break;
                mbar_expect_tx(&S.p_full[ps], kUPackBytes);
                tma_load_2d(pack + (size_t)ps * kUPackBytes, &tmap_p, (box0 + bx) * kUBoxBytes, row0, &S.p_full[ps], policy);
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------------ digit tiles: they depend on the MMAs only
        if (lane == 0) {
            for (int st = 0; st < n_st; st++) {
                const int bs = st % kUNB;
                if (st >= kUNB) {     // the tile's previous user (stage st - 4) must be through both M-tiles
                    const unsigned par = (unsigned)(((st / kUNB) - 1) & 1);
                    if (!(mbar_wait(&S.mma_done[0][bs], par, err, dl) && mbar_wait(&S.mma_done[1][bs], par, err, dl))) break;
                }
                mbar_expect_tx(&S.b_full[bs], btile_bytes);
                tma_load_2d_nohint(btile + (size_t)bs * btile_bytes, &tmap_d, (box0 * 4 + st) * kUStage, PAIR ? (int)crank * (A.ng / 2) : 0,
                                   &S.b_full[bs]);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ------------------------------------------------------------------ MMA issuer of M-tile mt: one wait, four MMAs, one commit per stage
        if (lane == 0 && crank == 0) {
            const int mt = warp - 1;
            const unsigned idesc = umma_idesc_i8(PAIR ? 256 : 128, A.ng);
            const unsigned bt0 = smem_u32(btile);
            const unsigned acc = tb + mt * acc_stride, a0 = tb + stage_col + mt * stage_mt;
            bool ok = true;
            long long c_wa = 0, c_mma = 0, c_cm = 0, t0 = 0, t1 = 0;
            const long long t_begin = PROF ? clock64() : 0;
            for (int st = 0; st < n_st && ok; st++) {
                const int bs = st % kUNB, as = st % ns;
                if (PROF) t0 = clock64();
                ok = mbar_wait(&S.a_full[mt][as], (unsigned)((st / ns) & 1), err, dl);
                if (!ok) break;
                if (PROF) { t1 = clock64(); c_wa += t1 - t0; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned bb = bt0 + (unsigned)bs * btile_bytes;
#pragma unroll
                for (int ks = 0; ks < 4; ks++)
                    if (PAIR) umma_i8_ts_pair(acc, a0 + as * 32 + ks * 8, umma_desc_sw128(bb + ks * 32), idesc, (st > 0 || ks > 0) ? 1u : 0u);
                    else umma_i8_ts(acc, a0 + as * 32 + ks * 8, umma_desc_sw128(bb + ks * 32), idesc, (st > 0 || ks > 0) ? 1u : 0u);
                if (PROF) { t0 = clock64(); c_mma += t0 - t1; }
                if (PAIR) umma_commit_arrive_pair(&S.mma_done[mt][bs]); else umma_commit_arrive(&S.mma_done[mt][bs]);
                if (PROF) { t1 = clock64(); c_cm += t1 - t0; }
            }
            if (PAIR) umma_commit_arrive_pair(&S.acc_full[mt]); else umma_commit_arrive(&S.acc_full[mt]);
            if (PROF && mt == 0) {
                long long *o = A.prof + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
                o[0] = clock64() - t_begin; o[1] = 0; o[2] = c_wa; o[3] = c_mma; o[4] = c_cm; o[5] = n_st;
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ expanders: thread <-> row; set `es` serves stages st % 2 == es
        const int e = warp - 4, es = e >> 3, mt = (e >> 2) & 1, quarter = warp & 3;
        const int r = mt * 128 + quarter * 32 + lane;             // row inside the CTA tile
        const unsigned lane_base = (unsigned)(quarter * 32) << 16;
        const unsigned prow = (unsigned)r * kUBoxBytes, psw = (unsigned)(r & 7);
        bool ok = true;
        long long c_wp = 0, c_ld = 0, c_we = 0, c_st = 0, c_ar = 0, t0 = 0, t1 = 0;
        const long long t_begin = PROF ? clock64() : 0;
        for (int st = es; st < n_st && ok; st += 2) {
            const int bx = st >> 2, ps = bx % kUNP, bs = st % kUNB;
            if (PROF) t0 = clock64();
            if ((st & 3) < 2) ok = mbar_wait(&S.p_full[ps], (unsigned)((bx / kUNP) & 1), err, dl);     // first stage of this set in the box
            if (!ok) break;
            if (PROF) { t1 = clock64(); c_wp += t1 - t0; }
            const uint8_t *src = pack + (size_t)ps * kUPackBytes + prow;
            const unsigned c0 = (unsigned)(st & 3) * 2;
            const uint4 wa = *reinterpret_cast<const uint4 *>(src + ((c0 ^ psw) << 4));
            const uint4 wb = *reinterpret_cast<const uint4 *>(src + (((c0 + 1) ^ psw) << 4));
            const uint32_t ws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            uint32_t x[32];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int t = 0; t < 4; t++) x[4 * i + t] = (ws[i] >> (2 * t)) & 0x03030303u;
            if (PROF) { t0 = clock64(); c_ld += t0 - t1; }
            // A slot st % ns is free when the MMAs of stage st - ns (this M-tile) are done
            if (st >= ns) ok = mbar_wait(&S.mma_done[mt][(st - ns) % kUNB], (unsigned)(((st - ns) / kUNB) & 1), err, dl);
            if (!ok) break;
            if (PROF) { t1 = clock64(); c_we += t1 - t0; }
            tmem_st32(tb + lane_base + stage_col + mt * stage_mt + (unsigned)(st % ns) * 32u, x);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (PROF) { t0 = clock64(); c_st += t0 - t1; }
            // the issuer waits on a_full only: make sure the stage's digit tile has landed before handing over
            ok = mbar_wait(&S.b_full[bs], (unsigned)((st / kUNB) & 1), err, dl);
            if (!ok) break;
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(&S.a_full[mt][st % ns], 0u); else mbar_arrive(&S.a_full[mt][st % ns]);
                if ((st & 3) >= 2 || st + 2 >= n_st) mbar_arrive(&S.p_empty[ps]);       // this set's last stage in the box
            }
            if (PROF) { t1 = clock64(); c_ar += t1 - t0; }
        }
        if (PROF && warp == 4 && lane == 0) {
            long long *o = A.prof + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
            o[8] = clock64() - t_begin; o[9] = c_wp; o[10] = c_ld; o[11] = c_we; o[12] = c_st; o[13] = c_ar;
        }
        // ------------------------------------------------------------------ epilogue: set es takes the column groups g % 2 == es
        if (ok) ok = mbar_wait(&S.acc_full[mt], 0u, err, dl);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            const int64_t row = (int64_t)row0 + r;
            for (int g = es; g * 8 < A.ncols; g += 2) {
                int v[8 * kUND];
#pragma unroll
                for (int i = 0; i < kUND; i++) tmem_ld8_nowait(tb + lane_base + mt * acc_stride + g * (8 * kUND) + i * 8, v + 8 * i);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int cc = 0; cc < 8; cc++) {
                    const int col = g * 8 + cc;
                    long long lo = 0, hi = 0;
#pragma unroll
                    for (int l = 2; l >= 0; l--) lo = lo * 256 + v[cc * kUND + l];
#pragma unroll
                    for (int l = 5; l >= 3; l--) hi = hi * 256 + v[cc * kUND + l];
                    if (col < A.ncols && row < A.R) {
                        red_add_u64(A.out_lo + (size_t)col * A.ldo + row, (unsigned long long)lo);
                        red_add_u64(A.out_hi + (size_t)col * A.ldo + row, (unsigned long long)hi);
                    }
                }
            }
        }
    }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (PAIR) cluster_sync_all(); else __syncthreads();     // the peer may still read this CTA's shared / tensor memory until here
    if (warp == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    }
}

// ---- sample-major copy of the packed matrix: PT[n][j / 4] bits 2 (j % 4) = code of (sample n, variant j) -------------------
// tile = 128 variants x 128 B (512 samples) through shared memory; thread <-> sample writes the 32 bytes (128 variants) of its row
__global__ void __launch_bounds__(512) transpose_2bit_kernel(const uint8_t *__restrict__ packed, size_t pitch, int64_t M, int64_t n_rows_t,
                                                             uint8_t *__restrict__ pt, size_t pitch_t) {
    __shared__ uint32_t tile[128][33];
    const int64_t v0 = (int64_t)blockIdx.y * 128, b0 = (int64_t)blockIdx.x * 128;   // variant / packed-byte origin
    for (int i = threadIdx.x; i < 128 * 32; i += 512) {
        const int v = i >> 5, wq = i & 31;
        uint32_t w = 0;                                     // variants beyond M: code 0 (their digits of e are zero anyway)
        if (v0 + v < M && (size_t)(b0 + 4 * wq) < pitch) w = *reinterpret_cast<const uint32_t *>(packed + (size_t)(v0 + v) * pitch + b0 + 4 * wq);
        tile[v][wq] = w;
    }
    __syncthreads();
    const int s = threadIdx.x;                              // sample inside the tile
    const int64_t n = b0 * 4 + s;
    if (n >= n_rows_t) return;
    const int wq = s >> 4, sh = 2 * (s & 15);
    uint32_t o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) acc |= ((tile[16 * k + i][wq] >> sh) & 3u) << (2 * i);
        o[k] = acc;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(pt + (size_t)n * pitch_t + (size_t)v0 / 4);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---- per-column scalars of b: max|b|, sum(b), quantisation unit and shift (one block per column, fixed tree) -----------------
// scal layout per column: [0] max|b|, [1] sum b, [2] unit_b, [3] shift_b, [4] max|e|, [5] H, [6] unit_e, [7] shift_e
constexpr int kUScal = 8;
__device__ __forceinline__ int quant_shift6(double maxabs, double *unit) {
    if (!(maxabs > 0) || !isfinite(maxabs)) {
        *unit = (maxabs == 0) ? 0.0 : __longlong_as_double(0x7ff8000000000000LL);
        return 0;
    }
    const int sh = 45 - ilogb(maxabs);          // |v| 2^sh < 2^46: the top digit stays within [-64, 64]
    *unit = scalbn(1.0, -sh);
    return sh;
}
// 46-bit fixed point -> six signed base-256 digits (d_l in [-128, 127], the top digit takes the rest)
__device__ __forceinline__ void to_digits6(double v, int shift, int8_t (&d)[kUND]) {
    long long B = __double2ll_rn(scalbn(v, shift));
#pragma unroll
    for (int l = 0; l < kUND - 1; l++) {
        const int dl = (int)(((B & 255) ^ 128) - 128);
        d[l] = (int8_t)dl;
        B = (B - dl) >> 8;
    }
    d[kUND - 1] = (int8_t)B;
}
__global__ void __launch_bounds__(1024) umma_colstats_kernel(const double *__restrict__ b, int64_t N, double *__restrict__ scal) {
    __shared__ double sm[32];
    const double *col = b + (size_t)blockIdx.x * N;
    double mx = 0, s = 0;
    for (int64_t i = threadIdx.x; i < N; i += 1024) { const double v = col[i]; mx = fmax(mx, fabs(v)); s += v; }
    mx = block_reduce_max<1024>(mx, sm);
    s = block_reduce_sum<1024>(s, sm);
    if (threadIdx.x == 0) {
        if (!isfinite(s)) mx = s;            // NaN / Inf anywhere in b poisons the product like it does in the reference
        double unit;
        const int sh = quant_shift6(mx, &unit);
        double *o = scal + (size_t)blockIdx.x * kUScal;
        o[0] = mx; o[1] = s; o[2] = unit; o[3] = (double)sh;
    }
}
// digits of K columns in MMA-B order: D[(c * 6 + l)][perm(i)], perm(i) = 16 (i / 16) + 4 (i % 4) + (i % 16) / 4.
// sel = 0: quantise with the b scalars (slots 2, 3), sel = 1: with the e scalars (slots 6, 7).  grid (ceil(Cpad / 256), K)
__global__ void umma_digits_kernel(const double *__restrict__ v, int64_t n_valid, int64_t ldv, int64_t cpad, const double *__restrict__ scal,
                                   int sel, int8_t *__restrict__ D) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cpad) return;
    const int c = blockIdx.y;
    const double unit = scal[(size_t)c * kUScal + (sel ? 6 : 2)];
    const int sh = (int)scal[(size_t)c * kUScal + (sel ? 7 : 3)];
    int8_t d[kUND];
    const double x = (i < n_valid && unit > 0) ? v[(size_t)c * ldv + i] : 0.0;
    to_digits6(x, sh, d);
    const int64_t p = (i & ~(int64_t)15) + 4 * (i & 3) + ((i & 15) >> 2);
#pragma unroll
    for (int l = 0; l < kUND; l++) D[(size_t)(c * kUND + l) * cpad + p] = d[l];
}

// exact column totals of the quantised values in limb form (what a GEMM against an all-ones operand would return): tot[c] = {sum of
// d0 + 256 d1 + 65536 d2, sum of d3 + 256 d4 + 65536 d5} over i < n.  One block of 1024 threads per column, integer arithmetic.
__global__ void __launch_bounds__(1024) umma_digit_totals_kernel(const double *__restrict__ v, int64_t n, int64_t ldv, const double *__restrict__ scal,
                                                                 long long *__restrict__ tot) {
    __shared__ long long sm[2][32];
    const int c = blockIdx.x;
    const double unit = scal[(size_t)c * kUScal + 2];
    const int sh = (int)scal[(size_t)c * kUScal + 3];
    long long lo = 0, hi = 0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        int8_t d[kUND];
        to_digits6(unit > 0 ? v[(size_t)c * ldv + i] : 0.0, sh, d);
        lo += (long long)d[0] + 256ll * d[1] + 65536ll * d[2];
        hi += (long long)d[3] + 256ll * d[4] + 65536ll * d[5];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = lo; sm[1][threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = 0, b = 0;
        for (int w = 0; w < 32; w++) { a += sm[0][w]; b += sm[1][w]; }
        tot[2 * c] = a; tot[2 * c + 1] = b;
    }
}

// after phase A: T' limbs -> dot, e, hm per (variant, column); per-column max|e| and H = sum h (deterministic).  grid (G, K)
__global__ void __launch_bounds__(256) umma_finalize_kernel(const unsigned long long *__restrict__ t_lo, const unsigned long long *__restrict__ t_hi,
                                                            const double *__restrict__ upart, int n_utiles, const double *__restrict__ lut,
                                                            int64_t M, double inv_mtotal, double *__restrict__ e, double *__restrict__ hm,
                                                            double *partial, unsigned int *counter, double *scal) {
    __shared__ double sm[8];
    __shared__ bool last;
    const int c = blockIdx.y, G = gridDim.x;
    double *sc = scal + (size_t)c * kUScal;
    const double unit_b = sc[2], sumb = sc[1];
    double mx = 0, hs = 0;
    for (int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x; j < M; j += (int64_t)G * 256) {
        const long long lo = (long long)t_lo[(size_t)c * M + j], hi = (long long)t_hi[(size_t)c * M + j];
        const double tsum = (double)lo + 16777216.0 * (double)hi;        // exact integer (256^3 = 2^24), one rounding
        double uj = 0;
        for (int t = 0; t < n_utiles; t++) uj += upart[((size_t)c * n_utiles + t) * M + j];
        const double l0 = lut[4 * j], inv = lut[4 * j + 1] - l0;
        const double T = (unit_b == 0 ? 0.0 : unit_b * tsum) - 3.0 * uj;
        const double dot = inv * T + l0 * (sumb - uj);
        const double ej = dot * inv * inv_mtotal, hj = dot * l0 * inv_mtotal;
        e[(size_t)c * M + j] = ej;
        hm[(size_t)c * M + j] = hj + 3.0 * ej;
        mx = fmax(mx, fabs(ej));
        hs += hj;
        if (!isfinite(ej)) mx = ej;
    }
    mx = block_reduce_max<256>(mx, sm);
    hs = block_reduce_sum<256>(hs, sm);
    double *pp = partial + (size_t)c * 2 * G;
    if (threadIdx.x == 0) {
        pp[blockIdx.x] = mx;
        pp[G + blockIdx.x] = hs;
        __threadfence();
        last = atomicInc(counter + c, G - 1) == (unsigned)(G - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const volatile double *p = pp;
        double m = 0, t = 0;
        for (int i = 0; i < G; i++) { m = fmax(m, p[i]); t += p[G + i]; }
        if (!isfinite(t)) m = t;
        double unit;
        const int sh = quant_shift6(m, &unit);
        sc[4] = m; sc[5] = t; sc[6] = unit; sc[7] = (double)sh;
    }
}

// out[c][n] = unit_e R[n][c] + H_c - corr[c][n].  grid (ceil(N / 256), K)
__global__ void umma_combine_kernel(const unsigned long long *__restrict__ r_lo, const unsigned long long *__restrict__ r_hi, int64_t N,
                                    const double *__restrict__ cpart, int n_ctiles, const double *__restrict__ scal,
                                    double *__restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int c = blockIdx.y;
    const double unit_e = scal[(size_t)c * kUScal + 6], H = scal[(size_t)c * kUScal + 5];
    const long long lo = (long long)r_lo[(size_t)c * N + n], hi = (long long)r_hi[(size_t)c * N + n];
    const double r = (double)lo + 16777216.0 * (double)hi;
    double corr = 0;
    for (int t = 0; t < n_ctiles; t++) corr += cpart[((size_t)c * n_ctiles + t) * N + n];
    out[(size_t)c * N + n] = (unit_e == 0 ? 0.0 : unit_e * r) + H - corr;
}

// ---- few right-hand sides: the same exact integer GEMM on mma.sync (legacy IMMA) -----------------------------------------------------
// A tcgen05.mma costs its ~96 clk dispatch whatever N, so with 3-4 right-hand sides (18-24 digit columns of 192) a phase of the
// kernels above takes 3.5-3.9 ms -- as long as with 30 columns.  mma.sync.m16n8k32 has N = 8: 24 digit columns are three
// instructions per A fragment, i.e. three times the tensor work of the single-RHS two-pass kernels, which is about their HBM time.
// Same inputs (row-major 2-bit matrix, digit matrix in the permuted element order) and same outputs (exact 64-bit limbs added with
// red.add.u64) as umma_gemm_kernel, so prepare / finalise / combine and the results do not change by a bit.
//   CTA = 8 warps x 2 row-blocks of 16 rows = 256 rows x a contraction range, two CTAs per SM; 2-stage cp.async ring; a stage = 2 K-steps of 64
//   packed bytes (256 elements) per row + the 24 digit rows of those 512 elements.  Per K-step a warp loads its digit fragments once
//   (3 n-tiles x 64 B per lane) and uses them for both row-blocks: per row-block and 2-bit plane t four LOP3 (bytes & 3 << 2t, the
//   products carry 4^t) feed six IMMAs.  A split holds at most 2,048 K-steps (int32 safe).
#ifndef SGB_SG_WARPS
#define SGB_SG_WARPS 8
#endif
#ifndef SGB_SG_STAGES
#define SGB_SG_STAGES 2
#endif
constexpr int kSgWarps = SGB_SG_WARPS, kSgRB = 2, kSgNT = 3;
constexpr int kSgThreads = kSgWarps * 32;
constexpr int kSgRows = kSgWarps * kSgRB * 16;          // 256
constexpr int kSgStageSteps = 2, kSgStages = SGB_SG_STAGES;
static_assert(kSgWarps == 8 && (kSgThreads / 8) % 2 == 0, "copy mapping: one digit row per warp and n-tile");
constexpr int kSgARow = 64 * kSgStageSteps;             // packed bytes per row and stage
constexpr int kSgDRow = 256 * kSgStageSteps + 16;       // digit bytes per digit row and stage (+ one granule: rows start in different bank groups)
constexpr int kSgStageBytes = kSgRows * kSgARow + 8 * kSgNT * kSgDRow;
constexpr int kSgSmemBytes = kSgStages * kSgStageBytes;
constexpr int kSgMaxSteps = 2048;                       // K-steps per split: 128 terms of at most 48 * 128 per K-step and accumulator
static_assert(kSgSmemBytes <= 227 * 1024, "small-K GEMM shared memory");
static_assert(kSgRows * 25 * 4 <= kSgSmemBytes, "epilogue staging");

template <int NT>   // n-tiles of eight digit columns: 6 K <= 8 NT
__global__ void __launch_bounds__(kSgThreads, kSgThreads <= 256 ? 2 : 1) imma_small_gemm_kernel(const uint8_t *__restrict__ P, size_t pitch, int64_t R, int64_t ksteps,
                                                                        int split, const int8_t *__restrict__ D, int64_t cpad, int ncols,
                                                                        unsigned long long *__restrict__ out_lo,
                                                                        unsigned long long *__restrict__ out_hi, int64_t ldo) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int64_t r0cta = (int64_t)blockIdx.x * kSgRows;
    const int ks_per = (int)((ksteps + split - 1) / split);
    const int ks_begin = (int)blockIdx.y * ks_per, ks_end = min((int)ksteps, ks_begin + ks_per);
    const int n_stage = (ks_end > ks_begin) ? (ks_end - ks_begin + kSgStageSteps - 1) / kSgStageSteps : 0;

    // copies of one stage: per thread 8 granules of packed rows and 3 of digit rows; everything but the K-step is fixed per thread
    // (row pointers and shared-memory offsets are computed once: the address arithmetic of a naive loop costs more issue slots than
    // the tensor-core instructions of the stage)
    const int a_r = tid >> 3, a_gran = tid & 7;
    unsigned a_row[8];                                     // 32-bit row numbers (a pointer per copy would cost eight more registers)
#pragma unroll
    for (int k = 0; k < 8; k++) a_row[k] = (unsigned)min(r0cta + a_r + (kSgThreads / 8) * k, R - 1);
    const uint8_t *a_base = P + (size_t)(a_gran & 3) * 16;
    // odd rows: the two K-steps swap halves of the 128-byte line, so that rows g and g + 1 of a fragment load use different banks
    const unsigned a_dst = smem_u32(smem) + a_r * kSgARow + ((a_gran ^ ((a_r & 1) << 2)) << 4);
    const int d_n = tid >> 5, d_gran = tid & 31;
    constexpr int kDCopies = 8 * NT / kSgWarps;
    const int8_t *d_base = D + (size_t)d_n * cpad + (size_t)(d_gran & 15) * 16;
    // granule index ^ 2 in the upper half of every 16: lanes tq = 0..3 of one fragment load hit four different bank groups
    const unsigned d_dst = smem_u32(smem) + kSgRows * kSgARow + d_n * kSgDRow + ((d_gran ^ (((d_gran >> 3) & 1) << 1)) << 4);
    const int ks_last = (int)ksteps - 1;
    auto cp16 = [](unsigned dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); };
    auto issue = [&](int st) {
        if (st < n_stage) {
            const unsigned so = (unsigned)(st % kSgStages) * kSgStageBytes;
            const int ks0 = ks_begin + st * kSgStageSteps;
            const size_t ka = (size_t)min(ks0 + (a_gran >> 2), ks_last) * 64, kd = (size_t)min(ks0 + (d_gran >> 4), ks_last) * 256;
#pragma unroll
            for (int k = 0; k < 8; k++) cp16(a_dst + so + k * (kSgThreads / 8) * kSgARow, a_base + (size_t)a_row[k] * pitch + ka);
#pragma unroll
            for (int k = 0; k < kDCopies; k++) cp16(d_dst + so + k * kSgWarps * kSgDRow, d_base + (size_t)(k * kSgWarps) * cpad + kd);
        }
        cp_async_commit();
    };

    // accumulators [row-block][plane pair][n-tile]: pair 0 = planes 0, 1 (products at scale 1), pair 1 = planes 2, 3 (scale 16)
    int acc[kSgRB][2][NT][4];
#pragma unroll
    for (int rb = 0; rb < kSgRB; rb++)
#pragma unroll
        for (int pp = 0; pp < 2; pp++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[rb][pp][nt][q] = 0;

    for (int s = 0; s < kSgStages - 1; s++) issue(s);
    for (int st = 0; st < n_stage; st++) {
        cp_async_wait<kSgStages - 2>();
        __syncthreads();
        issue(st + kSgStages - 1);
        const uint8_t *buf = smem + (size_t)(st % kSgStages) * kSgStageBytes;
        const uint8_t *dbuf = buf + kSgRows * kSgARow;
        const int ks0 = ks_begin + st * kSgStageSteps;
#pragma unroll
        for (int s = 0; s < kSgStageSteps; s++) {
            if (ks0 + s >= ks_end) break;
            // digit fragments of this lane: digit row 8 nt + g, granules 4 tq + j of the K-step; word t of granule j = plane t of the
            // packed bytes 16 tq + 4 j .. + 3.  One IMMA takes the four packed bytes of word j of this lane's rows twice: plane 2 pp in
            // its first k-half, plane 2 pp + 1 (the word shifted right by two bits, same mask, hence the same scale) in its second --
            // the B operand is then the adjacent word pair (2 pp, 2 pp + 1) of ONE granule, i.e. a register pair straight out of the
            // 16-byte load (no moves), and planes 2 pp and 2 pp + 1 share an accumulator
            uint32_t wa[kSgRB][4], wb[kSgRB][4];
#pragma unroll
            for (int rb = 0; rb < kSgRB; rb++) {
                const int r0 = warp * (kSgRB * 16) + rb * 16 + g;
                const int gran = (s * 4 + tq) ^ ((r0 & 1) << 2);     // r0 and r0 + 8 have the same parity
                const uint4 wa4 = *reinterpret_cast<const uint4 *>(buf + r0 * kSgARow + (gran << 4));
                const uint4 wb4 = *reinterpret_cast<const uint4 *>(buf + (r0 + 8) * kSgARow + (gran << 4));
                wa[rb][0] = wa4.x; wa[rb][1] = wa4.y; wa[rb][2] = wa4.z; wa[rb][3] = wa4.w;
                wb[rb][0] = wb4.x; wb[rb][1] = wb4.y; wb[rb][2] = wb4.z; wb[rb][3] = wb4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // digit granule j of the three n-tiles (12 registers live at a time), used by both row-blocks
                uint4 bf[NT];
                const int gran = s * 16 + 4 * tq + j;
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
                    bf[nt] = *reinterpret_cast<const uint4 *>(dbuf + (nt * 8 + g) * kSgDRow + ((gran ^ (((gran >> 3) & 1) << 1)) << 4));
#pragma unroll
                for (int rb = 0; rb < kSgRB; rb++) {
                    const uint32_t sa = wa[rb][j] >> 2, sb = wb[rb][j] >> 2;
#pragma unroll
                    for (int pp = 0; pp < 2; pp++) {
                        const uint32_t m = pp ? 0x30303030u : 0x03030303u;
                        const uint32_t a0 = wa[rb][j] & m, a1 = wb[rb][j] & m, a2 = sa & m, a3 = sb & m;
#pragma unroll
                        for (int nt = 0; nt < NT; nt++) {
                            const uint32_t b0 = pp ? bf[nt].z : bf[nt].x, b1 = pp ? bf[nt].w : bf[nt].y;
                            imma_nv(acc[rb][pp][nt], a0, a1, a2, a3, b0, b1);
                        }
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // epilogue: plane factors removed (exact shifts), the 24 digit-column sums of a row through shared memory ([row][25] ints), then
    // one lane per (row, right-hand side, limb): sum_k S[3 limb + k] 256^k added to the limb plane
    int *stg = reinterpret_cast<int *>(smem);
#pragma unroll
    for (int rb = 0; rb < kSgRB; rb++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v = acc[rb][0][nt][q] + (acc[rb][1][nt][q] >> 4);
                const int row = warp * (kSgRB * 16) + rb * 16 + g + (q >> 1) * 8, n = nt * 8 + 2 * tq + (q & 1);
                stg[row * 25 + n] = v;
            }
    __syncthreads();
    for (int i = tid; i < kSgRows * 8; i += kSgThreads) {
        const int row = i >> 3, slot = i & 7, cidx = slot >> 1, limb = slot & 1;
        const int64_t rr = r0cta + row;
        if (rr >= R || cidx >= ncols || n_stage == 0) continue;
        const int *sp = stg + row * 25 + cidx * kUND + 3 * limb;
        const long long v = (long long)sp[0] + 256ll * sp[1] + 65536ll * sp[2];
        red_add_u64((limb ? out_hi : out_lo) + (size_t)cidx * ldo + rr, (unsigned long long)v);
    }
}

// ---- multi-column variant of sparse_ell_sum_kernel: part[c][t][r] = sum over the entries of row r in tile t of vec[c][column] ----
// Same lane-interleaved index blocks; an item is 16 groups (512 rows) so that C vector tiles and two index stages fit in
// shared memory; every index is fetched once per C gathers.
constexpr int kEmGroups = 16;
constexpr int kEmThreads = 512;                 // 16 warps x 1 group
constexpr int kEmCap = 20480;                   // index entries per stage (40 KB); an item holds ~512 x 30 at 0.5 % missing
template <int C> constexpr int em_smem() { return C * kEllSv * 8 + 2 * kEmCap * 2 + 2 * 24 * 8; }
template <int C>
__global__ void __launch_bounds__(kEmThreads) sparse_ell_multi_kernel(const int64_t *__restrict__ gstart, const uint16_t *__restrict__ ell,
                                                                      const double *__restrict__ vec, int64_t ldv, int ncols, int64_t R,
                                                                      int64_t Cn, int64_t G, int n_tiles, double *__restrict__ part) {
    extern __shared__ __align__(16) uint8_t smem_sp[];
    double *sv = reinterpret_cast<double *>(smem_sp);                                       // [C][kEllSv]
    uint16_t *sidx0 = reinterpret_cast<uint16_t *>(smem_sp + C * kEllSv * 8);
    int64_t *sgs0 = reinterpret_cast<int64_t *>(smem_sp + C * kEllSv * 8 + 2 * kEmCap * 2);   // 2 x 24 group starts
    __shared__ unsigned long long s_bar[2];
    unsigned ph[2] = {0, 0};
    if (threadIdx.x == 0) {
        sp_mbar_init(&s_bar[0], 1);
        sp_mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n_chunks = (G + kEmGroups - 1) / kEmGroups;
    const int64_t n_work = n_chunks * n_tiles;
    const int64_t w_begin = n_work * blockIdx.x / gridDim.x, w_end = n_work * (blockIdx.x + 1) / gridDim.x;
    if (threadIdx.x < 8 * C) sv[(threadIdx.x >> 3) * kEllSv + kSpTile + (threadIdx.x & 7)] = 0.0;      // the zero slots
    auto bounds = [&](int64_t w, int64_t &e0, int64_t &e1) {
        const int64_t t = w / n_chunks, chunk = w % n_chunks;
        e0 = gstart[t * G + chunk * kEmGroups];
        e1 = gstart[t * G + min(G, (chunk + 1) * kEmGroups)];
    };
    auto issue = [&](int64_t w, int stage, int64_t e0, int64_t e1, bool with_sv) {
        const int64_t t = w / n_chunks, chunk = w % n_chunks, g0 = chunk * kEmGroups, c0 = t * kSpTile;
        const int ng = (int)(min(G, g0 + kEmGroups) - g0);
        const unsigned nb_gs = (unsigned)(((ng + 2) & ~1) * 8);
        const unsigned nb_idx = (unsigned)(min(e1 - e0, (int64_t)kEmCap) * 2);
        const int ncol = (int)min((int64_t)kSpTile, Cn - c0);
        const unsigned nb_sv = with_sv ? (unsigned)ncol * 8u : 0u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        sp_mbar_expect_tx(&s_bar[stage], nb_gs + nb_idx + nb_sv * (unsigned)C);
        sp_bulk_g2s(sgs0 + stage * 24, gstart + t * G + g0, nb_gs, &s_bar[stage]);
        if (nb_idx) sp_bulk_g2s(sidx0 + (size_t)stage * kEmCap, ell + e0, nb_idx, &s_bar[stage]);
        if (nb_sv)
            for (int cc = 0; cc < C; cc++)
                sp_bulk_g2s(sv + cc * kEllSv, vec + (size_t)min(cc, ncols - 1) * ldv + c0, nb_sv, &s_bar[stage]);
    };
    int64_t nx0 = 0, nx1 = 0;
    if (threadIdx.x == 0 && w_begin < w_end) bounds(w_begin, nx0, nx1);
    int cur_tile = -1;
    int it = 0;
    for (int64_t w = w_begin; w < w_end; w++, it++) {
        const int stage = it & 1;
        const int t = (int)(w / n_chunks);
        const int64_t chunk = w % n_chunks, g0 = chunk * kEmGroups, c0 = (int64_t)t * kSpTile;
        const int ng = (int)(min(G, g0 + kEmGroups) - g0);
        const bool new_tile = t != cur_tile;
        cur_tile = t;
        __syncthreads();
        if (new_tile) {
            const int ncol = (int)min((int64_t)kSpTile, Cn - c0);
            const bool sv_bulk = ((ncol & 1) == 0) && ((reinterpret_cast<uintptr_t>(vec + c0) & 15) == 0) && ((ldv & 1) == 0);
            if (!sv_bulk)
                for (int cc = 0; cc < C; cc++)
                    for (int i = threadIdx.x; i < ncol; i += kEmThreads) cp_async8(sv + cc * kEllSv + i, vec + (size_t)min(cc, ncols - 1) * ldv + c0 + i);
            for (int cc = 0; cc < C; cc++)
                for (int i = ncol + threadIdx.x; i < kSpTile; i += kEmThreads) sv[cc * kEllSv + i] = 0.0;
            cp_async_commit();
            if (threadIdx.x == 0) {
                issue(w, stage, nx0, nx1, sv_bulk);
                if (w + 1 < w_end) bounds(w + 1, nx0, nx1);
            }
            cp_async_wait<0>();
        }
        if (threadIdx.x == 0 && w + 1 < w_end && (int)((w + 1) / n_chunks) == t) {
            issue(w + 1, stage ^ 1, nx0, nx1, false);
            if (w + 2 < w_end) bounds(w + 2, nx0, nx1);
        }
        sp_mbar_wait(&s_bar[stage], ph[stage]);
        ph[stage] ^= 1;
        if (new_tile) __syncthreads();
        const uint16_t *sidx = sidx0 + (size_t)stage * kEmCap;
        const int64_t *sgs = sgs0 + stage * 24;
        const int64_t e0 = sgs[0], e1 = sgs[ng];
        double acc[C];
#pragma unroll
        for (int cc = 0; cc < C; cc++) acc[cc] = 0;
        for (int64_t pc = e0;; pc += kEmCap) {
            const int64_t pe = min(e1, pc + kEmCap);
            if (pc != e0) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    sp_mbar_expect_tx(&s_bar[stage], (unsigned)((pe - pc) * 2));
                    sp_bulk_g2s(sidx0 + (size_t)stage * kEmCap, ell + pc, (unsigned)((pe - pc) * 2), &s_bar[stage]);
                }
                sp_mbar_wait(&s_bar[stage], ph[stage]);
                ph[stage] ^= 1;
            }
            if (warp < ng) {
                const int lo = (int)(max(sgs[warp], pc) - pc), hi = (int)(min(sgs[warp + 1], pe) - pc);
                double s0[C], s1[C];
#pragma unroll
                for (int cc = 0; cc < C; cc++) s0[cc] = s1[cc] = 0;
                int i = lo + lane;
                for (; i + 32 < hi; i += 64) {
                    const int a = sidx[i], b = sidx[i + 32];
#pragma unroll
                    for (int cc = 0; cc < C; cc++) { s0[cc] += sv[cc * kEllSv + a]; s1[cc] += sv[cc * kEllSv + b]; }
                }
                for (; i < hi; i += 32) {
                    const int a = sidx[i];
#pragma unroll
                    for (int cc = 0; cc < C; cc++) s0[cc] += sv[cc * kEllSv + a];
                }
#pragma unroll
                for (int cc = 0; cc < C; cc++) acc[cc] += s0[cc] + s1[cc];
            }
            if (pe >= e1) break;
        }
        const int64_t r = (g0 + warp) * 32 + lane;
        if (warp < ng && r < R) {
#pragma unroll
            for (int cc = 0; cc < C; cc++)
                if (cc < ncols) part[((size_t)cc * n_tiles + t) * R + r] = acc[cc];
        }
    }
}

// ---- row-gather variant for many columns: out[c][r] = sum over the entries of row r (CSR, 32-bit column ids) of vt[id][c] ----------
// vt is the vector block transposed to [column id][W] (W = 16 or 32 doubles = 128 / 256 B per id, zero padded), which stays in L2
// (55 / 110 MB at N = 430K): a sub-warp of W lanes walks one row, lane <-> right-hand side, every entry is one coalesced W * 8 byte
// load -- no shared-memory bank conflicts, no partial sums, the summation order of a row is its list order (deterministic).
template <int W>
__global__ void __launch_bounds__(256) sparse_rows_gather_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                                 const double *__restrict__ vt, int64_t zero_row, int ncols, int64_t R,
                                                                 double *__restrict__ out, int64_t ldo) {
    constexpr int RPW = 32 / W;                       // rows per warp
    constexpr int NB = W < 16 ? W : 16;               // loads in flight per lane: the gather is bound by L2 latency x occupancy
    const int lane = threadIdx.x & 31, sub = lane / W, col = lane % W;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double *vcol = vt + col;
    for (int64_t rb = warp0 * RPW; rb < R; rb += n_warps * RPW) {
        const int64_t r = rb + sub;
        int64_t e0 = 0, e1 = 0;
        if (r < R) { e0 = ptr[r]; e1 = ptr[r + 1]; }
        int64_t len = e1 - e0, maxlen = len;
        if (RPW > 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 16));
        if (RPW > 2) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, 8));
        double acc = 0;
        for (int64_t base = 0; base < maxlen; base += W) {
            const int64_t me = base + col;
            const int64_t my = (me < len) ? (int64_t)idx[e0 + me] : zero_row;     // padding reads the all-zero row: no predicates below
            const int cnt = (int)min((int64_t)W, maxlen - base);
            for (int t0 = 0; t0 < cnt; t0 += NB) {
                double v[NB];
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int64_t i = __shfl_sync(0xffffffffu, (t0 + u < W) ? my : zero_row, (t0 + u) & (W - 1), W);
                    v[u] = __ldg(vcol + (size_t)((t0 + u < cnt) ? i : zero_row) * W);
                }
#pragma unroll
                for (int u = 0; u < NB; u++) acc += v[u];                        // list order: deterministic
            }
        }
        if (r < R && col < ncols) out[(size_t)col * ldo + r] = acc;
    }
}
// Same walk with 16-byte loads: a sub-warp of W / 2 lanes per row, lane <-> two right-hand sides, so one load instruction of a warp
// covers 64 / W rows (half the load instructions per byte of the kernel above; env SGB_UMMA_GATHER_V2).
template <int W, int NBX = 16>
__global__ void __launch_bounds__(256) sparse_rows_gather2_kernel(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                                                                  const double *__restrict__ vt, int64_t zero_row, int ncols, int64_t R,
                                                                  double *__restrict__ out, int64_t ldo) {
    constexpr int L = W / 2;                          // lanes per row
    constexpr int RPW = 32 / L;                       // rows per warp
    constexpr int NB = L < NBX ? L : NBX;             // loads in flight per lane
    const int lane = threadIdx.x & 31, sub = lane / L, cl = lane % L;
    const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const double2 *vcol = reinterpret_cast<const double2 *>(vt) + cl;
    for (int64_t rb = warp0 * RPW; rb < R; rb += n_warps * RPW) {
        const int64_t r = rb + sub;
        int64_t e0 = 0, e1 = 0;
        if (r < R) { e0 = ptr[r]; e1 = ptr[r + 1]; }
        int64_t len = e1 - e0, maxlen = len;
#pragma unroll
        for (int o = 16; o >= L; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
        double2 acc = make_double2(0.0, 0.0);
        for (int64_t base = 0; base < maxlen; base += L) {
            const int64_t me = base + cl;
            const int64_t my = (me < len) ? (int64_t)idx[e0 + me] : zero_row;     // padding reads the all-zero row: no predicates below
            const int cnt = (int)min((int64_t)L, maxlen - base);
            for (int t0 = 0; t0 < cnt; t0 += NB) {
                double2 v[NB];
#pragma unroll
                for (int u = 0; u < NB; u++) {
                    const int64_t i = __shfl_sync(0xffffffffu, (t0 + u < L) ? my : zero_row, (t0 + u) & (L - 1), L);
                    v[u] = __ldg(vcol + (size_t)((t0 + u < cnt) ? i : zero_row) * L);
                }
#pragma unroll
                for (int u = 0; u < NB; u++) { acc.x += v[u].x; acc.y += v[u].y; }   // list order: deterministic, same sums as above
            }
        }
        if (r < R) {
            if (2 * cl < ncols) out[(size_t)(2 * cl) * ldo + r] = acc.x;
            if (2 * cl + 1 < ncols) out[(size_t)(2 * cl + 1) * ldo + r] = acc.y;
        }
    }
}
// dst[i][c] = src[c][i] for c < ncols, 0 for ncols <= c < W; row n is the all-zero row.  grid ceil((n + 1) / 256)
template <int W>
__global__ void transpose_cols_kernel(const double *__restrict__ src, int64_t ld, int ncols, int64_t n, double *__restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    double v[W];
#pragma unroll
    for (int c = 0; c < W; c++) v[c] = (c < ncols && i < n) ? src[(size_t)c * ld + i] : 0.0;
    double2 *d = reinterpret_cast<double2 *>(dst + (size_t)i * W);
#pragma unroll
    for (int c = 0; c < W; c += 2) d[c >> 1] = make_double2(v[c], v[c + 1]);
}

// ---- pair kernel: cta_group::2 MMAs, one M = 128 tile per CTA, eight A slots ------------------------------------------------------------
// tools/umma_probe.cu (profiles/r02_umma_probe_pair_b200.txt): a tcgen05.mma of one CTA costs ~124 clk whatever N <= 240, a pair MMA
// (M = 256 over two SMs) 96.3 clk for N <= 192 = 8,167 MAC/clk/SM at N = 192, 99.7 % of the int8 peak.  The kernel above is therefore
// bound by MMA dispatch (8 x 124 clk per stage and SM).  Here a cluster of two CTAs shares every MMA: each CTA owns 128 rows (one
// M-tile, accumulator of 192 columns), which leaves 256 tensor-memory columns = EIGHT A-operand slots, so the cross-CTA hand-over loop
// (remote a_full arrivals, multicast commits) is hidden; each CTA loads half of every digit tile (rows [ng/2 rank, +ng/2)), which
// also halves the L2 -> SM traffic of the digit matrix.
//   warp 0: packed boxes [128 rows x 128 B] (4 stages each, ring of 4)     warp 2: digit half tiles (ring of 8)
//   warp 1: issuer (leader CTA only): per stage one wait, four pair MMAs, one multicast commit
//   warp 3: tensor-memory allocation (cta_group::2)                        warps 4..19: four expander sets, set s serves st % 4 == s
constexpr int kPRows = 128, kPNP = 4, kPNB = 8, kPNA = 8;
constexpr int kPIssuers = 4;            // issuing threads (leader CTA), stages round-robin: for ONE thread a tcgen05.mma costs ~100 clk of issue
                                        // latency and a multicast commit ~240 clk (tools/umma_probe.cu), a pair MMA of N = 192 keeps the pipe 96 clk
constexpr int kPThreads = (20 + kPIssuers) * 32;    // warp 0: packed boxes, 1 / 3: digit tiles, 2: tensor-memory allocation, 4..19: expanders, 20..: issuers
constexpr int kPPackBytes = kPRows * kUBoxBytes;      // 16 KB
constexpr int kPBMax = (kUMaxN / 2) * 128;            // 12 KB
struct PairSmem {
    unsigned long long p_full[kPNP], p_empty[kPNP], b_full[kPNB], a_full[kPNA], mma_done[kPNB], acc_full, first_issued;
    unsigned long long deadline;
    unsigned tmem_base;
};
constexpr int kPSmemBytes = kPNP * kPPackBytes + kPNB * kPBMax + 1024 + (int)sizeof(PairSmem);

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1) umma_pair_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                                                                          const __grid_constant__ CUtensorMap tmap_d, UmmaArgs A) {
    extern __shared__ uint8_t smem_dyn[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    uint8_t *pack = base;
    uint8_t *btile = base + kPNP * kPPackBytes;
    PairSmem &S = *reinterpret_cast<PairSmem *>(btile + kPNB * kPBMax);
    unsigned crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const unsigned bhalf_rows = (unsigned)A.ng / 2, btile_bytes = bhalf_rows * 128u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    volatile int *err = A.err;
    const int box0 = blockIdx.y * A.boxes_per_split;
    const int n_box = min(A.boxes_per_split, A.boxes_total - box0);
    const int n_st = n_box * 4;
    const int row0 = blockIdx.x * kPRows;
    const volatile unsigned long long *dl = &S.deadline;
    if (tid == 0) {
        S.deadline = global_ns() + g_wait_timeout_ns;
        for (int i = 0; i < kPNP; i++) { mbar_init(&S.p_full[i], 1); mbar_init(&S.p_empty[i], 16); }
        for (int i = 0; i < kPNB; i++) { mbar_init(&S.b_full[i], 1); mbar_init(&S.mma_done[i], 1); }
        for (int i = 0; i < kPNA; i++) mbar_init(&S.a_full[i], 8);          // four expander warps of each CTA of the pair
        mbar_init(&S.acc_full, kPIssuers);                                   // one commit per issuer
        mbar_init(&S.first_issued, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = S.tmem_base;
    constexpr unsigned kStage0 = 256;                       // A slots: columns 256 + 32 slot; accumulator: columns [0, ng)

    if (n_st > 0) {
    if (warp == 0) {
        if (lane == 0) {
            unsigned long long policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            for (int bx = 0; bx < n_box; bx++) {
                const int ps = bx % kPNP;
                if (bx >= kPNP && !mbar_wait(&S.p_empty[ps], (unsigned)(((bx / kPNP) - 1) & 1), err, dl)) break;
                mbar_expect_tx(&S.p_full[ps], kPPackBytes);
                tma_load_2d(pack + (size_t)ps * kPPackBytes, &tmap_p, (box0 + bx) * kUBoxBytes, row0, &S.p_full[ps], policy);
            }
        }
    } else if (warp == 1 || warp == 3) {
        // digit half tiles, two producer threads on alternating stages (one thread's wait + expect_tx + TMA issue is ~500 clk per stage)
        if (lane == 0) {
            for (int st = (warp == 1 ? 0 : 1); st < n_st; st += 2) {
                const int bs = st % kPNB;
                if (st >= kPNB && !mbar_wait(&S.mma_done[bs], (unsigned)(((st / kPNB) - 1) & 1), err, dl)) break;
                // both halves of the tile are counted on the LEADER's barrier (the issuers wait there); the leader expects all the bytes
                if (crank == 0) mbar_expect_tx(&S.b_full[bs], 2 * btile_bytes);
                tma_load_2d_pair(btile + (size_t)bs * kPBMax, &tmap_d, (box0 * 4 + st) * kUStage, (int)(crank * bhalf_rows), &S.b_full[bs], 0u);
            }
        }
    } else if (warp >= 20) {
        // two issuers (leader CTA), stages alternating: for the issuing thread a stage is a serial wait (~200-350 clk) + four pair MMAs
        // (~385 clk, the thread is held while the queue is full) + multicast commit; one thread alone leaves the tensor pipe idle half
        // of the time.  Integer accumulation commutes, so the interleaving of the two MMA streams on the one accumulator does not
        // matter -- except that the zero-initialising MMA (stage 0, k-step 0) must be first: issuer 1 waits for it to be issued.
        if (lane == 0 && crank == 0) {
            const int me = warp - 20;
            const unsigned idesc = umma_idesc_i8(256, A.ng);
            const unsigned bt0 = smem_u32(btile);
            long long c_wa = 0, c_mma = 0, c_cm = 0, t0 = 0, t1 = 0;
            const bool PROF = A.prof != nullptr;
            const long long t_begin = PROF ? clock64() : 0;
            for (int st = me; st < n_st; st += kPIssuers) {
                const int sl = st % kPNA;
                if (PROF) t0 = clock64();
                if (!mbar_wait(&S.a_full[sl], (unsigned)((st / kPNA) & 1), err, dl)) break;
                if (!mbar_wait(&S.b_full[st % kPNB], (unsigned)((st / kPNB) & 1), err, dl)) break;
                if (PROF) { t1 = clock64(); c_wa += t1 - t0; }
                if (st == me && me != 0 && !mbar_wait(&S.first_issued, 0u, err, dl)) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned bb = bt0 + (unsigned)(st % kPNB) * kPBMax;
#pragma unroll
                for (int ks = 0; ks < 4; ks++)
                    umma_i8_ts_pair(tb, tb + kStage0 + sl * 32 + ks * 8, umma_desc_sw128(bb + ks * 32), idesc, (st > 0 || ks > 0) ? 1u : 0u);
                if (st == 0) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&S.first_issued);
                }
                if (PROF) { t0 = clock64(); c_mma += t0 - t1; }
                umma_commit_arrive_pair(&S.mma_done[st % kPNB]);
                if (PROF) { t1 = clock64(); c_cm += t1 - t0; }
            }
            umma_commit_arrive_pair(&S.acc_full);
            if (PROF && me == 0) {
                long long *o = A.prof + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
                o[0] = clock64() - t_begin; o[1] = 0; o[2] = c_wa; o[3] = c_mma; o[4] = c_cm; o[5] = (n_st + kPIssuers - 1) / kPIssuers;
            }
        }
    } else if (warp >= 4 && warp < 20) {
        const int e = warp - 4, set = e >> 2, quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const unsigned lane_base = (unsigned)(quarter * 32) << 16;
        const unsigned prow = (unsigned)r * kUBoxBytes, psw = (unsigned)(r & 7);
        bool ok = true;
        long long c_wp = 0, c_ld = 0, c_we = 0, c_st = 0, c_ar = 0, t0 = 0, t1 = 0;
        const bool PROF = A.prof != nullptr;
        const long long t_begin = PROF ? clock64() : 0;
        for (int st = set; st < n_st && ok; st += 4) {
            const int bx = st >> 2, ps = bx % kPNP, sl = st % kPNA;
            if (PROF) t0 = clock64();
            ok = mbar_wait(&S.p_full[ps], (unsigned)((bx / kPNP) & 1), err, dl);
            if (!ok) break;
            if (PROF) { t1 = clock64(); c_wp += t1 - t0; }
            const uint8_t *src = pack + (size_t)ps * kPPackBytes + prow;
            const unsigned c0 = (unsigned)(st & 3) * 2;
            const uint4 wa = *reinterpret_cast<const uint4 *>(src + ((c0 ^ psw) << 4));
            const uint4 wb = *reinterpret_cast<const uint4 *>(src + (((c0 + 1) ^ psw) << 4));
            uint32_t ws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            if (A.amode != 0) {          // indicator operands: one bit per code, in the low bit of its 2-bit field
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t w = ws[i];
                    ws[i] = (A.amode == 1 ? w : (A.amode == 2 ? (w >> 1) : (w & (w >> 1)))) & 0x55555555u;
                }
            }
            uint32_t x[32];
#pragma unroll
            for (int i = 0; i < 8; i++)
#pragma unroll
                for (int t = 0; t < 4; t++) x[4 * i + t] = (ws[i] >> (2 * t)) & 0x03030303u;
            if (PROF) { t0 = clock64(); c_ld += t0 - t1; }
            if (st >= kPNA) ok = mbar_wait(&S.mma_done[(st - kPNA) % kPNB], (unsigned)(((st - kPNA) / kPNB) & 1), err, dl);
            if (!ok) break;
            if (PROF) { t1 = clock64(); c_we += t1 - t0; }
            tmem_st32(tb + lane_base + kStage0 + (unsigned)sl * 32u, x);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (PROF) { t0 = clock64(); c_st += t0 - t1; }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_cluster(&S.a_full[sl], 0u);
                mbar_arrive(&S.p_empty[ps]);
            }
            if (PROF) { t1 = clock64(); c_ar += t1 - t0; }
        }
        if (PROF && warp == 4 && lane == 0 && crank == 0) {
            long long *o = A.prof + (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16;
            o[8] = clock64() - t_begin; o[9] = c_wp; o[10] = c_ld; o[11] = c_we; o[12] = c_st; o[13] = c_ar; o[14] = (n_st + 3) / 4;
        }
        if (ok) ok = mbar_wait(&S.acc_full, 0u, err, dl);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            const int64_t row = (int64_t)row0 + r;
            for (int g = set; g * 8 < A.ncols; g += 4) {
                int v[8 * kUND];
#pragma unroll
                for (int i = 0; i < kUND; i++) tmem_ld8_nowait(tb + lane_base + g * (8 * kUND) + i * 8, v + 8 * i);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int cc = 0; cc < 8; cc++) {
                    const int col = g * 8 + cc;
                    long long lo = 0, hi = 0;
#pragma unroll
                    for (int l = 2; l >= 0; l--) lo = lo * 256 + v[cc * kUND + l];
#pragma unroll
                    for (int l = 5; l >= 3; l--) hi = hi * 256 + v[cc * kUND + l];
                    if (col < A.ncols && row < A.R) {
                        red_add_u64(A.out_lo + (size_t)col * A.ldo + row, (unsigned long long)lo);
                        red_add_u64(A.out_hi + (size_t)col * A.ldo + row, (unsigned long long)hi);
                    }
                }
            }
        }
    }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
    }
}
