// tools/umma_probe.cu -- correctness and throughput probe for tcgen05.mma kind::i8 (u8 x s8 -> s32) on sm_100a.
//
// Part 1 (layouts): one CTA multiplies A[128 x K] (u8) by B[N x K] (s8) for several shared-memory layouts / descriptor
// encodings and compares D = A B^T with the host.  It answers, on the hardware, the questions the batched GRM kernel
// (csrc/grm_umma.cuh) depends on: meaning of LBO / SBO for the no-swizzle K-major and MN-major canonical layouts, the
// 128-byte-swizzled K-major layout a TMA box produces, and A taken from tensor memory.
// Part 2 (throughput): MAC / clk / SM of back-to-back MMAs for N = 8 .. 256, operands resident, SS and TS mode.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    const unsigned a = smem_u32(b);
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long *b) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void umma_i8_ss(unsigned d_tmem, uint64_t adesc, uint64_t bdesc, unsigned idesc, unsigned accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_i8_ts(unsigned d_tmem, unsigned a_tmem, uint64_t bdesc, unsigned idesc, unsigned accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tmem_ld8(unsigned taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout << 61
__host__ __device__ inline uint64_t make_desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned layout) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::i8: D s32, A u8, B s8
__host__ __device__ inline unsigned make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((unsigned)a_mn_major << 15) | ((unsigned)b_mn_major << 16) |
           ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

enum { L_K_NOSW = 0, L_MN_NOSW = 1, L_K_SW128 = 2, L_TMEM = 3 };

struct ProbeArgs {
    int N, ksteps;           // K = 32 * ksteps
    int a_layout, b_layout;  // L_*
    int swap_lbo_sbo;        // encode (SBO, LBO) instead of (LBO, SBO): tells which reading of the canonical layout is right
    const uint8_t *A;        // [128][K] row-major (logical)
    const int8_t *B;         // [N][K]
    int *D;                  // [128][N]
};

// byte offset of logical element (r, k) of a [R x 32*ksteps] operand inside its tile, plus the descriptor parameters
struct LayoutParams { unsigned lbo, sbo, kstep_stride, layout_code; };
__host__ __device__ inline LayoutParams layout_params(int layout, int rows, int ksteps) {
    LayoutParams p{};
    if (layout == L_K_NOSW) {            // [row group][16-byte K chunk][8 rows][16 B]
        p.lbo = 128; p.sbo = 128u * 2 * ksteps; p.kstep_stride = 256; p.layout_code = 0;
    } else if (layout == L_MN_NOSW) {    // [16-row unit][K group of 8][8 k][16 B of rows]
        p.lbo = 128; p.sbo = 128u * 4 * ksteps; p.kstep_stride = 512; p.layout_code = 0;
    } else {                             // SW128 K-major: rows of 128 B (4 k-steps), 8-row atoms of 1 KB
        p.lbo = 16; p.sbo = 1024; p.kstep_stride = 32; p.layout_code = 2;
    }
    (void)rows;
    return p;
}
__host__ __device__ inline unsigned elem_off(int layout, int ksteps, int r, int k) {
    if (layout == L_K_NOSW) return (unsigned)((r >> 3) * (128 * 2 * ksteps) + (k >> 4) * 128 + (r & 7) * 16 + (k & 15));
    if (layout == L_MN_NOSW) return (unsigned)((r >> 4) * (128 * 4 * ksteps) + (k >> 3) * 128 + (k & 7) * 16 + (r & 15));
    // SW128: K <= 128 here
    return (unsigned)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 4) ^ (r & 7)) & 7) << 4) + (k & 15));
}

__global__ void __launch_bounds__(128, 1) probe_layout_kernel(ProbeArgs P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = 32 * P.ksteps;
    uint8_t *sA = smem, *sB = smem + 65536;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        if (P.a_layout != L_TMEM) sA[elem_off(P.a_layout, P.ksteps, r, k)] = P.A[i];
    }
    for (int i = tid; i < P.N * K; i += 128) {
        const int r = i / K, k = i % K;
        sB[elem_off(P.b_layout, P.ksteps, r, k)] = (uint8_t)P.B[i];
    }
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base;
    const unsigned a_tmem = tb + 256;          // columns 256.. hold A in TS mode (8 columns per k-step)
    if (P.a_layout == L_TMEM) {
        // thread <-> row: 32 bytes (8 columns) per k-step
        for (int s = 0; s < P.ksteps; s++) {
            uint32_t r[8];
            for (int q = 0; q < 8; q++) {
                uint32_t v = 0;
                for (int by = 0; by < 4; by++) v |= (uint32_t)P.A[(size_t)tid * K + s * 32 + q * 4 + by] << (8 * by);
                r[q] = v;
            }
            tmem_st8(a_tmem + ((unsigned)(warp * 32) << 16) + s * 8, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (tid == 0) {
        const LayoutParams la = layout_params(P.a_layout == L_TMEM ? L_K_NOSW : P.a_layout, 128, P.ksteps), lb = layout_params(P.b_layout, P.N, P.ksteps);
        const unsigned idesc = make_idesc(128, P.N, P.a_layout == L_MN_NOSW, P.b_layout == L_MN_NOSW);
        for (int s = 0; s < P.ksteps; s++) {
            const uint64_t ad = P.swap_lbo_sbo ? make_desc(smem_u32(sA) + s * la.kstep_stride, la.sbo, la.lbo, la.layout_code)
                                               : make_desc(smem_u32(sA) + s * la.kstep_stride, la.lbo, la.sbo, la.layout_code);
            const uint64_t bd = P.swap_lbo_sbo ? make_desc(smem_u32(sB) + s * lb.kstep_stride, lb.sbo, lb.lbo, lb.layout_code)
                                               : make_desc(smem_u32(sB) + s * lb.kstep_stride, lb.lbo, lb.sbo, lb.layout_code);
            if (P.a_layout == L_TMEM) umma_i8_ts(tb, a_tmem + s * 8, bd, idesc, s > 0);
            else umma_i8_ss(tb, ad, bd, idesc, s > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < P.N; c0 += 8) {
        uint32_t r[8];
        tmem_ld8(tb + ((unsigned)(warp * 32) << 16) + c0, r);
        for (int q = 0; q < 8; q++) P.D[(size_t)tid * P.N + c0 + q] = (int)r[q];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

// ---- throughput: `iters` x (4 k-steps of M=128 x N) into `nacc` accumulators in turn; operands resident -------------------
struct RateArgs { int N, iters, nacc, ts; long long *cycles; };
__global__ void __launch_bounds__(128, 1) probe_rate_kernel(RateArgs P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base;
    if (P.ts) {
        uint32_t r[8] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u};
        for (int s = 0; s < 4; s++) tmem_st8(tb + 480 + ((unsigned)(warp * 32) << 16) + s * 8, r);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t0 = 0;
    if (tid == 0) {
        const unsigned idesc = make_idesc(128, P.N, 0, 0);
        const unsigned sA = smem_u32(smem), sB = smem_u32(smem) + 32768;
        t0 = clock64();
        for (int it = 0; it < P.iters; it++) {
            const unsigned d = tb + (unsigned)((it % P.nacc) * P.N);
            for (int s = 0; s < 4; s++) {
                const uint64_t bd = make_desc(sB + s * 32, 16, 1024, 2);
                if (P.ts) umma_i8_ts(d, tb + 480 + s * 8, bd, idesc, 1);
                else umma_i8_ss(d, make_desc(sA + s * 256, 128, 1024, 0), bd, idesc, 1);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    if (tid == 0) P.cycles[blockIdx.x] = clock64() - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

static int run_layout(const char *name, int N, int ksteps, int a_layout, int b_layout, int swap) {
    const int K = 32 * ksteps;
    std::vector<uint8_t> A((size_t)128 * K);
    std::vector<int8_t> B((size_t)N * K);
    srand(1234 + N + a_layout * 7 + b_layout * 13);
    for (auto &v : A) v = (uint8_t)(rand() & 0xFF);
    for (auto &v : B) v = (int8_t)((rand() & 0xFF) - 128);
    std::vector<int> ref((size_t)128 * N), got((size_t)128 * N, -1);
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            int s = 0;
            for (int k = 0; k < K; k++) s += (int)A[(size_t)m * K + k] * (int)B[(size_t)n * K + k];
            ref[(size_t)m * N + n] = s;
        }
    uint8_t *dA; int8_t *dB; int *dD;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, got.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, got.size() * 4));
    ProbeArgs P{N, ksteps, a_layout, b_layout, swap, dA, dB, dD};
    CK(cudaFuncSetAttribute(probe_layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    probe_layout_kernel<<<1, 128, 131072>>>(P);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s : CUDA error %s\n", name, cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < got.size(); i++) bad += got[i] != ref[i];
    printf("%-44s N=%3d K=%3d swap=%d : %zu / %zu mismatches%s\n", name, N, K, swap, bad, got.size(), bad ? "" : "   <== OK");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad == 0;
}

int pair_main();
int main(int argc, char **argv) {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device: %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const bool rate_only = argc > 1 && !strcmp(argv[1], "rate");
    if (argc > 1 && !strcmp(argv[1], "pair")) return pair_main();
    if (!rate_only) {
        for (int N : {16, 240}) {
            for (int swap = 0; swap < 2; swap++) {
                run_layout("A K-major noswz, B K-major noswz", N, 4, L_K_NOSW, L_K_NOSW, swap);
                run_layout("A MN-major noswz, B K-major noswz", N, 4, L_MN_NOSW, L_K_NOSW, swap);
            }
            run_layout("A K-major noswz, B K-major SW128", N, 4, L_K_NOSW, L_K_SW128, 0);
            run_layout("A MN-major noswz, B K-major SW128", N, 4, L_MN_NOSW, L_K_SW128, 0);
            run_layout("A in TMEM, B K-major SW128", N, 4, L_TMEM, L_K_SW128, 0);
            run_layout("A K-major SW128, B K-major SW128", N, 4, L_K_SW128, L_K_SW128, 0);
        }
        run_layout("A K-major noswz, B K-major noswz (1 k-step)", 64, 1, L_K_NOSW, L_K_NOSW, 0);
        run_layout("A MN-major noswz, B MN-major noswz", 64, 2, L_MN_NOSW, L_MN_NOSW, 0);
    }
    // throughput
    long long *dcyc;
    const int G = prop.multiProcessorCount;
    CK(cudaMalloc(&dcyc, sizeof(long long) * G));
    CK(cudaFuncSetAttribute(probe_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    printf("\nthroughput (all %d SMs busy, M = 128, 4 k-steps of 32 per iteration, 2000 iterations):\n", G);
    for (int ts = 0; ts < 2; ts++)
        for (int N : {8, 16, 32, 64, 128, 240, 256}) {
            if (ts && N > 240) continue;       // columns 480.. hold A
            for (int nacc : {1, 2}) {
                if (nacc * N > (ts ? 480 : 512)) continue;
                RateArgs R{N, 2000, nacc, ts, dcyc};
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                probe_rate_kernel<<<G, 128, 98304>>>(R);       // warm-up
                CK(cudaEventRecord(e0));
                probe_rate_kernel<<<G, 128, 98304>>>(R);
                CK(cudaEventRecord(e1));
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("rate kernel: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                std::vector<long long> cyc(G);
                CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * G, cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (long long v : cyc) mx = v > mx ? v : mx;
                const double mac = 2000.0 * 4 * 128.0 * N * 32;
                printf("  %s N=%3d acc=%d : %8lld clk  -> %7.1f MAC/clk/SM, %6.1f clk per MMA (floor %5.1f), chip %6.1f TMAC/s (event %.3f ms)\n",
                       ts ? "TS" : "SS", N, nacc, mx, mac / (double)mx, (double)mx / 8000.0, 128.0 * N / 256.0,
                       mac * G / (ms * 1e-3) / 1e12, ms);
            }
        }
    return 0;
}

// ---- cta_group::2: one tcgen05.mma drives the tensor cores of both SMs of a CTA pair (M = 256: 128 rows per CTA) ---------------------
// Layout check (A from tensor memory, each CTA its own 128 rows; B split by rows, N / 2 per CTA, at the same shared-memory offset) and
// MAC/clk/SM of back-to-back pair MMAs: is the ~124 clk per instruction a per-instruction dispatch cost that a pair MMA halves per SM?
__device__ __forceinline__ void umma_i8_ts_2cta(unsigned d_tmem, unsigned a_tmem, uint64_t bdesc, unsigned idesc, unsigned accum) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, p;\n}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(unsigned long long *b) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(b)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
struct Pair2Args { int N, ksteps, iters; const uint8_t *A; const int8_t *B; int *D; long long *cycles; int commit_every; int st_during; };
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe_pair_kernel(Pair2Args P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long bar, bar2;
    __shared__ unsigned tmem_base;
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = 32 * P.ksteps, NH = P.N / 2;
    // this CTA's half of B: rows [NH rank, NH rank + NH), SW128 K-major
    for (int i = tid; i < NH * K; i += 128) {
        const int r = i / K, k = i % K;
        smem[elem_off(L_K_SW128, P.ksteps, r, k)] = P.B ? (uint8_t)P.B[(size_t)(NH * rank + r) * K + k] : (uint8_t)1;
    }
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tmem_base;
    // this CTA's 128 rows of A into its own tensor memory, columns 256.. (8 per k-step)
    for (int s = 0; s < P.ksteps; s++) {
        uint32_t r[8];
        for (int q = 0; q < 8; q++) {
            uint32_t v = 0x01010101u;
            if (P.A) { v = 0; for (int by = 0; by < 4; by++) v |= (uint32_t)P.A[(size_t)(128 * rank + tid) * K + s * 32 + q * 4 + by] << (8 * by); }
            r[q] = v;
        }
        tmem_st8(tb + 256 + ((unsigned)(warp * 32) << 16) + s * 8, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = 0;
    if (rank == 0 && tid == 0) {
        const unsigned idesc = make_idesc(256, P.N, 0, 0);
        t0 = clock64();
        for (int it = 0; it < P.iters; it++) {
            for (int s = 0; s < P.ksteps; s++)
                umma_i8_ts_2cta(tb, tb + 256 + s * 8, make_desc(smem_u32(smem) + s * 32, 16, 1024, 2), idesc, (it > 0 || s > 0) ? 1u : 0u);
            if (P.commit_every && (it % P.commit_every) == P.commit_every - 1) {
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar2)), "h"((unsigned short)3) : "memory");
            }
        }
        umma_commit_2cta(&bar);
    } else if (P.st_during && warp >= 1 && P.cycles) {
        // the other warps keep writing A-operand-sized blocks into unrelated tensor-memory columns while the MMAs run
        uint32_t r[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        for (int it = 0; it < P.iters * 2; it++) {
            tmem_st8(tb + 384 + ((unsigned)(warp * 32) << 16) + (it & 7) * 8, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
    }
    mbar_wait(&bar, 0);
    if (rank == 0 && tid == 0 && P.cycles) P.cycles[blockIdx.x / 2] = clock64() - t0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (P.D)
        for (int c0 = 0; c0 < P.N; c0 += 8) {
            uint32_t r[8];
            tmem_ld8(tb + ((unsigned)(warp * 32) << 16) + c0, r);
            for (int q = 0; q < 8; q++) P.D[(size_t)(128 * rank + tid) * P.N + c0 + q] = (int)r[q];
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int pair_main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount & ~1;
    CK(cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int N : {32, 192, 256}) {
        const int ksteps = 4, K = 128;
        std::vector<uint8_t> A((size_t)256 * K);
        std::vector<int8_t> B((size_t)N * K);
        srand(77 + N);
        for (auto &v : A) v = (uint8_t)(rand() & 0xFF);
        for (auto &v : B) v = (int8_t)((rand() & 0xFF) - 128);
        std::vector<int> ref((size_t)256 * N), got((size_t)256 * N, -1);
        for (int m = 0; m < 256; m++)
            for (int n = 0; n < N; n++) {
                int s = 0;
                for (int k = 0; k < K; k++) s += (int)A[(size_t)m * K + k] * (int)B[(size_t)n * K + k];
                ref[(size_t)m * N + n] = s;
            }
        uint8_t *dA; int8_t *dB; int *dD;
        CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, got.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
        Pair2Args P{N, ksteps, 1, dA, dB, dD, nullptr, 0, 0};
        probe_pair_kernel<<<2, 128, 65536>>>(P);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair layout N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 2; }
        CK(cudaMemcpy(got.data(), dD, got.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < got.size(); i++) bad += got[i] != ref[i];
        printf("cta_group::2, A in TMEM (128 rows per CTA), B halves per CTA, M=256 N=%3d K=%d : %zu / %zu mismatches%s\n", N, K, bad, got.size(),
               bad ? "" : "   <== OK");
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    long long *dcyc;
    CK(cudaMalloc(&dcyc, sizeof(long long) * G));
    printf("\ncta_group::2 throughput (%d SMs = %d pairs, M = 256, 4 k-steps per iteration, 2000 iterations):\n", G, G / 2);
    for (int variant = 0; variant < 3; variant++)
    for (int N : {32, 64, 128, 192, 256}) {
        if (variant && N != 192) continue;
        Pair2Args P{N, 4, 2000, nullptr, nullptr, nullptr, dcyc, variant == 1 ? 1 : 0, variant == 2 ? 1 : 0};
        if (variant == 1) printf("  (a multicast commit after every 4 MMAs)\n");
        if (variant == 2) printf("  (three warps keep issuing tcgen05.st + wait::st into other columns)\n");
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        probe_pair_kernel<<<G, 128, 65536>>>(P);
        CK(cudaEventRecord(e0));
        probe_pair_kernel<<<G, 128, 65536>>>(P);
        CK(cudaEventRecord(e1));
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("pair rate kernel: CUDA error %s\n", cudaGetErrorString(e)); return 2; }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> cyc(G / 2);
        CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * (G / 2), cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (long long v : cyc) mx = v > mx ? v : mx;
        const double mac = 8000.0 * 256.0 * N * 32;
        printf("  pair TS N=%3d : %8lld clk -> %6.1f clk per pair MMA, %7.1f MAC/clk/SM, chip %6.1f TMAC/s (event %.3f ms)\n", N, mx, (double)mx / 8000.0,
               mac / (double)mx / 2.0, mac * (G / 2) / (ms * 1e-3) / 1e12, ms);
    }
    return 0;
}
