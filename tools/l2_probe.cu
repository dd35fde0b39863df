// tools/l2_probe.cu -- how fast can all SMs pull an L2-resident operand matrix through TMA, and does cluster multicast help?
// The batched GRM kernel streams the digit matrix (96 MB at N = 430K, K = 32) once per 256-row block: 37 GB per phase out of L2.
//   mode 0: every CTA loads the same tile sequence (what the GEMM CTAs of one wave do)
//   mode 1: every CTA loads a different part of the matrix (no request merging possible)
//   mode 2: clusters of 2, each CTA loads half of every tile and multicasts it to both (SM ingress = 2x the L2 reads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/l2_probe tools/l2_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    const unsigned a = smem_u32(b);
    unsigned ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *tmap, int x, int y, unsigned long long *bar, unsigned short mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar)), "h"(mask) : "memory");
}

constexpr int kRows = 192, kSlots = 6, kTile = kRows * 128;

// mode 0 / 1: plain; the consumer is a single thread that waits for each tile and immediately frees the slot
__global__ void __launch_bounds__(64, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, int n_tiles, int tiles_total, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long full[kSlots];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; i++) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int off = (mode == 1) ? (int)((long long)blockIdx.x * tiles_total / gridDim.x) : 0;
        for (int t = 0; t < n_tiles + kSlots; t++) {
            if (t >= kSlots) mbar_wait(&full[(t - kSlots) % kSlots], (unsigned)(((t - kSlots) / kSlots) & 1));
            if (t < n_tiles) {
                const int s = t % kSlots;
                mbar_expect_tx(&full[s], kTile);
                tma_load_2d(smem + (size_t)s * kTile, &tm, ((t + off) % tiles_total) * 128, 0, &full[s]);
            }
        }
    }
}

// mode 2: cluster of 2, each CTA loads rows [96 r, 96 r + 96) of the tile and multicasts them to both CTAs
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64, 1) stream_mc_kernel(const __grid_constant__ CUtensorMap tm_half, int n_tiles, int tiles_total) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ unsigned long long full[kSlots];
    unsigned rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSlots; i++) mbar_init(&full[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (threadIdx.x == 0) {
        // a slot may only be refilled when BOTH CTAs have consumed it; this probe has no consumer work, so the two producers simply
        // run in lock step through a cluster barrier every kSlots tiles
        for (int t0 = 0; t0 < n_tiles; t0 += kSlots) {
            for (int t = t0; t < t0 + kSlots && t < n_tiles; t++) {
                const int s = t % kSlots;
                mbar_expect_tx(&full[s], kTile);
                tma_load_2d_mc(smem + (size_t)s * kTile + rank * (kTile / 2), &tm_half, (t % tiles_total) * 128, (int)rank * (kRows / 2), &full[s], (unsigned short)3);
            }
            for (int t = t0; t < t0 + kSlots && t < n_tiles; t++) mbar_wait(&full[t % kSlots], (unsigned)((t / kSlots) & 1));
            asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
        }
    } else {
        for (int t0 = 0; t0 < n_tiles; t0 += kSlots) asm volatile("barrier.cluster.arrive.release;\nbarrier.cluster.wait.acquire;" ::: "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

static CUtensorMap make_map(void *base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {cols, rows}, gstride[1] = {cols};
    const cuuint32_t box[2] = {128, box_rows}, estr[2] = {1, 1};
    CUresult r = ((EncodeFn)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return tm;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    for (uint64_t cols : {(uint64_t)430080, (uint64_t)100352}) {
        uint8_t *d;
        CK(cudaMalloc(&d, cols * kRows));
        CK(cudaMemset(d, 1, cols * kRows));
        const int tiles_total = (int)(cols / 128);
        const int n_tiles = 20000;
        CUtensorMap tm = make_map(d, cols, kRows, kRows), tmh = make_map(d, cols, kRows, kRows / 2);
        CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kTile));
        CK(cudaFuncSetAttribute(stream_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * kTile));
        printf("digit matrix %d rows x %llu B = %.1f MB; %d tiles of %d B per CTA, %d CTAs\n", kRows, (unsigned long long)cols, cols * kRows / 1e6,
               n_tiles, kTile, G);
        for (int mode = 0; mode < 3; mode++) {
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaEventRecord(e0));
                if (mode < 2) stream_kernel<<<G, 64, kSlots * kTile>>>(tm, n_tiles, tiles_total, mode);
                else stream_mc_kernel<<<G, 64, kSlots * kTile>>>(tmh, n_tiles, tiles_total);
                CK(cudaEventRecord(e1));
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
            }
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double ingress = (double)G * n_tiles * kTile;
            printf("  mode %d (%s): %.3f ms, SM ingress %.2f TB/s (%.1f B/clk/SM at 1965 MHz)\n", mode,
                   mode == 0 ? "same tiles everywhere" : mode == 1 ? "different tiles per CTA" : "cluster 2, half tile each, multicast", ms,
                   ingress / (ms * 1e-3) / 1e12, ingress / (ms * 1e-3) / G / 1.965e9);
        }
        cudaFree(d);
    }
    return 0;
}
