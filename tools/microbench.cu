// Pipe-rate microbenchmarks that decide the product kernel design (DESIGN.md "Why int8 tensor cores").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Prints per-SM per-clock rates (from clock64 inside the kernel) and whole-chip rates (CUDA events).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void k_dfma(double *out, long long *cyc, double a, double b) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dadd_pred(double *out, long long *cyc, const uint32_t *wsrc, double b) {
    double x[8];
    for (int i = 0; i < 8; i++) x[i] = i;
    uint32_t w = wsrc[threadIdx.x & 31];
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) if (w & (1u << ((it + i) & 31))) x[i] += b;
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_lop3(uint32_t *out, long long *cyc, uint32_t a, uint32_t b) {
    uint32_t x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 2654435761u + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6A;" : "+r"(x[i]) : "r"(a), "r"(b));
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int WITH_LOP>
__global__ void k_imma(int *out, long long *cyc, uint32_t seed) {
    int c[8][4];
    uint32_t a[4], b[2], w[4];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
    for (int j = 0; j < 4; j++) { a[j] = seed * (threadIdx.x + j + 1); w[j] = a[j] ^ 0x5a5a5a5a; }
    b[0] = seed + threadIdx.x; b[1] = seed ^ threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (WITH_LOP) {
                // 4 LOP3 per IMMA: the mask work of the real kernel (w & (0x03030303 << 2k))
#pragma unroll
                for (int j = 0; j < 4; j++) asm volatile("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(a[j]) : "r"(w[j]), "r"(0x03030303u << (2 * (i & 3))), "r"(0xffffffffu - it));
            }
            imma(c[i], a, b);
        }
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_dmma(double *out, long long *cyc, double seed) {
    double c[8][2];
    for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
    double a = seed * threadIdx.x, b = seed + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// DFMA and DMMA issued by different warps of the same SM: do the two FP64 paths overlap?
__global__ void k_dfma_dmma(double *out, long long *cyc, double seed) {
    const int warp = threadIdx.x >> 5;
    double s = 0;
    long long t0 = clock64();
    if (warp & 1) {
        double c[8][2];
        for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
        double a = seed * threadIdx.x, b = seed + threadIdx.x;
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    } else {
        double x[8];
        for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fma(x[i], seed, 0.5);
        }
        for (int i = 0; i < 8; i++) s += x[i];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ldmatrix_probe(uint32_t *out) {
    __shared__ __align__(128) uint8_t sm[16 * 16 * 2];
    for (int i = threadIdx.x; i < 512; i += 32) sm[i] = (uint8_t)i;
    __syncwarp();
    // rows of 16 bytes; lane l supplies the address of row l (lanes 0..15 used)
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(sm + (threadIdx.x & 15) * 16);
    uint32_t r0 = 0, r1 = 0;
    asm volatile("ldmatrix.sync.aligned.m16n16.x1.trans.shared.b8 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
    out[threadIdx.x * 2] = r0;
    out[threadIdx.x * 2 + 1] = r1;
}

template <typename F>
int run(const char *name, F launch, double ops_per_thread_iter, int threads, int blocks_per_sm, int sms, const char *unit) {
    long long *cyc;
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * blocks_per_sm));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(cyc);  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    launch(cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(sms * blocks_per_sm);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (long long v : h) mean += v;
    mean /= h.size();
    const double ops_block = ops_per_thread_iter * ITERS * threads;
    printf("%-28s %8.1f %s/clk/SM   (%.2f T%s/s chip, %.3f ms, mean %.0f cycles/block, %d thr x %d blk/SM)\n", name,
           ops_block * blocks_per_sm / mean, unit, ops_block * blocks_per_sm * sms / (ms * 1e-3) / 1e12, unit, ms, mean, threads,
           blocks_per_sm);
    cudaFree(cyc);
    return 0;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
    void *buf;
    CK(cudaMalloc(&buf, 1 << 26));
    uint32_t hw[32];
    for (int i = 0; i < 32; i++) hw[i] = 0x9E3779B9u * (i + 1);
    uint32_t *dw;
    CK(cudaMalloc(&dw, sizeof(hw)));
    CK(cudaMemcpy(dw, hw, sizeof(hw), cudaMemcpyHostToDevice));
    for (int bps : {1, 2}) {
        const int T = 512;
        run("DFMA (fp64 fma)", [&](long long *c) { k_dfma<<<sms * bps, T>>>((double *)buf, c, 1.0000001, 1e-9); }, 8, T, bps, sms, "FMA");
        run("predicated DADD", [&](long long *c) { k_dadd_pred<<<sms * bps, T>>>((double *)buf, c, dw, 1e-9); }, 8, T, bps, sms, "ADD");
        run("LOP3", [&](long long *c) { k_lop3<<<sms * bps, T>>>((uint32_t *)buf, c, 0x12345678u, 0x9abcdef0u); }, 8, T, bps, sms, "op");
        run("IMMA m16n8k32 u8*s8 (MACs)", [&](long long *c) { k_imma<0><<<sms * bps, T>>>((int *)buf, c, 12345u); }, 8 * 4096.0 / 32, T, bps, sms, "MAC");
        run("IMMA + 4 LOP3 each (MACs)", [&](long long *c) { k_imma<1><<<sms * bps, T>>>((int *)buf, c, 12345u); }, 8 * 4096.0 / 32, T, bps, sms, "MAC");
        run("DMMA m8n8k4 f64 (FMAs)", [&](long long *c) { k_dmma<<<sms * bps, T>>>((double *)buf, c, 1.0000001); }, 8 * 256.0 / 32, T, bps, sms, "FMA");
        run("DFMA || DMMA (warps split)", [&](long long *c) { k_dfma_dmma<<<sms * bps, T>>>((double *)buf, c, 1.0000001); }, 8 * (1 + 256.0 / 32) / 2, T, bps, sms, "FMA");
    }
    // IMMA rate as a function of warps per scheduler (the fused kernel has 2 compute warps per SMSP)
    for (int T : {128, 256, 384, 512, 768, 1024}) {
        run("IMMA (MACs)", [&](long long *c) { k_imma<0><<<sms, T>>>((int *)buf, c, 12345u); }, 8 * 4096.0 / 32, T, 1, sms, "MAC");
        run("IMMA + 4 LOP3 (MACs)", [&](long long *c) { k_imma<1><<<sms, T>>>((int *)buf, c, 12345u); }, 8 * 4096.0 / 32, T, 1, sms, "MAC");
    }
    uint32_t *probe;
    CK(cudaMalloc(&probe, 64 * 4));
    k_ldmatrix_probe<<<1, 32>>>(probe);
    CK(cudaDeviceSynchronize());
    uint32_t hp[64];
    CK(cudaMemcpy(hp, probe, sizeof(hp), cudaMemcpyDeviceToHost));
    printf("ldmatrix.m16n16.x1.trans.b8 fragment (smem byte value = 16*row + col):\n");
    for (int l = 0; l < 32; l++) printf("lane %2d: r0 = %08x  r1 = %08x\n", l, hp[2 * l], hp[2 * l + 1]);
    return 0;
}
