// Multi-GPU plumbing: one process per GPU, variants sharded, one sum all-reduce of the N x k product
// block per GRM product over NVLink/NVSwitch.  The reference is single-process (TBB threads,
// saige_fitnull.cpp:44-87); the per-thread buffers it reduces at :523-535 become per-GPU partial
// vectors reduced here.
//
// NCCL is bound at run time with dlopen/dlsym so that (a) the single-GPU library has no NCCL
// dependency and (b) when the host process is Python with torch already loaded, the very same
// libnccl.so.2 that torch uses is picked up (no second copy of the library in the process).
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ctx.h"

namespace sgb {

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };
enum { ncclUint8 = 1 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;   // optional: peer set-up only
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &api() {
    static NcclApi a;
    if (a.handle) return a;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) throw Error(SGB_ERR_COMM, std::string("cannot load libnccl.so.2: ") + dlerror());
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.handle, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.handle, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.handle, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.handle, "ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))dlsym(a.handle, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.handle, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce)
        throw Error(SGB_ERR_COMM, "libnccl.so.2 lacks a required symbol");
    return a;
}

void check(int rc, const char *what) {
    if (rc != ncclSuccess) {
        const char *s = api().GetErrorString ? api().GetErrorString(rc) : "?";
        throw Error(SGB_ERR_COMM, std::string(what) + ": " + s);
    }
}

}  // namespace

// ---- all-reduce of one N-vector over peer memory --------------------------------------------------------------------------------------
// The product of a single right-hand side ends in a sum over the ranks of 3.44 MB (N = 430K): latency, not bandwidth.  On one NVSwitch
// node every rank maps the others' exchange buffers (cudaIpc handles travel over the NCCL communicator at first use) and ONE kernel
// per rank does the whole collective: copy the input into its own send buffer S -> signal -> wait for all ranks -> sum its slice of
// the vector from all S in rank order (remote loads) and push the result into every rank's receive buffer D (remote stores) ->
// signal -> wait -> copy D to the destination.  Same bits on every rank (fixed order).  Anything else -- several columns, ranks on
// different hosts, a failed mapping on any rank -- uses ncclAllReduce.
constexpr int kPeerMax = 8;
constexpr int kPeerThreads = 1024;           // one block per SM (cooperative launch: the grid-wide rendezvous need all blocks resident)
struct PeerArgs {
    double *S[kPeerMax];                       // send buffers of all ranks, as mapped here
    double *D[kPeerMax];                       // receive buffers
    unsigned long long *flags[kPeerMax];       // [2][kPeerMax] arrival epochs per rank: phase A (S filled), phase B (D filled)
    int rank, world;
    int one_shot;                              // two ranks: no receive buffers, every rank sums both send buffers itself
    long long n;
    unsigned long long epoch;
    double *buf;
    CombineSrc comb;                           // comb.rout != nullptr: the input is formed here from the product's parts (no combine launch)
    unsigned int *grid_bar;                    // local: two counters for the two grid-wide rendezvous
    int *err;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long peer_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_sys_f64x2(const double *p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
// all blocks of this grid have arrived (cooperative launch: every block is resident); `target` grows by gridDim.x per use
__device__ __forceinline__ bool peer_grid_sync(unsigned int *counter, unsigned int target, volatile int *err, unsigned long long deadline) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while ((int)(*(volatile unsigned int *)counter - target) < 0) {
            if (*err || peer_now() > deadline) { *err = 1; ok = false; break; }
        }
        __threadfence();
    }
    return __syncthreads_and(ok ? 1 : 0) != 0;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs A, unsigned int bar_base) {
    const int R = A.world, me = A.rank;
    const long long n = A.n;
    volatile int *err = A.err;
    const unsigned long long deadline = peer_now() + A.timeout_ns;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    // one-shot (two ranks): the send buffer alternates between the two halves of the block with the epoch -- a rank raises its flag
    // of epoch k only after its kernel of epoch k-1 ended, i.e. after it read this rank's buffer of epoch k-1, and this rank has seen
    // that flag before it starts epoch k+1, which is the next writer of that buffer
    double *const mine = (A.one_shot && (A.epoch & 1)) ? A.D[me] : A.S[me];
    // 1. input -> own send buffer; for the fused product the input is formed on the way: out_n = R_n + H - corr_n (what
    //    combine_fused_kernel computes, same order of additions)
    if (A.comb.rout != nullptr) {
        __shared__ double sh_h;
        if (threadIdx.x == 0) {
            double h = 0;
            for (int i = 0; i < A.comb.n_hpart; i++) h += A.comb.h_part[i];
            sh_h = h;
        }
        __syncthreads();
        const double h = sh_h;
        const int T = A.comb.n_ctiles;
        for (long long i = tid; i < n; i += nthr) {
            const double r = A.comb.rout[i];
            double corr = 0;
            int t = 0;
            for (; t + 8 <= T; t += 8) {       // eight loads in flight, added in tile order
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = A.comb.cpart[(size_t)(t + u) * n + i];
#pragma unroll
                for (int u = 0; u < 8; u++) corr += v[u];
            }
            for (; t < T; t++) corr += A.comb.cpart[(size_t)t * n + i];
            mine[i] = r + h - corr;
        }
    } else {
        for (long long i = tid; i < n; i += nthr) mine[i] = A.buf[i];
    }
    if (!peer_grid_sync(A.grid_bar, bar_base + gridDim.x, err, deadline)) return;
    // 2. tell every rank (release at system scope: the stores above, ordered before it by the rendezvous, become visible with it)
    if (blockIdx.x == 0 && threadIdx.x < R) st_release_sys(A.flags[threadIdx.x] + me, A.epoch);
    // 3. wait for every rank's send buffer
    {
        bool ok = true;
        if (threadIdx.x < R) {
            const unsigned long long *f = A.flags[me] + threadIdx.x;
            while (ld_acquire_sys(f) < A.epoch) {
                if (*err || peer_now() > deadline) { *err = 1; ok = false; break; }
            }
        }
        if (!__syncthreads_and(ok ? 1 : 0)) return;
    }
    if (A.one_shot) {
        // 4'. every rank sums all send buffers itself, in rank order (same bits everywhere), straight into the destination
        const long long pairs = n / 2;
        for (long long q = tid; q < pairs; q += nthr) {
            const long long i = 2 * q;
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int r = 0; r < kPeerMax; r++)
                if (r < R) {
                    const double *src = (A.epoch & 1) ? A.D[r] : A.S[r];
                    const double2 v = ld_sys_f64x2(src + i);
                    acc.x += v.x; acc.y += v.y;
                }
            *reinterpret_cast<double2 *>(A.buf + i) = acc;     // cudaMalloc'd N-vectors: 16-byte aligned at even i
        }
        if ((n & 1) && tid == 0) {
            double acc = 0;
            for (int r = 0; r < R; r++) acc += *(volatile double *)(((A.epoch & 1) ? A.D[r] : A.S[r]) + n - 1);
            A.buf[n - 1] = acc;
        }
        return;
    }
    // 4. my slice (pairs of doubles): sum over the ranks in rank order, result to every rank's receive buffer
    const long long pairs = (n + 1) / 2, per = (pairs + R - 1) / R, p0 = (long long)me * per, p1 = min(pairs, p0 + per);
    for (long long q = p0 + tid; q < p1; q += nthr) {
        const long long i = 2 * q;
        double2 acc = make_double2(0.0, 0.0);
        if (i + 1 < n) {
            double2 v[kPeerMax];
#pragma unroll
            for (int r = 0; r < kPeerMax; r++)
                if (r < R) v[r] = ld_sys_f64x2(A.S[r] + i);
#pragma unroll
            for (int r = 0; r < kPeerMax; r++)
                if (r < R) { acc.x += v[r].x; acc.y += v[r].y; }
#pragma unroll
            for (int r = 0; r < kPeerMax; r++)
                if (r < R) *reinterpret_cast<double2 *>(A.D[r] + i) = acc;
        } else {
            for (int r = 0; r < R; r++) acc.x += *(volatile double *)(A.S[r] + i);
            for (int r = 0; r < R; r++) A.D[r][i] = acc.x;
        }
    }
    __threadfence_system();
    if (!peer_grid_sync(A.grid_bar + 1, bar_base + gridDim.x, err, deadline)) return;
    // 5. / 6. second round of flags, then the full vector from the own receive buffer
    if (blockIdx.x == 0 && threadIdx.x < R) st_release_sys(A.flags[threadIdx.x] + kPeerMax + me, A.epoch);
    {
        bool ok = true;
        if (threadIdx.x < R) {
            const unsigned long long *f = A.flags[me] + kPeerMax + threadIdx.x;
            while (ld_acquire_sys(f) < A.epoch) {
                if (*err || peer_now() > deadline) { *err = 1; ok = false; break; }
            }
        }
        if (!__syncthreads_and(ok ? 1 : 0)) return;
    }
    for (long long i = tid; i < n; i += nthr) A.buf[i] = A.D[me][i];
}

struct PeerState {
    bool tried = false, ready = false;
    size_t cap = 0;                            // doubles per buffer
    DevBuf<unsigned char> block;               // own [flags 2 x 8 u64 | pad to 256 | S | D]
    void *mapped[kPeerMax] = {nullptr};        // peers' blocks (own entry: block.get())
    DevBuf<unsigned int> grid_bar;
    DevBuf<int> err;
    PinBuf<int> herr;
    unsigned long long epoch = 0;
    unsigned int bar_base = 0;
    int blocks = 0;
    bool one_shot = false;
};

struct Comm {
    ncclComm_t comm = nullptr;
    PeerState peer;
};

void comm_unique_id(unsigned char id[128]) {
    ncclUniqueId u;
    check(api().GetUniqueId(&u), "ncclGetUniqueId");
    memcpy(id, u.internal, 128);
}

void comm_init(Context &c, const unsigned char id[128], int rank, int world) {
    if (world < 1 || rank < 0 || rank >= world) throw Error(SGB_ERR_INVALID, "invalid rank / world_size");
    comm_destroy(c);
    c.rank = rank;
    c.world = world;
    if (world == 1) return;
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    SGB_CUDA(cudaSetDevice(c.dev));
    c.comm = new Comm();
    check(api().CommInitRank(&c.comm->comm, world, u, rank), "ncclCommInitRank");
}

namespace {

struct PeerRecord {
    cudaIpcMemHandle_t handle;
    unsigned long long host_hash;
    int dev, ok;
};

// Collective (every rank calls it at its first single-vector all-reduce): allocate the exchange block, swap handles over NCCL,
// map the peers.  All ranks agree on the outcome.
void peer_setup(Context &c, size_t count) {
    PeerState &P = c.comm->peer;
    P.tried = true;
    const int R = c.world;
    int ok = (R <= kPeerMax && api().AllGather != nullptr && getenv("SGB_NO_PEER_ALLREDUCE") == nullptr) ? 1 : 0;
    // Measured on one NVSwitch node (same box, alternating runs, profiles/r02_peer_vs_nccl_ab.txt): 2 ranks +1 %, 8 ranks +4.0 % over
    // combine + ncclAllReduce, 4 ranks -1 % (one-shot and two-shot alike): three and four ranks stay on NCCL unless SGB_PEER_FORCE=1
    if ((R == 3 || R == 4) && getenv("SGB_PEER_FORCE") == nullptr) ok = 0;
    const size_t cap = (count + 31) / 32 * 32, bytes = 256 + 2 * cap * sizeof(double);
    PeerRecord mine{};
    if (ok) {
        try {
            P.block.ensure(bytes);
            SGB_CUDA(cudaMemsetAsync(P.block.get(), 0, bytes, c.stream));
            P.grid_bar.ensure(2);
            SGB_CUDA(cudaMemsetAsync(P.grid_bar.get(), 0, 2 * sizeof(unsigned int), c.stream));
            P.err.ensure(1);
            SGB_CUDA(cudaMemsetAsync(P.err.get(), 0, sizeof(int), c.stream));
            P.herr.ensure(1);
            *P.herr.p = 0;
            SGB_CUDA(cudaIpcGetMemHandle(&mine.handle, P.block.get()));
        } catch (const Error &) {
            cudaGetLastError();
            ok = 0;
        }
    }
    char host[256] = {0};
    gethostname(host, sizeof(host) - 1);
    unsigned long long h = 1469598103934665603ull;
    for (const char *q = host; *q; q++) h = (h ^ (unsigned char)*q) * 1099511628211ull;
    mine.host_hash = h; mine.dev = c.dev; mine.ok = ok;
    // round 1: records
    DevBuf<unsigned char> send, recv;
    send.ensure(sizeof(PeerRecord)); recv.ensure(sizeof(PeerRecord) * R);
    std::vector<PeerRecord> all(R);
    auto gather = [&](const void *src, void *dst_host, size_t each) {
        SGB_CUDA(cudaMemcpyAsync(send.get(), src, each, cudaMemcpyHostToDevice, c.stream));
        check(api().AllGather(send.get(), recv.get(), each, ncclUint8, c.comm->comm, c.stream), "ncclAllGather");
        SGB_CUDA(cudaMemcpyAsync(dst_host, recv.get(), each * R, cudaMemcpyDeviceToHost, c.stream));
        SGB_CUDA(cudaStreamSynchronize(c.stream));
    };
    if (api().AllGather == nullptr) return;          // same on every rank (same library): nothing was exchanged
    gather(&mine, all.data(), sizeof(PeerRecord));
    for (int r = 0; r < R; r++) ok = ok && all[r].ok && all[r].host_hash == mine.host_hash;
    if (ok) {
        for (int r = 0; r < R && ok; r++) {
            if (r == c.rank) { P.mapped[r] = P.block.get(); continue; }
            if (cudaIpcOpenMemHandle(&P.mapped[r], all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                P.mapped[r] = nullptr;
                ok = 0;
            }
        }
    }
    // round 2: did every rank map every peer?
    std::vector<PeerRecord> st(R);
    mine.ok = ok;
    gather(&mine, st.data(), sizeof(PeerRecord));
    for (int r = 0; r < R; r++) ok = ok && st[r].ok;
    if (!ok) {
        for (int r = 0; r < R; r++)
            if (r != c.rank && P.mapped[r]) { cudaIpcCloseMemHandle(P.mapped[r]); P.mapped[r] = nullptr; }
        P.block.release();
        return;
    }
    {   // grid: one block per SM, fewer if the kernel's occupancy says so (all blocks must be resident)
        int sms = 0, per_sm = 0;
        SGB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c.dev));
        SGB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer_allreduce_kernel, kPeerThreads, 0));
        P.blocks = std::max(1, sms * std::min(per_sm, 1));
    }
    // one-shot (every rank sums all send buffers itself) up to this many ranks, reduce-scatter + all-gather above
    int one_shot_max = 2;
    if (const char *e = getenv("SGB_PEER_ONE_SHOT_MAX")) one_shot_max = atoi(e);
    P.one_shot = (R <= one_shot_max) && getenv("SGB_PEER_TWO_SHOT") == nullptr;
    P.cap = cap;
    P.ready = true;
}

void peer_allreduce(Context &c, double *buf, size_t count, const CombineSrc *comb) {
    PeerState &P = c.comm->peer;
    PeerArgs a{};
    if (comb) a.comb = *comb;
    for (int r = 0; r < c.world; r++) {
        unsigned char *base = (unsigned char *)P.mapped[r];
        a.flags[r] = (unsigned long long *)base;
        a.S[r] = (double *)(base + 256);
        a.D[r] = a.S[r] + P.cap;
    }
    a.rank = c.rank; a.world = c.world; a.one_shot = P.one_shot ? 1 : 0; a.n = (long long)count; a.epoch = ++P.epoch; a.buf = buf;
    a.grid_bar = P.grid_bar.get(); a.err = P.err.get(); a.timeout_ns = 2000000000ull;
    if (const char *e = getenv("SGB_WAIT_TIMEOUT_MS")) a.timeout_ns = (unsigned long long)std::max(1, atoi(e)) * 1000000ull;
    unsigned int bar_base = P.bar_base;
    void *kargs[] = {&a, &bar_base};
    // cooperative launch: the kernel's two grid-wide rendezvous need all 64 blocks resident at once
    SGB_CUDA(cudaLaunchCooperativeKernel((const void *)peer_allreduce_kernel, dim3(P.blocks), dim3(kPeerThreads), kargs, 0, c.stream));
    SGB_CUDA(cudaMemcpyAsync(P.herr.p, P.err.get(), sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    c.comm_err = P.herr.p;
    P.bar_base += (unsigned int)P.blocks;
    c.stats.n_kernel_launches++;
}

}  // namespace

void comm_destroy(Context &c) {
    c.comm_err = nullptr;
    if (c.comm) {
        PeerState &P = c.comm->peer;
        if (P.ready) {
            cudaStreamSynchronize(c.stream);
            for (int r = 0; r < c.world && r < kPeerMax; r++)
                if (r != c.rank && P.mapped[r]) cudaIpcCloseMemHandle(P.mapped[r]);
        }
        if (c.comm->comm) api().CommDestroy(c.comm->comm);
        delete c.comm;
        c.comm = nullptr;
    }
    c.rank = 0;
    c.world = 1;
}

void comm_allreduce_sum(Context &c, double *buf, size_t count) {
    if (c.world <= 1) return;
    if (!c.comm) throw Error(SGB_ERR_STATE, "communicator not initialised");
    // one N-vector (the product of a single right-hand side): the peer-memory kernel; everything else: NCCL
    PeerState &P = c.comm->peer;
    if (count == (size_t)c.N && c.N > 0) {
        if (!P.tried) peer_setup(c, count);
        if (P.ready && count <= P.cap) {
            peer_allreduce(c, buf, count, nullptr);
            return;
        }
    }
    check(api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, c.comm->comm, c.stream), "ncclAllReduce");
}

// The last step of the fused single-RHS product and the sum over the ranks in ONE kernel: out = sum over ranks of (R + H - corr).
// False: the peer path is not available (one rank, other hosts, mapping refused) -- the caller launches its combine kernel and the
// all-reduce follows in grm_mv_device.
bool comm_combine_allreduce(Context &c, const CombineSrc &src, double *out) {
    if (c.world <= 1 || !c.comm || c.N <= 0) return false;
    PeerState &P = c.comm->peer;
    if (!P.tried) peer_setup(c, (size_t)c.N);
    if (!P.ready || (size_t)c.N > P.cap) return false;
    peer_allreduce(c, out, (size_t)c.N, &src);
    return true;
}

}  // namespace sgb
