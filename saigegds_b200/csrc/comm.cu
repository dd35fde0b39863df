// Multi-GPU plumbing: one process per GPU, variants sharded, one sum all-reduce of the N x k product
// block per GRM product over NVLink/NVSwitch.  The reference is single-process (TBB threads,
// saige_fitnull.cpp:44-87); the per-thread buffers it reduces at :523-535 become per-GPU partial
// vectors reduced here.
//
// NCCL is bound at run time with dlopen/dlsym so that (a) the single-GPU library has no NCCL
// dependency and (b) when the host process is Python with torch already loaded, the very same
// libnccl.so.2 that torch uses is picked up (no second copy of the library in the process).
#include <dlfcn.h>

#include <cstring>

#include "ctx.h"

namespace sgb {

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };  // ncclDataType_t: ncclDouble
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
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

struct Comm {
    ncclComm_t comm = nullptr;
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

void comm_destroy(Context &c) {
    if (c.comm) {
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
    check(api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, c.comm->comm, c.stream), "ncclAllReduce");
}

}  // namespace sgb
