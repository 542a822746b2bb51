// NCCL collectives behind the C ABI: the only exchange steps of this path are the gather of the
// packed per-fit results and the sum of moment buffers AFTER a sharded batch (the reference loops over bootstrap copies
// in one Python process: src/lsqfit/__init__.py:1612-1624; here every rank fits its shard and the results meet once).
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch has already loaded is reused), so the library
// itself has no link-time dependency and still loads on machines without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>
#include <string>
#include "../../include/b200lm.h"
#include "handle.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;           // 0 = ncclSuccess
enum { kNcclFloat64 = 8, kNcclSum = 0 };   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    if (a.lib) return a;
    a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.lib, "ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce;
    return a;
}

int nccl_fail(ncclResult_t r, const char* what) {
    const char* msg = api().GetErrorString ? api().GetErrorString(r) : "?";
    return b200lm::set_error(nullptr, B200LM_ECUDA, std::string(what) + ": NCCL error: " + msg);
}

}  // namespace

struct b200lm_comm_s {
    ncclComm_t comm = nullptr;
    int device = 0, rank = 0, world = 1;
};

using namespace b200lm;

extern "C" {

int b200lm_comm_unique_id(char* out128) {
    if (!out128) return set_error(nullptr, B200LM_EINVAL, "NULL argument");
    if (!api().ok) return set_error(nullptr, B200LM_ECUDA, "NCCL (libnccl.so.2) is not available");
    ncclUniqueId id;
    ncclResult_t r = api().GetUniqueId(&id);
    if (r) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(out128, id.internal, 128);
    return B200LM_OK;
}

int b200lm_comm_init(int device, int rank, int world, const char* id128, b200lm_comm* out) {
    if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return set_error(nullptr, B200LM_EINVAL, "bad comm_init argument");
    *out = nullptr;
    if (!api().ok) return set_error(nullptr, B200LM_ECUDA, "NCCL (libnccl.so.2) is not available");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    b200lm_comm_s* c = new b200lm_comm_s();
    c->device = device; c->rank = rank; c->world = world;
    ncclResult_t r = api().CommInitRank(&c->comm, world, id, rank);
    if (r) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    *out = c;
    return B200LM_OK;
}

void b200lm_comm_destroy(b200lm_comm c) {
    if (!c) return;
    if (c->comm && api().ok) api().CommDestroy(c->comm);
    delete c;
}

int b200lm_comm_rank(b200lm_comm c) { return c ? c->rank : B200LM_EINVAL; }
int b200lm_comm_world(b200lm_comm c) { return c ? c->world : B200LM_EINVAL; }

int b200lm_gather(b200lm_comm c, const double* d_send, double* d_recv, long long count_per_rank, void* stream) {
    if (!c || !d_send || !d_recv || count_per_rank < 0) return set_error(nullptr, B200LM_EINVAL, "bad gather argument");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    ncclResult_t r = api().AllGather(d_send, d_recv, (size_t)count_per_rank, kNcclFloat64, c->comm, (cudaStream_t)stream);
    if (r) return nccl_fail(r, "ncclAllGather");
    return B200LM_OK;
}

int b200lm_allreduce_sum(b200lm_comm c, double* d_buf, long long count, void* stream) {
    if (!c || !d_buf || count < 0) return set_error(nullptr, B200LM_EINVAL, "bad allreduce argument");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    ncclResult_t r = api().AllReduce(d_buf, d_buf, (size_t)count, kNcclFloat64, kNcclSum, c->comm, (cudaStream_t)stream);
    if (r) return nccl_fail(r, "ncclAllReduce");
    return B200LM_OK;
}

}  // extern "C"
