// comm.cu -- the one collective of the path: the gather of the per-GPU PSF (or encircled-energy) stacks to one rank
// (the `paos_gather_psf` row of 8b).  The reference has no counterpart: its joblib fan-out
// (paos/core/pipeline.py:140-150) returns results through process pickling.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy torch has already loaded when the caller is a
// torch.distributed process, else the system one), so that the library keeps linking against libcudart only and loads on a
// box without NCCL.  The gather is grouped point-to-point (ncclSend / ncclRecv inside one ncclGroup), which takes ragged
// per-rank counts and lands every block at its final offset of the destination stack; the root's own block is a
// device-to-device copy on the same stream.  Nothing of ours computes here, so there is nothing to fuse: the overlap that
// matters is with the sweep itself, and the caller gets it by gathering finished chunks on a side stream while later
// wavelengths still propagate (paos_b200/sweep.py: ChunkGather).
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/paos_b200.h"

namespace {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId_ {
    char internal[128];
};
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;  // ncclUint8 / ncclChar+1 in nccl.h's ncclDataType_t

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId_*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId_, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

Nccl g_nccl;
std::once_flag g_nccl_once;
thread_local char g_comm_error[512];

void load_nccl() {
    // prefer a copy that is already mapped into the process (torch's bundled NCCL), then the usual search path
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_nccl.handle = h;
#define PAOS_SYM(field, name)                                   \
    *(void**)(&g_nccl.field) = dlsym(h, name);                  \
    if (!g_nccl.field) return;
    PAOS_SYM(GetUniqueId, "ncclGetUniqueId")
    PAOS_SYM(CommInitRank, "ncclCommInitRank")
    PAOS_SYM(CommDestroy, "ncclCommDestroy")
    PAOS_SYM(Send, "ncclSend")
    PAOS_SYM(Recv, "ncclRecv")
    PAOS_SYM(GroupStart, "ncclGroupStart")
    PAOS_SYM(GroupEnd, "ncclGroupEnd")
    PAOS_SYM(GetErrorString, "ncclGetErrorString")
#undef PAOS_SYM
    g_nccl.ok = true;
}

int comm_fail(int code, const char* what, const char* detail) {
    snprintf(g_comm_error, sizeof g_comm_error, "%s: %s", what, detail ? detail : "");
    return code;
}

}  // namespace

struct paos_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

extern "C" {

const char* paos_comm_last_error(void) { return g_comm_error; }

int paos_comm_unique_id(void* id128) {
    if (!id128) return comm_fail(PAOS_ERR_ARG, "paos_comm_unique_id", "null argument");
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) return comm_fail(PAOS_ERR_UNSUPPORTED, "NCCL", "libnccl.so.2 could not be loaded");
    ncclUniqueId_ id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return comm_fail(PAOS_ERR_CUDA, "ncclGetUniqueId", g_nccl.GetErrorString(r));
    std::memcpy(id128, &id, sizeof id);
    return PAOS_OK;
}

int paos_comm_create(paos_comm** out, const void* id128, int rank, int world, int device) {
    if (!out || !id128) return comm_fail(PAOS_ERR_ARG, "paos_comm_create", "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return comm_fail(PAOS_ERR_ARG, "paos_comm_create", "bad rank / world size");
    std::call_once(g_nccl_once, load_nccl);
    if (!g_nccl.ok) return comm_fail(PAOS_ERR_UNSUPPORTED, "NCCL", "libnccl.so.2 could not be loaded");
    if (cudaSetDevice(device) != cudaSuccess) return comm_fail(PAOS_ERR_CUDA, "cudaSetDevice", cudaGetErrorString(cudaGetLastError()));
    ncclUniqueId_ id;
    std::memcpy(&id, id128, sizeof id);
    paos_comm* c = new paos_comm();
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != 0) {
        delete c;
        return comm_fail(PAOS_ERR_CUDA, "ncclCommInitRank", g_nccl.GetErrorString(r));
    }
    *out = c;
    return PAOS_OK;
}

int paos_comm_destroy(paos_comm* c) {
    if (!c) return PAOS_OK;
    if (c->comm && g_nccl.ok) g_nccl.CommDestroy(c->comm);
    delete c;
    return PAOS_OK;
}

int paos_gather_psf(paos_comm* c, const void* local_dev, const size_t* bytes_per_rank, const size_t* dst_offsets, void* dst_dev, int root,
                    void* stream) {
    if (!c || !bytes_per_rank) return comm_fail(PAOS_ERR_ARG, "paos_gather_psf", "null argument");
    if (root < 0 || root >= c->world) return comm_fail(PAOS_ERR_ARG, "paos_gather_psf", "root out of range");
    if (c->rank == root && !dst_dev) return comm_fail(PAOS_ERR_ARG, "paos_gather_psf", "the root needs a destination");
    const size_t mine = bytes_per_rank[c->rank];
    if (mine && !local_dev) return comm_fail(PAOS_ERR_ARG, "paos_gather_psf", "null local block");
    if (cudaSetDevice(c->device) != cudaSuccess) return comm_fail(PAOS_ERR_CUDA, "cudaSetDevice", cudaGetErrorString(cudaGetLastError()));
    cudaStream_t st = (cudaStream_t)stream;
    ncclResult_t r = g_nccl.GroupStart();
    if (r != 0) return comm_fail(PAOS_ERR_CUDA, "ncclGroupStart", g_nccl.GetErrorString(r));
    if (c->rank == root) {
        size_t off = 0;
        for (int q = 0; q < c->world; ++q) {
            char* at = static_cast<char*>(dst_dev) + (dst_offsets ? dst_offsets[q] : off);
            if (q == root) {
                if (mine && at != local_dev && cudaMemcpyAsync(at, local_dev, mine, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
                    g_nccl.GroupEnd();
                    return comm_fail(PAOS_ERR_CUDA, "cudaMemcpyAsync", cudaGetErrorString(cudaGetLastError()));
                }
            } else if (bytes_per_rank[q]) {
                r = g_nccl.Recv(at, bytes_per_rank[q], kNcclUint8, q, c->comm, st);
                if (r != 0) {
                    g_nccl.GroupEnd();
                    return comm_fail(PAOS_ERR_CUDA, "ncclRecv", g_nccl.GetErrorString(r));
                }
            }
            off += bytes_per_rank[q];
        }
    } else if (mine) {
        r = g_nccl.Send(local_dev, mine, kNcclUint8, root, c->comm, st);
        if (r != 0) {
            g_nccl.GroupEnd();
            return comm_fail(PAOS_ERR_CUDA, "ncclSend", g_nccl.GetErrorString(r));
        }
    }
    r = g_nccl.GroupEnd();
    if (r != 0) return comm_fail(PAOS_ERR_CUDA, "ncclGroupEnd", g_nccl.GetErrorString(r));
    return PAOS_OK;
}

}  // extern "C"
