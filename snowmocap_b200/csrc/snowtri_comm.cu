// Final all-gather of the 3D joints across the GPUs of one box (north_star: "NCCL over NVLink appears
// only as a final all-gather of 3D joints").  The path itself shards by frame and needs no collective; this entry
// point lets a host that binds the C ABI collect every rank's dense output block without going through Python.
//
// NCCL is loaded with dlopen at first use ("libnccl.so.2": the copy PyTorch already mapped into the process, or the
// system one), so libsnowtri.so has no link-time dependency on it.  Either the caller passes its own ncclComm_t, or
// the handle owns one: rank 0 calls snowtri_comm_unique_id, the host ships the 128 bytes to the other ranks by any
// means it has (torch.distributed, MPI, a file), every rank calls snowtri_comm_init.
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include "snowtri_internal.h"

namespace {

struct NcclId {
    char internal[128];   // ncclUniqueId
};
typedef int (*GetUniqueIdFn)(NcclId*);
typedef int (*CommInitRankFn)(void**, int, NcclId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllGatherFn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*ErrorStringFn)(int);

struct Nccl {
    void* lib;
    GetUniqueIdFn get_unique_id;
    CommInitRankFn comm_init_rank;
    CommDestroyFn comm_destroy;
    AllGatherFn all_gather;
    ErrorStringFn error_string;
};

Nccl g_nccl;   // process-wide; filled once

const char* load_nccl() {
    if (g_nccl.lib) return nullptr;
    void* lib = nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return "libnccl.so.2 not found (dlopen)";
    Nccl t;
    t.lib = lib;
    t.get_unique_id = (GetUniqueIdFn)dlsym(lib, "ncclGetUniqueId");
    t.comm_init_rank = (CommInitRankFn)dlsym(lib, "ncclCommInitRank");
    t.comm_destroy = (CommDestroyFn)dlsym(lib, "ncclCommDestroy");
    t.all_gather = (AllGatherFn)dlsym(lib, "ncclAllGather");
    t.error_string = (ErrorStringFn)dlsym(lib, "ncclGetErrorString");
    if (!t.get_unique_id || !t.comm_init_rank || !t.comm_destroy || !t.all_gather || !t.error_string)
        return "libnccl.so.2 lacks an expected symbol";
    g_nccl = t;
    return nullptr;
}

}  // namespace

extern "C" int snowtri_comm_unique_id(void* id128) {
    if (!id128) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_comm_unique_id: NULL buffer");
    if (const char* e = load_nccl()) return fail(nullptr, SNOWTRI_E_UNSUPPORTED, "snowtri_comm_unique_id: %s", e);
    NcclId id;
    const int rc = g_nccl.get_unique_id(&id);
    if (rc) return fail(nullptr, SNOWTRI_E_CUDA, "ncclGetUniqueId: %s", g_nccl.error_string(rc));
    memcpy(id128, id.internal, sizeof(id.internal));
    return SNOWTRI_OK;
}

extern "C" int snowtri_comm_init(snowtri_t* h, const void* id128, int nranks, int rank) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_comm_init: NULL handle");
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, SNOWTRI_E_ARG, "snowtri_comm_init: bad argument");
    if (h->nccl_comm) return fail(h, SNOWTRI_E_ARG, "snowtri_comm_init: the handle already owns a communicator");
    if (const char* e = load_nccl()) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_comm_init: %s", e);
    CUDA_TRY(h, cudaSetDevice(h->device));
    NcclId id;
    memcpy(id.internal, id128, sizeof(id.internal));
    void* comm = nullptr;
    const int rc = g_nccl.comm_init_rank(&comm, nranks, id, rank);
    if (rc) return fail(h, SNOWTRI_E_CUDA, "ncclCommInitRank: %s", g_nccl.error_string(rc));
    h->nccl_comm = comm;
    h->nccl_nranks = nranks;
    h->nccl_rank = rank;
    return SNOWTRI_OK;
}

extern "C" int snowtri_comm_destroy(snowtri_t* h) {
    if (!h || !h->nccl_comm) return SNOWTRI_OK;
    cudaSetDevice(h->device);
    if (g_nccl.lib) g_nccl.comm_destroy(h->nccl_comm);
    h->nccl_comm = nullptr;
    h->nccl_nranks = 0;
    return SNOWTRI_OK;
}

extern "C" int snowtri_allgather(snowtri_t* h, const void* d_send, void* d_recv, size_t bytes_per_rank,
                                 void* nccl_comm_or_null, void* stream) {
    if (!h) return fail(nullptr, SNOWTRI_E_ARG, "snowtri_allgather: NULL handle");
    if (bytes_per_rank == 0) return SNOWTRI_OK;
    if (!d_send || !d_recv) return fail(h, SNOWTRI_E_ARG, "snowtri_allgather: NULL buffer");
    void* comm = nccl_comm_or_null ? nccl_comm_or_null : h->nccl_comm;
    if (!comm) return fail(h, SNOWTRI_E_ARG, "snowtri_allgather: no communicator (pass one or call snowtri_comm_init)");
    if (const char* e = load_nccl()) return fail(h, SNOWTRI_E_UNSUPPORTED, "snowtri_allgather: %s", e);
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int rc = g_nccl.all_gather(d_send, d_recv, bytes_per_rank, /*ncclChar*/ 0, comm, (cudaStream_t)stream);
    if (rc) return fail(h, SNOWTRI_E_CUDA, "ncclAllGather: %s", g_nccl.error_string(rc));
    return SNOWTRI_OK;
}
