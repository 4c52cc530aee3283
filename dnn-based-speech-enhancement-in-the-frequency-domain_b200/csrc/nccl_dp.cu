// Data-parallel plumbing of the C ABI without torch: one NCCL communicator per process / GPU, one
// sum all-reduce of the flat fp32 gradient buffer per step.  libnccl is bound at RUN time with dlopen (the library links
// only libcudart): `libnccl.so.2` is resolved through the process (torch has it loaded when torch.distributed is in use),
// LD_LIBRARY_PATH, or the path in SEFD_NCCL_LIB.  Declarations below restate the public NCCL C API (nccl.h) for the four
// entry points used.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sefd.h"
#include "common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;       // NCCL_UNIQUE_ID_BYTES
typedef int ncclResult_t;                                   // ncclSuccess = 0
enum { NCCL_SUM = 0, NCCL_FLOAT32 = 7 };

struct Api {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

Api* api() {
    static Api a;
    static bool tried = false;
    if (tried) return a.lib ? &a : nullptr;
    tried = true;
    const char* names[3] = {getenv("SEFD_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (int i = 0; i < 3 && !a.lib; ++i)
        if (names[i] && names[i][0]) a.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!a.lib) return nullptr;
    a.GetUniqueId = (ncclResult_t(*)(ncclUniqueId*))dlsym(a.lib, "ncclGetUniqueId");
    a.CommInitRank = (ncclResult_t(*)(ncclComm_t*, int, ncclUniqueId, int))dlsym(a.lib, "ncclCommInitRank");
    a.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(a.lib, "ncclCommDestroy");
    a.AllReduce = (ncclResult_t(*)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(a.lib, "ncclAllReduce");
    a.GetErrorString = (const char* (*)(ncclResult_t))dlsym(a.lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce) {
        dlclose(a.lib);
        a.lib = nullptr;
        return nullptr;
    }
    return &a;
}

}  // namespace

struct sefd_comm {
    ncclComm_t comm;
    int rank, world;
};

extern "C" {

int sefd_nccl_unique_id_bytes(void) { return 128; }

int sefd_nccl_unique_id(void* out128) {
    Api* a = api();
    SEFD_REQUIRE(a != nullptr, "nccl: libnccl.so.2 could not be loaded (set SEFD_NCCL_LIB or LD_LIBRARY_PATH): %s", dlerror());
    SEFD_REQUIRE(out128 != nullptr, "nccl_unique_id: null buffer");
    ncclUniqueId id;
    const ncclResult_t r = a->GetUniqueId(&id);
    SEFD_REQUIRE(r == 0, "ncclGetUniqueId: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
    memcpy(out128, &id, sizeof(id));
    return 0;
}

sefd_comm* sefd_nccl_init(int rank, int world, const void* unique_id128) {
    Api* a = api();
    if (!a) {
        sefd_set_error("nccl: libnccl.so.2 could not be loaded (set SEFD_NCCL_LIB or LD_LIBRARY_PATH)");
        return nullptr;
    }
    if (world < 1 || rank < 0 || rank >= world || !unique_id128) {
        sefd_set_error("nccl_init: bad rank %d / world %d / id", rank, world);
        return nullptr;
    }
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof(id));
    ncclComm_t c = nullptr;
    const ncclResult_t r = a->CommInitRank(&c, world, id, rank);
    if (r != 0) {
        sefd_set_error("ncclCommInitRank: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
        return nullptr;
    }
    sefd_comm* s = new sefd_comm();
    s->comm = c; s->rank = rank; s->world = world;
    return s;
}

int sefd_nccl_allreduce(sefd_comm* comm, float* buf, long long n, void* stream) {
    SEFD_REQUIRE(comm && buf && n > 0, "nccl_allreduce: bad argument");
    Api* a = api();
    SEFD_REQUIRE(a != nullptr, "nccl: library not loaded");
    const ncclResult_t r = a->AllReduce(buf, buf, (size_t)n, NCCL_FLOAT32, NCCL_SUM, comm->comm, (cudaStream_t)stream);
    SEFD_REQUIRE(r == 0, "ncclAllReduce: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
    return 0;
}

int sefd_nccl_world(const sefd_comm* comm) { return comm ? comm->world : 1; }

void sefd_nccl_destroy(sefd_comm* comm) {
    if (!comm) return;
    Api* a = api();
    if (a) a->CommDestroy(comm->comm);
    delete comm;
}

}  // extern "C"
