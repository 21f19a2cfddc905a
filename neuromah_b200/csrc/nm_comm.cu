// Data-parallel communication: a thin NCCL binding (dlopen, no link-time dependency)
// for the gradient-slab all-reduce over NVLink 5 / NVSwitch and the FedAvg weight
// average (Federated/utils.py:20-23).  One communicator per process (one process per GPU).
#include "nm_common.cuh"
#include <dlfcn.h>

namespace {

// Minimal NCCL ABI (stable since NCCL 2.x): opaque comm, 128-byte unique id.
typedef struct { char internal[128]; } nm_ncclUniqueId;
typedef void* nm_ncclComm_t;
constexpr int kNcclFloat = 7;   // ncclFloat32
constexpr int kNcclSum = 0;     // ncclSum

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nm_ncclUniqueId*) = nullptr;
    int (*CommInitRank)(nm_ncclComm_t*, int, nm_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(nm_ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nm_ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, nm_ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} g_nccl;

int nccl_fail(int rc, const char* where) {
    snprintf(nm::g_err, sizeof(nm::g_err), "%s: NCCL error %d (%s)", where, rc,
             g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return NM_ERR_NCCL;
}

}  // namespace

#define NM_NCCL(call, where) do { int _r = (call); if (_r != 0) return nccl_fail(_r, where); } while (0)
#define NM_NCCL_LOADED() do { if (!g_nccl.handle) return nm::fail(NM_ERR_NCCL, "%s: %s", __func__, "nm_comm_load() not called"); } while (0)

extern "C" {

int nm_comm_load(const char* libnccl_path) {
    if (g_nccl.handle) return NM_OK;
    const char* path = (libnccl_path && *libnccl_path) ? libnccl_path : "libnccl.so.2";
    void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nm::fail(NM_ERR_NCCL, "dlopen(%s) failed: %s", path, dlerror());
#define NM_SYM(field, name)                                                         \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                      \
    if (!g_nccl.field) { dlclose(h); return nm::fail(NM_ERR_NCCL, "missing symbol %s in %s", name, path); }
    NM_SYM(GetUniqueId, "ncclGetUniqueId")
    NM_SYM(CommInitRank, "ncclCommInitRank")
    NM_SYM(CommDestroy, "ncclCommDestroy")
    NM_SYM(AllReduce, "ncclAllReduce")
    NM_SYM(Broadcast, "ncclBroadcast")
    NM_SYM(GetErrorString, "ncclGetErrorString")
#undef NM_SYM
    g_nccl.handle = h;
    return NM_OK;
}

int nm_comm_unique_id(void* id128) {
    NM_NCCL_LOADED();
    NM_REQUIRE(id128, "id128 is NULL");
    nm_ncclUniqueId id;
    NM_NCCL(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
    return NM_OK;
}

int nm_comm_init(void** comm, const void* id128, int rank, int world) {
    NM_NCCL_LOADED();
    NM_REQUIRE(comm && id128, "null pointer");
    NM_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
    nm_ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    nm_ncclComm_t c = nullptr;
    NM_NCCL(g_nccl.CommInitRank(&c, world, id, rank), "ncclCommInitRank");
    *comm = c;
    return NM_OK;
}

int nm_comm_destroy(void* comm) {
    NM_NCCL_LOADED();
    if (comm) NM_NCCL(g_nccl.CommDestroy((nm_ncclComm_t)comm), "ncclCommDestroy");
    return NM_OK;
}

int nm_comm_allreduce_sum_f32(void* comm, float* buf, size_t n, void* stream) {
    NM_NCCL_LOADED();
    NM_REQUIRE(comm && buf, "null pointer");
    if (!n) return NM_OK;
    NM_NCCL(g_nccl.AllReduce(buf, buf, n, kNcclFloat, kNcclSum, (nm_ncclComm_t)comm, nm::S(stream)), "ncclAllReduce");
    nm::g_launches.fetch_add(1);
    return NM_OK;
}

int nm_comm_broadcast_f32(void* comm, float* buf, size_t n, int root, void* stream) {
    NM_NCCL_LOADED();
    NM_REQUIRE(comm && buf, "null pointer");
    if (!n) return NM_OK;
    NM_NCCL(g_nccl.Broadcast(buf, buf, n, kNcclFloat, root, (nm_ncclComm_t)comm, nm::S(stream)), "ncclBroadcast");
    nm::g_launches.fetch_add(1);
    return NM_OK;
}

}  // extern "C"
