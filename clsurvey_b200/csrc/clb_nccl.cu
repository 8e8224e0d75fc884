// Data-parallel exchange: one ncclAllReduce(sum, fp32) of the flat gradient buffer per step and one
// of the flat Fisher/Omega buffer per task.  The reference is single-GPU (no collective anywhere); this is new.
// libnccl is resolved at run time (dlopen) so that the library loads on boxes without NCCL and uses the same
// libnccl.so.2 that torch already mapped into the process.
#include <dlfcn.h>
#include <nccl.h>

#include "clb_common.cuh"

namespace clb {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return CLB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cannot dlopen libnccl.so.2: %s", dlerror());
        return CLB_ENCCL;
    }
#define SYM(field, name)                                                        \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                  \
    if (!g_nccl.field) { set_error("libnccl: missing symbol %s", name); return CLB_ENCCL; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(AllReduce, "ncclAllReduce")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return CLB_OK;
}
#define CLB_NCCL(call)                                                                       \
    do {                                                                                     \
        ncclResult_t r__ = (call);                                                           \
        if (r__ != ncclSuccess) {                                                            \
            set_error("%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
            return CLB_ENCCL;                                                                \
        }                                                                                    \
    } while (0)
}  // namespace clb

using namespace clb;

extern "C" {
int clb_nccl_unique_id(void* out128) {
    CLB_CHECK_ARG(out128 != nullptr);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    CLB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return CLB_OK;
}
int clb_nccl_init(const void* id128, int rank, int world, void** comm_out) {
    CLB_CHECK_ARG(id128 && comm_out && world >= 1 && rank >= 0 && rank < world);
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm;
    CLB_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    *comm_out = comm;
    return CLB_OK;
}
int clb_nccl_allreduce_f32(void* comm, float* buf, int64_t n, void* stream) {
    CLB_CHECK_ARG(comm && buf && n >= 0);
    int rc = nccl_load();
    if (rc) return rc;
    CLB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, (ncclComm_t)comm, as_stream(stream)));
    return CLB_OK;
}
int clb_nccl_destroy(void* comm) {
    CLB_CHECK_ARG(comm != nullptr);
    int rc = nccl_load();
    if (rc) return rc;
    CLB_NCCL(g_nccl.CommDestroy((ncclComm_t)comm));
    return CLB_OK;
}
}
