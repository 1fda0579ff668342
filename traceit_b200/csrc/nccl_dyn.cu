// NCCL plumbing for sharded worlds: one all-gather of float4(x,y,z,m) per step over
// NVLink 5 / NVSwitch.  The reference has no distributed code at all.
//
// NCCL is bound at run time with dlopen so that (a) single-GPU users need no NCCL, and
// (b) inside a Python process that already imported torch, the SAME libnccl.so.2 torch
// loaded is reused (two NCCL copies in one process is asking for trouble).
#include <dlfcn.h>

#include <cstring>

#include "world.cuh"

namespace tr {

// Minimal re-declaration of the NCCL C API we use (stable since NCCL 2.x).
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess == 0
typedef int ncclDataType_t; // ncclFloat32 == 7

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.handle) return TR_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* nm : names) {
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return TR_ERR_NCCL;
    }
#define TR_SYM(field, name)                                         \
    g_nccl.field = (decltype(g_nccl.field))dlsym(h, name);          \
    if (!g_nccl.field) {                                            \
        set_error("libnccl lacks symbol %s", name);                 \
        return TR_ERR_NCCL;                                         \
    }
    TR_SYM(GetUniqueId, "ncclGetUniqueId");
    TR_SYM(CommInitRank, "ncclCommInitRank");
    TR_SYM(CommDestroy, "ncclCommDestroy");
    TR_SYM(AllGather, "ncclAllGather");
    TR_SYM(Broadcast, "ncclBroadcast");
    TR_SYM(GroupStart, "ncclGroupStart");
    TR_SYM(GroupEnd, "ncclGroupEnd");
    TR_SYM(GetErrorString, "ncclGetErrorString");
#undef TR_SYM
    g_nccl.handle = h;
    return TR_OK;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    set_error("NCCL error '%s' in %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return TR_ERR_NCCL;
}

#define TR_NCCL(call)                                        \
    do {                                                     \
        ncclResult_t r__ = (call);                           \
        if (r__ != 0) return nccl_fail(r__, #call);          \
    } while (0)

int nccl_unique_id(void* out128) {
    if (!out128) return TR_ERR_INVALID_ARG;
    TR_TRY(load_nccl());
    ncclUniqueId id;
    TR_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return TR_OK;
}

int nccl_init(tr_world* w, const void* id128) {
    TR_TRY(load_nccl());
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    TR_NCCL(g_nccl.CommInitRank(&comm, w->nranks, id, w->rank));
    w->nccl_comm = comm;
    return TR_OK;
}

// In-place all-gather of the freshly integrated positions: every rank contributes its
// owned block of `buf` and receives everyone else's.  Blocks are ceil(N/R) bodies; when N
// is not a multiple of R the ragged tail is exchanged as a group of broadcasts instead.
int nccl_allgather_pos(tr_world* w, float4* buf) {
    if (w->nranks <= 1) return TR_OK;
    ncclComm_t comm = (ncclComm_t)w->nccl_comm;
    const uint64_t per = (w->n + w->nranks - 1) / w->nranks;
    PendingTimer t;
    timer_begin(w, T_ALLGATHER, &t);
    if (per * (uint64_t)w->nranks == w->n) {
        TR_NCCL(g_nccl.AllGather(buf + w->own_first, buf, per * 4, 7 /*ncclFloat32*/, comm, w->stream));
    } else {
        TR_NCCL(g_nccl.GroupStart());
        for (int r = 0; r < w->nranks; ++r) {
            uint64_t lo = per * (uint64_t)r;
            if (lo >= w->n) break;
            uint64_t cnt = (lo + per <= w->n) ? per : w->n - lo;
            TR_NCCL(g_nccl.Broadcast(buf + lo, buf + lo, cnt * 4, 7, r, comm, w->stream));
        }
        TR_NCCL(g_nccl.GroupEnd());
    }
    timer_end(w, &t);
    return TR_OK;
}

void nccl_destroy(tr_world* w) {
    if (w->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)w->nccl_comm);
    w->nccl_comm = nullptr;
}

}  // namespace tr
