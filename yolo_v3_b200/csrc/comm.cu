// Multi-GPU plumbing of the path (absent in the reference): images are independent, so
// the batch is sharded across ranks and the only collectives are one broadcast of the packed weight
// blob and a per-batch all-gather of the fixed-capacity detection rows.  NCCL is bound at run time
// with dlopen so that libyolo_b200.so loads (and its symbols can be checked) on a host without
// NCCL or a GPU; in a torch process the already-loaded libnccl.so.2 is reused.
#include <dlfcn.h>

#include <cstring>
#include <string>

#include "yb_internal.h"

namespace yb {
namespace {

typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
typedef int ncclResult_i;   // 0 == ncclSuccess

struct Nccl {
    void* h = nullptr;
    ncclResult_i (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
    ncclResult_i (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int) = nullptr;
    ncclResult_i (*CommDestroy)(ncclComm_p) = nullptr;
    ncclResult_i (*CommAbort)(ncclComm_p) = nullptr;          // optional
    ncclResult_i (*Broadcast)(const void*, void*, size_t, int /*dtype*/, int, ncclComm_p, cudaStream_t) = nullptr;
    ncclResult_i (*AllGather)(const void*, void*, size_t, int /*dtype*/, ncclComm_p, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_i) = nullptr;
    std::string err;
};

// Bound once per process; the function-local static makes the first call thread-safe.
Nccl* nccl() {
    static Nccl n = [] {
        Nccl n;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            n.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.h) break;
        }
        if (!n.h) { n.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return n; }
#define YB_SYM(field, name)                                                         \
    *reinterpret_cast<void**>(&n.field) = dlsym(n.h, name);                          \
    if (!n.field) { n.err = std::string("libnccl lacks ") + name; n.h = nullptr; return n; }
        YB_SYM(GetUniqueId, "ncclGetUniqueId")
        YB_SYM(CommInitRank, "ncclCommInitRank")
        YB_SYM(CommDestroy, "ncclCommDestroy")
        YB_SYM(Broadcast, "ncclBroadcast")
        YB_SYM(AllGather, "ncclAllGather")
        YB_SYM(GetErrorString, "ncclGetErrorString")
#undef YB_SYM
        *reinterpret_cast<void**>(&n.CommAbort) = dlsym(n.h, "ncclCommAbort");
        return n;
    }();
    return &n;
}

constexpr int kNcclChar = 0;   // ncclInt8 / ncclChar

int check(Nccl* n, ncclResult_i r, const char* what, std::string& err) {
    if (r == 0) return YB_OK;
    err = std::string(what) + ": " + (n->GetErrorString ? n->GetErrorString(r) : "nccl error");
    return YB_E_NCCL;
}

}  // namespace

int comm_unique_id(uint8_t* id, std::string& err) {
    Nccl* n = nccl();
    if (!n->h) { err = n->err; return YB_E_NCCL; }
    ncclUniqueId_t u;
    int rc = check(n, n->GetUniqueId(&u), "ncclGetUniqueId", err);
    if (rc) return rc;
    std::memcpy(id, u.internal, 128);
    return YB_OK;
}

int comm_init(void** comm, const uint8_t* id, int rank, int world, std::string& err) {
    Nccl* n = nccl();
    if (!n->h) { err = n->err; return YB_E_NCCL; }
    ncclUniqueId_t u;
    std::memcpy(u.internal, id, 128);
    ncclComm_p cm = nullptr;
    int rc = check(n, n->CommInitRank(&cm, world, u, rank), "ncclCommInitRank", err);
    if (rc) return rc;
    *comm = cm;
    return YB_OK;
}

int comm_bcast(void* comm, void* buf, size_t bytes, int root, cudaStream_t s, std::string& err) {
    Nccl* n = nccl();
    return check(n, n->Broadcast(buf, buf, bytes, kNcclChar, root, static_cast<ncclComm_p>(comm), s), "ncclBroadcast", err);
}

int comm_allgather(void* comm, const void* send, void* recv, size_t bytes, cudaStream_t s, std::string& err) {
    Nccl* n = nccl();
    return check(n, n->AllGather(send, recv, bytes, kNcclChar, static_cast<ncclComm_p>(comm), s), "ncclAllGather", err);
}

// Tear-down must never wait for the peers: ranks leave at different times (rank 0 of bench.py keeps working alone after
// the others have gone), and ncclCommDestroy synchronises with the other ranks of the communicator -- observed in round 2
// as a deadlock against a rank blocked in an unrelated barrier.  ncclCommAbort frees the communicator without that handshake;
// every collective this library enqueues has completed by the time a context is destroyed (the caller synchronises its streams).
void comm_destroy(void* comm) {
    Nccl* n = nccl();
    if (!n->h || !comm) return;
    if (n->CommAbort) n->CommAbort(static_cast<ncclComm_p>(comm));
    else n->CommDestroy(static_cast<ncclComm_p>(comm));
}

}  // namespace yb
