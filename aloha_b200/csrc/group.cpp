// group.cpp -- several engines (one per B200) as one limb-sharded machine: NCCL collectives over SPM rows
// behind the C-ABI (include/aloha_b200.h, aloha_group_*).
//
// The reference is a single chip; its only data movers are the AXI DMA channels between DDR and the
// scratchpad (src/mem_buf/axi_data_rd_top.sv, axi_data_wr_top.sv; driven by
// sim/top/top_noaxilite_tb.sv:372-394,450-520).  A limb-sharded machine adds one more mover of the same
// shape -- scratchpad rows to scratchpad rows of the peer chips -- which is what these entry points are:
// they name SPM rows, run asynchronously beside the VP like the DMA block does, and are ordered against
// run_vp by the same start/done discipline (here: CUDA events between the engine's stream and the group's
// communication stream).  The only cross-limb step of the path is the base extension of the key-switch
// stream: an all-gather of the coefficient-form digits and a broadcast of the special-prime
// accumulators.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): a process that never creates a group does not need
// it, and a process that already carries an NCCL (torch) shares that copy.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "engine.hpp"

using namespace alb;

namespace {

// ---- the handful of NCCL entry points used, declared as in nccl.h (stable ABI since 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclUint64 = 5 };
struct Nccl {
    void *so = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

Nccl &nccl() {
    static Nccl N;
    if (N.so || !N.error.empty()) return N;
    // ALOHA_NCCL_LIB: a particular NCCL build (full path), tried before the system's
    const char *env = std::getenv("ALOHA_NCCL_LIB");
    const char *names[] = {env && *env ? env : "libnccl.so.2", "libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
        if ((N.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!N.so) {
        N.error = std::string("cannot load libnccl.so.2: ") + dlerror();
        return N;
    }
    auto sym = [&](const char *name) {
        void *p = dlsym(N.so, name);
        if (!p && N.error.empty()) N.error = std::string("libnccl lacks ") + name;
        return p;
    };
    N.GetUniqueId = (decltype(N.GetUniqueId))sym("ncclGetUniqueId");
    N.CommInitRank = (decltype(N.CommInitRank))sym("ncclCommInitRank");
    N.CommInitAll = (decltype(N.CommInitAll))sym("ncclCommInitAll");
    N.CommDestroy = (decltype(N.CommDestroy))sym("ncclCommDestroy");
    N.AllGather = (decltype(N.AllGather))sym("ncclAllGather");
    N.Broadcast = (decltype(N.Broadcast))sym("ncclBroadcast");
    N.GroupStart = (decltype(N.GroupStart))sym("ncclGroupStart");
    N.GroupEnd = (decltype(N.GroupEnd))sym("ncclGroupEnd");
    N.GetErrorString = (decltype(N.GetErrorString))sym("ncclGetErrorString");
    return N;
}

struct Member {
    aloha *E = nullptr;
    ncclComm_t comm = nullptr;
    cudaStream_t cstream = nullptr;          // communication stream of this member's device
    std::vector<cudaEvent_t> arrived;        // per source rank: its block of the last all-gather is in this SPM
    cudaEvent_t bcast_done = nullptr;        // the last broadcast has landed
    cudaEvent_t ready = nullptr;             // engine work the next transfer depends on
};

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct aloha_group {
    std::vector<Member> m;     // 1 member (this process's rank of a multi-process group) or all of them (local)
    int rank = 0, nranks = 1;  // rank of m[0]; local groups: m[i] has rank i
    bool local = false;
    std::string last_error;
};

namespace {

int gfail(aloha_group *G, int code, const std::string &msg) {
    G->last_error = msg;
    for (auto &mb : G->m) if (mb.E) mb.E->last_error = msg;
    return code;
}
#define NC(call)                                                                                        \
    do {                                                                                                \
        int r_ = (call);                                                                                \
        if (r_ != ncclSuccess) return gfail(G, ALOHA_E_CUDA, std::string(#call) + ": " + nccl().GetErrorString(r_)); \
    } while (0)
#define GCU(call)                                                                                       \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return gfail(G, ALOHA_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

int init_member(aloha_group *G, Member &mb, int nranks) {
    DevGuard dg(mb.E->device);
    GCU(cudaStreamCreateWithFlags(&mb.cstream, cudaStreamNonBlocking));
    mb.arrived.resize(nranks);
    for (auto &e : mb.arrived) GCU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    GCU(cudaEventCreateWithFlags(&mb.bcast_done, cudaEventDisableTiming));
    GCU(cudaEventCreateWithFlags(&mb.ready, cudaEventDisableTiming));
    return ALOHA_OK;
}

// Rows [row, row + nrows) of every member's SPM are about to be overwritten by a transfer: registers
// aliasing them move out first, the 'x' tracking learns about the write, and the transfer is ordered after
// everything queued on the engine so far.
int before_transfer(aloha_group *G, Member &mb, uint32_t row, uint64_t nrows) {
    if ((uint64_t)row + nrows > mb.E->cfg.spm_rows) return gfail(G, ALOHA_E_RANGE, "collective rows beyond SPM");
    int rc = aloha_spm_mark_written(mb.E, row, (uint32_t)nrows);
    if (rc) return rc;
    DevGuard dg(mb.E->device);
    GCU(cudaEventRecord(mb.ready, mb.E->stream));
    GCU(cudaStreamWaitEvent(mb.cstream, mb.ready, 0));
    return ALOHA_OK;
}

}  // namespace

extern "C" {

int aloha_group_unique_id(uint8_t id[ALOHA_GROUP_ID_BYTES]) {
    if (!id) return ALOHA_E_ARG;
    Nccl &N = nccl();
    if (!N.error.empty()) return ALOHA_E_CUDA;
    ncclUniqueId u;
    if (N.GetUniqueId(&u) != ncclSuccess) return ALOHA_E_CUDA;
    static_assert(sizeof u == ALOHA_GROUP_ID_BYTES, "ncclUniqueId is 128 bytes");
    std::memcpy(id, &u, sizeof u);
    return ALOHA_OK;
}

const char *aloha_group_last_error(const aloha_group_t *G) {
    if (G) return G->last_error.c_str();
    return nccl().error.empty() ? "null group" : nccl().error.c_str();
}

int aloha_group_create(aloha_t *E, const uint8_t id[ALOHA_GROUP_ID_BYTES], int rank, int nranks, aloha_group_t **out) {
    if (!E || !id || !out || nranks < 1 || rank < 0 || rank >= nranks) return ALOHA_E_ARG;
    aloha_group *G = new aloha_group();
    *out = G;
    G->rank = rank;
    G->nranks = nranks;
    G->m.resize(1);
    G->m[0].E = E;
    Nccl &N = nccl();
    if (!N.error.empty()) return gfail(G, ALOHA_E_CUDA, N.error);
    int rc = init_member(G, G->m[0], nranks);
    if (rc) return rc;
    DevGuard dg(E->device);
    ncclUniqueId u;
    std::memcpy(&u, id, sizeof u);
    NC(N.CommInitRank(&G->m[0].comm, nranks, u, rank));
    return ALOHA_OK;
}

int aloha_group_create_local(aloha_t *const *engines, int n, aloha_group_t **out) {
    if (!engines || !out || n < 1) return ALOHA_E_ARG;
    for (int i = 0; i < n; ++i) if (!engines[i]) return ALOHA_E_ARG;
    aloha_group *G = new aloha_group();
    *out = G;
    G->local = true;
    G->nranks = n;
    G->m.resize(n);
    std::vector<int> devs(n);
    for (int i = 0; i < n; ++i) {
        G->m[i].E = engines[i];
        devs[i] = engines[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return gfail(G, ALOHA_E_ARG, "a local group needs one engine per device");
    }
    Nccl &N = nccl();
    if (!N.error.empty()) return gfail(G, ALOHA_E_CUDA, N.error);
    for (auto &mb : G->m) {
        int rc = init_member(G, mb, n);
        if (rc) return rc;
    }
    std::vector<ncclComm_t> comms(n);
    NC(N.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) G->m[i].comm = comms[i];
    return ALOHA_OK;
}

void aloha_group_destroy(aloha_group_t *G) {
    if (!G) return;
    for (auto &mb : G->m) {
        if (!mb.E) continue;
        DevGuard dg(mb.E->device);
        if (mb.cstream) cudaStreamSynchronize(mb.cstream);
        if (mb.comm) nccl().CommDestroy(mb.comm);
        for (auto &e : mb.arrived) if (e) cudaEventDestroy(e);
        if (mb.bcast_done) cudaEventDestroy(mb.bcast_done);
        if (mb.ready) cudaEventDestroy(mb.ready);
        if (mb.cstream) cudaStreamDestroy(mb.cstream);
    }
    delete G;
}

int aloha_group_size(const aloha_group_t *G) { return G ? G->nranks : ALOHA_E_ARG; }
int aloha_group_rank(const aloha_group_t *G) { return G ? G->rank : ALOHA_E_ARG; }

// `count` all-gathers as one transfer, the c-th over the rows starting at spm_row + c * stride_rows: rank r
// contributes rows [start + r * rows_per_rank, + rows_per_rank) of its own SPM; afterwards every member holds
// all nranks blocks at the same rows.  ALOHA_GROUP_CHUNKED moves the blocks one source rank at a time (nranks
// grouped broadcasts), each with its own completion event, so that aloha_group_wait(source) lets the engine
// start on the blocks that have arrived while the rest are still in flight.
int aloha_group_all_gather_rows(aloha_group_t *G, uint32_t spm_row, uint32_t rows_per_rank, uint32_t count,
                                uint32_t stride_rows, uint32_t flags) {
    if (!G || !rows_per_rank || !count) return ALOHA_E_ARG;
    Nccl &N = nccl();
    const uint64_t total = (uint64_t)rows_per_rank * G->nranks, words = (uint64_t)rows_per_rank * kLanes;
    if (count > 1 && stride_rows < total) return gfail(G, ALOHA_E_ARG, "all-gather blocks overlap (stride_rows < nranks * rows_per_rank)");
    for (auto &mb : G->m)
        for (uint32_t c = 0; c < count; ++c) {
            int rc = before_transfer(G, mb, spm_row + c * stride_rows, total);
            if (rc) return rc;
        }
    const bool grouped = G->local || count > 1;
    auto block = [&](Member &mb, uint32_t c, int src) {
        return mb.E->d_spm + ((u64)spm_row + (u64)c * stride_rows) * kLanes + (u64)src * words;
    };
    if (flags & ALOHA_GROUP_CHUNKED) {
        for (int src = 0; src < G->nranks; ++src) {
            if (grouped) NC(N.GroupStart());
            for (auto &mb : G->m) {
                DevGuard dg(mb.E->device);
                for (uint32_t c = 0; c < count; ++c)
                    NC(N.Broadcast(block(mb, c, src), block(mb, c, src), words, ncclUint64, src, mb.comm, mb.cstream));
            }
            if (grouped) NC(N.GroupEnd());
            for (auto &mb : G->m) {
                DevGuard dg(mb.E->device);
                GCU(cudaEventRecord(mb.arrived[src], mb.cstream));
            }
        }
    } else {
        if (grouped) NC(N.GroupStart());
        for (size_t i = 0; i < G->m.size(); ++i) {
            Member &mb = G->m[i];
            DevGuard dg(mb.E->device);
            const int r = G->local ? (int)i : G->rank;
            for (uint32_t c = 0; c < count; ++c)
                NC(N.AllGather(block(mb, c, r), block(mb, c, 0), words, ncclUint64, mb.comm, mb.cstream));
        }
        if (grouped) NC(N.GroupEnd());
        for (auto &mb : G->m) {
            DevGuard dg(mb.E->device);
            for (auto &e : mb.arrived) GCU(cudaEventRecord(e, mb.cstream));
        }
    }
    return ALOHA_OK;
}

int aloha_group_broadcast_rows(aloha_group_t *G, uint32_t spm_row, uint32_t nrows, int root) {
    if (!G || !nrows || root < 0 || root >= G->nranks) return ALOHA_E_ARG;
    Nccl &N = nccl();
    for (auto &mb : G->m) {
        int rc = before_transfer(G, mb, spm_row, nrows);
        if (rc) return rc;
    }
    if (G->local) NC(N.GroupStart());
    for (auto &mb : G->m) {
        DevGuard dg(mb.E->device);
        u64 *p = mb.E->d_spm + (u64)spm_row * kLanes;
        NC(N.Broadcast(p, p, (u64)nrows * kLanes, ncclUint64, root, mb.comm, mb.cstream));
    }
    if (G->local) NC(N.GroupEnd());
    for (auto &mb : G->m) {
        DevGuard dg(mb.E->device);
        GCU(cudaEventRecord(mb.bcast_done, mb.cstream));
    }
    return ALOHA_OK;
}

// Engine work issued after this call starts only when the named transfer has landed:
// source >= 0: that rank's block of the most recent all-gather; ALOHA_GROUP_ALL: every block;
// ALOHA_GROUP_BCAST: the most recent broadcast.
int aloha_group_wait(aloha_group_t *G, int source) {
    if (!G || source >= G->nranks || source < ALOHA_GROUP_BCAST) return ALOHA_E_ARG;
    for (auto &mb : G->m) {
        // queued (deferred) engine calls were issued before this wait: they must not end up behind it
        int rc = aloha_flush(mb.E);
        if (rc) return rc;
        DevGuard dg(mb.E->device);
        if (source == ALOHA_GROUP_BCAST) GCU(cudaStreamWaitEvent(mb.E->stream, mb.bcast_done, 0));
        else if (source == ALOHA_GROUP_ALL) { for (auto &e : mb.arrived) GCU(cudaStreamWaitEvent(mb.E->stream, e, 0)); }
        else GCU(cudaStreamWaitEvent(mb.E->stream, mb.arrived[source], 0));
    }
    return ALOHA_OK;
}

}  // extern "C"
