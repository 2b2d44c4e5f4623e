// sim_nccl.cpp -- the few NCCL entry points aloha_b200/csrc/group.cpp binds, for the simulated device (TEST
// INFRASTRUCTURE; see cuda_runtime.h in this directory).  Built as a library whose SONAME is libnccl.so.2 and loaded
// before a group is created, so that group.cpp's dlopen("libnccl.so.2") finds it.
//
// Only communicators whose ranks all live in ONE process exist here (ncclCommInitAll, or ncclCommInitRank with one
// rank): the ranks' calls arrive between ncclGroupStart and ncclGroupEnd, the k-th collective of every rank belongs
// together, and it is carried out when the group closes -- a copy between the ranks' "device" buffers, reported to the
// runtime's stream bookkeeping as one operation per rank:
//   all-gather: nobody finishes before everybody has started (every rank receives from every rank);
//   broadcast:  a receiver finishes after the root has started; the root waits for nobody.
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <vector>

#include "sim.hpp"

namespace sim {
// (defined in sim_cuda.cpp) one collective: stream of every rank, whom each rank's completion depends on, the copies
void collective(const std::vector<cudaStream_t> &streams, int root, const std::function<void(int rank)> &body);
}

extern "C" {

typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0, ncclInvalidArgument = 4, ncclInvalidUsage = 5 };

struct Clique { int n; };
struct ncclComm { std::shared_ptr<Clique> clique; int rank; };
typedef ncclComm *ncclComm_t;

namespace {
struct Call { int kind; ncclComm_t comm; const void *send; void *recv; size_t bytes; int root; cudaStream_t stream; };
thread_local int g_depth = 0;
thread_local std::vector<Call> g_calls;

int flush() {
    // per clique: the k-th call of each rank
    std::map<Clique *, std::map<int, std::vector<Call>>> by;
    for (auto &c : g_calls) by[c.comm->clique.get()][c.comm->rank].push_back(c);
    g_calls.clear();
    for (auto &kv : by) {
        const int n = kv.first->n;
        if ((int)kv.second.size() != n) { std::fprintf(stderr, "sim nccl: a collective was not issued on every rank\n"); return ncclInvalidUsage; }
        const size_t rounds = kv.second.begin()->second.size();
        for (auto &r : kv.second) if (r.second.size() != rounds) return ncclInvalidUsage;
        for (size_t k = 0; k < rounds; ++k) {
            std::vector<Call> c(n);
            std::vector<cudaStream_t> streams(n);
            for (int r = 0; r < n; ++r) { c[r] = kv.second[r][k]; streams[r] = c[r].stream; }
            for (int r = 1; r < n; ++r)
                if (c[r].kind != c[0].kind || c[r].bytes != c[0].bytes || c[r].root != c[0].root) return ncclInvalidUsage;
            if (c[0].kind == 0) {                          // all-gather: rank r receives every rank's block, in rank order
                std::vector<std::vector<uint8_t>> blocks(n);
                for (int r = 0; r < n; ++r) blocks[r].assign((const uint8_t *)c[r].send, (const uint8_t *)c[r].send + c[r].bytes);
                sim::collective(streams, -1, [&](int r) {
                    sim::access(c[r].send, c[r].bytes, false, "all-gather send block");
                    for (int s = 0; s < n; ++s) {
                        uint8_t *slot = (uint8_t *)c[r].recv + (size_t)s * c[r].bytes;
                        if (s == r && slot == c[r].send) continue;       // in place: the rank's own block is not rewritten
                        sim::access(slot, c[r].bytes, true, "all-gather receive block");
                        std::memcpy(slot, blocks[s].data(), c[r].bytes);
                    }
                });
            } else {                                       // broadcast from root
                const int root = c[0].root;
                if (root < 0 || root >= n) return ncclInvalidArgument;
                std::vector<uint8_t> block((const uint8_t *)c[root].send, (const uint8_t *)c[root].send + c[root].bytes);
                sim::collective(streams, root, [&](int r) {
                    if (r == root) sim::access(c[r].send, c[r].bytes, false, "broadcast source");
                    if (r != root || c[r].recv != c[r].send) {
                        sim::access(c[r].recv, c[r].bytes, true, "broadcast destination");
                        std::memcpy(c[r].recv, block.data(), c[r].bytes);
                    }
                });
            }
        }
    }
    return sim::status() == cudaSuccess ? ncclSuccess : ncclInvalidUsage;
}
int submit(const Call &c) {
    if (!c.comm || !sim::device_range(c.recv, 1)) return ncclInvalidArgument;
    g_calls.push_back(c);
    return g_depth ? ncclSuccess : flush();
}
}  // namespace

int ncclGetUniqueId(ncclUniqueId *id) { std::memset(id, 0x5a, sizeof *id); return ncclSuccess; }
int ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId, int rank) {
    if (nranks != 1 || rank != 0) {
        std::fprintf(stderr, "sim nccl: communicators across processes do not exist on the simulated device\n");
        return ncclInvalidUsage;
    }
    *comm = new ncclComm{std::make_shared<Clique>(Clique{1}), 0};
    return ncclSuccess;
}
int ncclCommInitAll(ncclComm_t *comms, int n, const int *) {
    auto q = std::make_shared<Clique>(Clique{n});
    for (int r = 0; r < n; ++r) comms[r] = new ncclComm{q, r};
    return ncclSuccess;
}
int ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }
int ncclGroupStart() { ++g_depth; return ncclSuccess; }
int ncclGroupEnd() {
    if (g_depth <= 0) return ncclInvalidUsage;
    return --g_depth ? ncclSuccess : flush();
}
int ncclAllGather(const void *send, void *recv, size_t count, int dtype, ncclComm_t comm, cudaStream_t st) {
    if (dtype != 5) return ncclInvalidArgument;            // ncclUint64
    return submit(Call{0, comm, send, recv, count * 8, 0, st});
}
int ncclBroadcast(const void *send, void *recv, size_t count, int dtype, int root, ncclComm_t comm, cudaStream_t st) {
    if (dtype != 5) return ncclInvalidArgument;
    return submit(Call{1, comm, send, recv, count * 8, root, st});
}
const char *ncclGetErrorString(int e) { return e == ncclSuccess ? "no error" : "error (sim nccl)"; }

}  // extern "C"
