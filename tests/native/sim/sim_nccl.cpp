// sim_nccl.cpp -- the few NCCL entry points aloha_b200/csrc/group.cpp binds, for the simulated device (TEST
// INFRASTRUCTURE; see cuda_runtime.h in this directory).  Built as a library whose SONAME is libnccl.so.2 and loaded
// before a group is created, so that group.cpp's dlopen("libnccl.so.2") finds it.
//
// Only communicators whose ranks all live in ONE process exist here: ncclCommInitAll (one thread drives every rank,
// as aloha_group_create_local does) or ncclCommInitRank from one thread per rank with a shared id (what one process
// per GPU does on a real box).  The k-th collective of every rank belongs together; whichever rank hands in its part
// last carries it out -- a copy between the ranks' "device" buffers, reported to the runtime's stream bookkeeping as
// one operation per rank:
//   all-gather: nobody finishes before everybody has started (every rank receives from every rank);
//   broadcast:  a receiver finishes after the root has started; the root waits for nobody.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "sim.hpp"

namespace sim {
// (defined in sim_cuda.cpp) one collective: stream of every rank, whom each rank's completion depends on, the copies
void collective(const std::vector<cudaStream_t> &streams, int root, const std::function<void(int rank)> &body);
}

extern "C" {

typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0, ncclInvalidArgument = 4, ncclInvalidUsage = 5 };

struct Call { int kind; const void *send; void *recv; size_t bytes; int root; cudaStream_t stream; };
// The ranks of a communicator: all in one process, driven by one thread (ncclCommInitAll) or by one thread each
// (ncclCommInitRank with the same id).  A collective is carried out by whichever rank submits its part last.
struct Clique {
    int n;
    std::mutex mu;
    std::condition_variable cv;
    std::map<uint64_t, std::map<int, Call>> parts;    // sequence number of the collective -> rank -> its call
    std::map<uint64_t, int> result;                   // finished collectives not yet seen by every rank
    std::map<uint64_t, int> seen;
};
struct ncclComm { std::shared_ptr<Clique> clique; int rank; uint64_t next = 0; };
typedef ncclComm *ncclComm_t;

namespace {
struct Queued { ncclComm_t comm; Call call; };
thread_local int g_depth = 0;
thread_local std::vector<Queued> g_calls;
std::mutex g_ids_mu;
std::map<std::string, std::weak_ptr<Clique>> g_by_id;

int carry_out(int n, std::map<int, Call> &c) {
    std::vector<cudaStream_t> streams(n);
    for (int r = 0; r < n; ++r) streams[r] = c[r].stream;
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
    return sim::status() == cudaSuccess ? ncclSuccess : ncclInvalidUsage;
}

// Hand in every queued call first (so that a single thread driving all ranks never waits for itself), then wait
// for the collectives to have been carried out.  NCCL itself only enqueues; blocking here is the same thing seen
// from a runtime whose streams run at once.
int flush() {
    std::vector<Queued> calls;
    calls.swap(g_calls);
    std::vector<std::pair<ncclComm_t, uint64_t>> mine;
    for (auto &q : calls) {
        Clique &Q = *q.comm->clique;
        const uint64_t seq = q.comm->next++;
        mine.emplace_back(q.comm, seq);
        std::unique_lock<std::mutex> lk(Q.mu);
        auto &parts = Q.parts[seq];
        parts[q.comm->rank] = q.call;
        if ((int)parts.size() == Q.n) {
            lk.unlock();                                   // (the other ranks only wait; nobody touches parts[seq] now)
            const int rc = carry_out(Q.n, parts);
            lk.lock();
            Q.result[seq] = rc;
            Q.parts.erase(seq);
            Q.cv.notify_all();
        }
    }
    int rc = ncclSuccess;
    for (auto &m : mine) {
        Clique &Q = *m.first->clique;
        std::unique_lock<std::mutex> lk(Q.mu);
        if (!Q.cv.wait_for(lk, std::chrono::seconds(120), [&] { return Q.result.count(m.second) != 0; })) {
            std::fprintf(stderr, "sim nccl: rank %d waited 120 s for the other ranks of collective %llu\n", m.first->rank, (unsigned long long)m.second);
            return ncclInvalidUsage;
        }
        if (Q.result[m.second] != ncclSuccess) rc = Q.result[m.second];
        if (++Q.seen[m.second] == Q.n) { Q.result.erase(m.second); Q.seen.erase(m.second); }
    }
    return rc;
}
int submit(ncclComm_t comm, const Call &c) {
    if (!comm || !sim::device_range(c.recv, 1)) return ncclInvalidArgument;
    g_calls.push_back(Queued{comm, c});
    return g_depth ? ncclSuccess : flush();
}
}  // namespace

int ncclGetUniqueId(ncclUniqueId *id) {
    static std::atomic<uint64_t> next{1};
    std::memset(id, 0, sizeof *id);
    const uint64_t v = next++;
    std::memcpy(id->internal, &v, sizeof v);
    std::memcpy(id->internal + 8, "aloha-sim", 9);
    return ncclSuccess;
}
int ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
    if (nranks < 1 || rank < 0 || rank >= nranks) return ncclInvalidArgument;
    std::lock_guard<std::mutex> lk(g_ids_mu);                 // ranks = threads of this process that share the id
    std::shared_ptr<Clique> q = g_by_id[std::string(id.internal, sizeof id.internal)].lock();
    if (!q) {
        q = std::make_shared<Clique>();
        q->n = nranks;
        g_by_id[std::string(id.internal, sizeof id.internal)] = q;
    }
    if (q->n != nranks) return ncclInvalidArgument;
    *comm = new ncclComm{q, rank};
    return ncclSuccess;
}
int ncclCommInitAll(ncclComm_t *comms, int n, const int *) {
    auto q = std::make_shared<Clique>();
    q->n = n;
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
    return submit(comm, Call{0, send, recv, count * 8, 0, st});
}
int ncclBroadcast(const void *send, void *recv, size_t count, int dtype, int root, ncclComm_t comm, cudaStream_t st) {
    if (dtype != 5) return ncclInvalidArgument;
    return submit(comm, Call{1, send, recv, count * 8, root, st});
}
const char *ncclGetErrorString(int e) { return e == ncclSuccess ? "no error" : "error (sim nccl)"; }

}  // extern "C"
