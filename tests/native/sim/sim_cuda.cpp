// sim_cuda.cpp -- the simulated runtime behind tests/native/sim/cuda_runtime.h (TEST INFRASTRUCTURE; see that
// header).  Every stream operation completes before the call returns, so the order of effects is program order.
// Whether the ORDERING the engine asks for would be enough on a real device is checked separately: every operation
// carries a vector clock over the streams (events, stream waits and synchronisations join clocks; the host's own
// knowledge flows into whatever it enqueues next), every device-memory access is logged with it, and two accesses to
// the same bytes from different streams, one of them a write, that are not ordered by those clocks are a violation.
// Device allocations are filled with a poison pattern (a read of never-written device memory shows up as a
// mismatch against the oracle) and tracked, so a free of a foreign pointer or a copy beyond an allocation aborts.
#include "sim.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <algorithm>
#include <mutex>
#include <string>

struct SimStream { int id; };
struct SimEvent { bool recorded = false; std::map<uintptr_t, uint64_t> clock; };
struct SimGraph { std::vector<std::function<void()>> nodes; };
struct SimGraphExec { std::vector<std::function<void()>> nodes; };

namespace sim {

static std::mutex g_mu;
static std::map<const uint8_t *, size_t> g_allocs;           // device allocations: base -> bytes
static thread_local SimGraph *g_capture = nullptr;
static thread_local cudaError_t g_last = cudaSuccess;
static thread_local int g_device = 0;
static std::string g_violation;

[[noreturn]] void die(const char *what) {
    std::fprintf(stderr, "aloha sim: %s\n", what);
    std::abort();
}

bool device_range(const void *p, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_allocs.upper_bound((const uint8_t *)p);
    if (it == g_allocs.begin()) return false;
    --it;
    return (const uint8_t *)p + bytes <= it->first + it->second;
}

// ---- happens-before bookkeeping
typedef std::map<uintptr_t, uint64_t> Clock;               // stream (by handle value; 0 = the legacy stream) -> operations seen
static std::map<uintptr_t, Clock> g_stream_clock;
static Clock g_host_clock;                                 // what the host has waited for
struct Access { uintptr_t lo, hi; bool write; uintptr_t stream; uint64_t tick; const char *what; };
static std::vector<Access> g_log;
static thread_local uintptr_t g_cur_stream = 0;
static thread_local uint64_t g_cur_tick = 0;
static thread_local bool g_in_op = false;

static void join(Clock &into, const Clock &from) {
    for (auto &kv : from) { uint64_t &v = into[kv.first]; v = std::max(v, kv.second); }
}
static void begin_op(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    const uintptr_t s = (uintptr_t)st;
    Clock &c = g_stream_clock[s];
    join(c, g_host_clock);                                 // enqueued after everything the host has observed
    g_cur_stream = s;
    g_cur_tick = ++c[s];
    g_in_op = true;
}
static void host_learns(const Clock &c) {
    join(g_host_clock, c);
    // accesses the host knows to be complete are ordered before everything enqueued from now on
    g_log.erase(std::remove_if(g_log.begin(), g_log.end(), [&](const Access &a) {
        auto it = g_host_clock.find(a.stream);
        return it != g_host_clock.end() && a.tick <= it->second;
    }), g_log.end());
}
void access(const void *p, size_t bytes, bool write, const char *what) {
    static const bool null_device = std::getenv("ALOHA_SIM_NULL") != nullptr;    // host-time measurements: no bookkeeping
    if (!g_in_op || !bytes || null_device) return;
    std::lock_guard<std::mutex> lk(g_mu);
    const uintptr_t lo = (uintptr_t)p, hi = lo + bytes;
    const Clock &mine = g_stream_clock[g_cur_stream];
    for (const Access &a : g_log) {
        if (a.stream == g_cur_stream || a.hi <= lo || hi <= a.lo || !(a.write || write)) continue;
        auto it = mine.find(a.stream);
        if (it == mine.end() || it->second < a.tick) {
            if (g_violation.empty()) g_violation = std::string("stream race (sim): ") + what + " is not ordered after " + a.what + " on another stream";
            std::fprintf(stderr, "aloha sim: stream race: %s (stream %#lx) vs %s (stream %#lx)\n", what, (unsigned long)g_cur_stream, a.what, (unsigned long)a.stream);
        }
    }
    g_log.push_back(Access{lo, hi, write, g_cur_stream, g_cur_tick, what});
    if (g_log.size() > 200000) g_log.erase(g_log.begin(), g_log.begin() + 100000);     // (bounded: old entries first)
}

// One collective over several streams (sim_nccl.cpp): every rank's part is an operation of its own stream; a rank's
// completion is ordered after the START of the ranks it receives from (root < 0: all of them; else: the root only).
void collective(const std::vector<cudaStream_t> &streams, int root, const std::function<void(int rank)> &body) {
    if (g_capture) die("collective inside a stream capture");
    const int n = (int)streams.size();
    std::vector<uint64_t> tick(n);
    std::vector<Clock> started(n);
    for (int r = 0; r < n; ++r) {
        begin_op(streams[r]);
        tick[r] = g_cur_tick;
        std::lock_guard<std::mutex> lk(g_mu);
        started[r] = g_stream_clock[(uintptr_t)streams[r]];
        started[r][(uintptr_t)streams[r]] = tick[r] - 1;      // what had been queued before the collective itself
    }
    for (int r = 0; r < n; ++r) {
        {
            std::lock_guard<std::mutex> lk(g_mu);
            Clock &c = g_stream_clock[(uintptr_t)streams[r]];
            for (int s = 0; s < n; ++s)
                if (s != r && (root < 0 || s == root) && !(root >= 0 && r == root)) join(c, started[s]);
        }
        g_cur_stream = (uintptr_t)streams[r];
        g_cur_tick = tick[r];
        g_in_op = true;
        body(r);
        g_in_op = false;
    }
}

void enqueue(cudaStream_t st, std::function<void()> fn) {
    if (g_capture) { g_capture->nodes.push_back(std::move(fn)); return; }
    begin_op(st);
    fn();
    g_in_op = false;
}
bool capturing() { return g_capture != nullptr; }
void violation(const char *what) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_violation.empty()) g_violation = std::string("launch failure (sim): ") + what;
    std::fprintf(stderr, "aloha sim: kernel contract violated: %s\n", what);
}
cudaError_t status() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_violation.empty() ? cudaSuccess : cudaErrorLaunchFailure;
}

static CUresult encode_tiled(CUtensorMap *map, CUtensorMapDataType, cuuint32_t rank, void *base, const cuuint64_t *dims,
                             const cuuint64_t *strides, const cuuint32_t *box, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle swz, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (rank != 2 || dims[0] != 16 || strides[0] != 128 || box[0] != 16 || box[1] != 16 || swz != CU_TENSOR_MAP_SWIZZLE_128B)
        return CUDA_ERROR_INVALID_VALUE;
    if (((uintptr_t)base & 127) || !device_range(base, dims[1] * 128)) return CUDA_ERROR_INVALID_VALUE;
    std::memset(map, 0, sizeof *map);
    map->opaque[0] = (uint64_t)(uintptr_t)base;
    map->opaque[1] = dims[1];
    return CUDA_SUCCESS;
}

}  // namespace sim

using namespace sim;

// a new test starts from a clean slate (tests/sim_engine.py simulated())
extern "C" void sim_reset_violation() {
    std::lock_guard<std::mutex> lk(sim::g_mu);
    sim::g_violation.clear();
}

const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument (sim)";
    case cudaErrorMemoryAllocation: return "out of memory (sim)";
    case cudaErrorNotReady: return "not ready (sim)";
    case cudaErrorLaunchFailure: return g_violation.empty() ? "launch failure (sim)" : g_violation.c_str();
    default: return "error (sim)";
    }
}
cudaError_t cudaGetLastError() { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
cudaError_t cudaGetDevice(int *d) { *d = g_device; return cudaSuccess; }
// ALOHA_SIM_DEVICES (default 1): how many devices the box pretends to have; they all share the host's memory
static int device_count() {
    const char *e = std::getenv("ALOHA_SIM_DEVICES");
    const int n = e ? std::atoi(e) : 1;
    return n < 1 ? 1 : n;
}
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= device_count()) return cudaErrorInvalidValue; g_device = d; return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = device_count(); return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int d) {
    if (d < 0 || d >= device_count()) return cudaErrorInvalidValue;
    std::memset(p, 0, sizeof *p);
    std::snprintf(p->name, sizeof p->name, "simulated sm_100 (host memory)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void **out, size_t bytes) {
    void *p = nullptr;
    if (posix_memalign(&p, 256, bytes ? bytes : 256)) return cudaErrorMemoryAllocation;
    uint64_t *w = (uint64_t *)p;
    for (size_t i = 0; i < bytes / 8; ++i) w[i] = 0xDEADBEEFCAFEF00Dull ^ (i * 0x9E3779B97F4A7C15ull);
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs[(const uint8_t *)p] = bytes;
    *out = p;
    return cudaSuccess;
}
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_allocs.find((const uint8_t *)p);
        if (it == g_allocs.end()) die("cudaFree of a pointer cudaMalloc did not return");
        const uintptr_t lo = (uintptr_t)it->first, hi = lo + it->second;
        g_allocs.erase(it);
        // cudaFree waits for the device: nothing touching the block can still be in flight
        g_log.erase(std::remove_if(g_log.begin(), g_log.end(), [&](const Access &a) { return a.lo < hi && lo < a.hi; }), g_log.end());
    }
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **out, size_t bytes, unsigned) {
    *out = std::malloc(bytes ? bytes : 1);
    return *out ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }

static cudaError_t copy(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t st) {
    if (capturing()) die("memcpy inside a stream capture");
    const bool dev_dst = kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice;
    const bool dev_src = kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice;
    if (dev_dst && !device_range(dst, n)) die("copy beyond a device allocation (dst)");
    if (dev_src && !device_range(src, n)) die("copy beyond a device allocation (src)");
    static const bool null_device = std::getenv("ALOHA_SIM_NULL") != nullptr;    // (see sim_kernels.cpp: host time only)
    enqueue(st, [=]() {
        if (dev_src) access(src, n, false, "memcpy source");
        if (dev_dst) access(dst, n, true, "memcpy destination");
        if (!null_device || n <= 65536) std::memmove(dst, src, n);            // (small copies are tables: keep them)
    });
    return status();
}
static void host_waits_for(cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_mu);
    host_learns(g_stream_clock[(uintptr_t)st]);
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) {
    const cudaError_t e = copy(d, s, n, k, nullptr);       // the legacy stream, and the host waits for it
    host_waits_for(nullptr);
    return e;
}
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st) { return copy(d, s, n, k, st); }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t st) {
    if (!device_range(d, n)) die("memset beyond a device allocation");
    enqueue(st, [=]() { access(d, n, true, "memset"); std::memset(d, v, n); });
    return status();
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { static int next = 1; *s = new SimStream{next++}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    host_waits_for(s);                                      // (destruction waits for the stream's work)
    { std::lock_guard<std::mutex> lk(g_mu); g_stream_clock.erase((uintptr_t)s); }
    delete s;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t s) {
    if (capturing()) die("synchronize inside a stream capture");
    host_waits_for(s);
    return status();
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    if (!e) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lk(g_mu);
    join(g_stream_clock[(uintptr_t)s], e->clock);           // (an event never recorded orders nothing, as on the device)
    return cudaSuccess;
}
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) {
    if (g_capture) return cudaErrorInvalidValue;
    g_capture = new SimGraph();
    return cudaSuccess;
}
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) {
    if (!g_capture) return cudaErrorInvalidValue;
    *g = g_capture;
    g_capture = nullptr;
    return cudaSuccess;
}
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *x, cudaGraph_t g, unsigned long long) {
    if (!g) return cudaErrorInvalidValue;
    *x = new SimGraphExec{g->nodes};
    return cudaSuccess;
}
cudaError_t cudaGraphLaunch(cudaGraphExec_t x, cudaStream_t st) {
    if (!x) return cudaErrorInvalidValue;
    for (auto &fn : x->nodes) { begin_op(st); fn(); g_in_op = false; }      // the nodes of a captured stream run in order
    return status();
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t x) { delete x; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new SimEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mu);
    Clock &c = g_stream_clock[(uintptr_t)s];
    join(c, g_host_clock);
    e->recorded = true;
    e->clock = c;
    return cudaSuccess;
}
// A query may say "complete" or "not yet" for work the host has not waited for: both happen on a device.  The
// default here is the slow device (not yet, until the host has synchronised past the event), which keeps the
// caller on its stream-wait path; ALOHA_SIM_EAGER_EVENTS=1 is the fast one (complete, and the host learns it).
cudaError_t cudaEventQuery(cudaEvent_t e) {
    if (!e->recorded) return cudaErrorNotReady;
    std::lock_guard<std::mutex> lk(g_mu);
    bool known = true;
    for (auto &kv : e->clock) {
        auto it = g_host_clock.find(kv.first);
        if (it == g_host_clock.end() || it->second < kv.second) known = false;
    }
    if (known) return cudaSuccess;
    const char *eager = std::getenv("ALOHA_SIM_EAGER_EVENTS");
    if (!eager || eager[0] != '1') return cudaErrorNotReady;
    host_learns(e->clock);
    return cudaSuccess;
}
cudaError_t cudaGetDriverEntryPoint(const char *name, void **fn, unsigned long long, cudaDriverEntryPointQueryResult *q) {
    if (std::strcmp(name, "cuTensorMapEncodeTiled") == 0) {
        *fn = (void *)&sim::encode_tiled;
        *q = cudaDriverEntryPointSuccess;
        return cudaSuccess;
    }
    *fn = nullptr;
    *q = cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}
