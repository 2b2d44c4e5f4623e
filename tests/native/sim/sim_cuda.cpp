// sim_cuda.cpp -- the simulated runtime behind tests/native/sim/cuda_runtime.h (TEST INFRASTRUCTURE; see that
// header).  Every stream operation completes before the call returns, so the order of effects is program order:
// this checks WHAT the engine enqueues, not whether its event ordering between streams is sufficient.
// Device allocations are filled with a poison pattern (a read of never-written device memory shows up as a
// mismatch against the oracle) and tracked, so a free of a foreign pointer or a copy beyond an allocation aborts.
#include "sim.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

struct SimStream { int id; };
struct SimEvent { bool recorded = false; };
struct SimGraph { std::vector<std::function<void()>> nodes; };
struct SimGraphExec { std::vector<std::function<void()>> nodes; };

namespace sim {

static std::mutex g_mu;
static std::map<const uint8_t *, size_t> g_allocs;           // device allocations: base -> bytes
static thread_local SimGraph *g_capture = nullptr;
static thread_local cudaError_t g_last = cudaSuccess;
static int g_device = 0;
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

void enqueue(std::function<void()> fn) {
    if (g_capture) g_capture->nodes.push_back(std::move(fn));
    else fn();
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
        g_allocs.erase(it);
    }
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaHostAlloc(void **out, size_t bytes, unsigned) {
    *out = std::malloc(bytes ? bytes : 1);
    return *out ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }

static cudaError_t copy(void *dst, const void *src, size_t n, cudaMemcpyKind kind) {
    if (capturing()) die("memcpy inside a stream capture");
    if ((kind == cudaMemcpyHostToDevice || kind == cudaMemcpyDeviceToDevice) && !device_range(dst, n)) die("copy beyond a device allocation (dst)");
    if ((kind == cudaMemcpyDeviceToHost || kind == cudaMemcpyDeviceToDevice) && !device_range(src, n)) die("copy beyond a device allocation (src)");
    std::memmove(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) { return copy(d, s, n, k); }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return copy(d, s, n, k); }
cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
    if (!device_range(d, n)) die("memset beyond a device allocation");
    std::memset(d, v, n);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { static int next = 1; *s = new SimStream{next++}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { if (capturing()) die("synchronize inside a stream capture"); return status(); }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t e, unsigned) { return e ? cudaSuccess : cudaErrorInvalidValue; }
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
cudaError_t cudaGraphLaunch(cudaGraphExec_t x, cudaStream_t) {
    if (!x) return cudaErrorInvalidValue;
    for (auto &fn : x->nodes) fn();
    return status();
}
cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t x) { delete x; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new SimEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->recorded = true; return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t e) { return e->recorded ? cudaSuccess : cudaErrorNotReady; }
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
