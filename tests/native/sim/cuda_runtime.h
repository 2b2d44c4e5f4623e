// cuda_runtime.h (simulated) -- TEST INFRASTRUCTURE, not the CUDA runtime.
//
// tests/native/sim builds the engine's HOST sources (aloha_b200/csrc/engine.cpp, host.cpp, group.cpp) against
// this header instead of the CUDA toolkit's, and against sim_kernels.cpp instead of the sm_100a kernels, so the
// instruction-stream batcher (symbolic execution, store forwarding, copy-on-write, fusion, levelling, plan cache,
// the DMA paths and the C host driver) can be exercised on a machine without a GPU.  "Device" memory is host
// memory, streams execute at once and in program order, a captured graph is a list of closures.  Nothing under
// aloha_b200/ refers to this directory; the product library links the real runtime and has no CPU path.
#pragma once
#include <cstddef>
#include <cstdint>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __align__(n) __attribute__((aligned(n)))

enum cudaError_t {
    cudaSuccess = 0,
    cudaErrorInvalidValue = 1,
    cudaErrorMemoryAllocation = 2,
    cudaErrorNotReady = 600,
    cudaErrorNotSupported = 801,
    cudaErrorLaunchFailure = 719
};
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeGlobal = 0, cudaStreamCaptureModeThreadLocal = 1 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaEnableDefault = 0 };

typedef struct SimStream *cudaStream_t;
typedef struct SimEvent *cudaEvent_t;
typedef struct SimGraph *cudaGraph_t;
typedef struct SimGraphExec *cudaGraphExec_t;

struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount;
};

const char *cudaGetErrorString(cudaError_t);
cudaError_t cudaGetLastError();
cudaError_t cudaGetDevice(int *);
cudaError_t cudaSetDevice(int);
cudaError_t cudaGetDeviceCount(int *);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *, int);
cudaError_t cudaMalloc(void **, size_t);
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
cudaError_t cudaFree(void *);
cudaError_t cudaHostAlloc(void **, size_t, unsigned);
cudaError_t cudaFreeHost(void *);
cudaError_t cudaMemcpy(void *, const void *, size_t, cudaMemcpyKind);
cudaError_t cudaMemcpyAsync(void *, const void *, size_t, cudaMemcpyKind, cudaStream_t);
cudaError_t cudaMemsetAsync(void *, int, size_t, cudaStream_t);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *, unsigned);
cudaError_t cudaStreamDestroy(cudaStream_t);
cudaError_t cudaStreamSynchronize(cudaStream_t);
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned);
cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode);
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *);
cudaError_t cudaGraphInstantiate(cudaGraphExec_t *, cudaGraph_t, unsigned long long);
cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t);
cudaError_t cudaGraphDestroy(cudaGraph_t);
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *, unsigned);
cudaError_t cudaEventDestroy(cudaEvent_t);
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t);
cudaError_t cudaEventQuery(cudaEvent_t);
cudaError_t cudaGetDriverEntryPoint(const char *, void **, unsigned long long, cudaDriverEntryPointQueryResult *);
