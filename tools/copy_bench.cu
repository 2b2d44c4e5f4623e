// copy_bench.cu -- what the memory system gives a kernel shaped like the tiled automorphism, piece by piece:
// (a) 16-byte ld/st copy, (b) 8-byte ld/st copy, (c) 8-byte cp.async -> shared -> 8-byte store (the aut kernel's
// data path with the identity map), (d) the same with 16-byte cp.async / stores, (e) = (c) with 16 slots per thread.
// 1 GiB in, 1 GiB out (larger than L2), CUDA-event timed; prints GB/s (read + write).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

__global__ void __launch_bounds__(256) copy16(const ulonglong2 *s, ulonglong2 *d, size_t n) {
    size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (i + k * 256 < n) d[i + k * 256] = s[i + k * 256];
}
__global__ void __launch_bounds__(256) copy8(const u64 *s, u64 *d, size_t n) {
    size_t i = (size_t)blockIdx.x * 2048 + threadIdx.x;
    u64 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = s[i + k * 256];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[i + k * 256] = v[k];
}
template <int ITERS>
__global__ void __launch_bounds__(256) stage8(const u64 *s, u64 *d, size_t n) {
    __shared__ u64 tile[256 * ITERS + 64];
    size_t i = (size_t)blockIdx.x * (256 * ITERS) + threadIdx.x;
    unsigned base = (unsigned)__cvta_generic_to_shared(tile);
#pragma unroll
    for (int k = 0; k < ITERS; ++k)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(base + (threadIdx.x + k * 256) * 8), "l"(s + i + k * 256) : "memory");
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
    __syncthreads();
    // read "transposed": thread t takes word (t * ITERS + k) -- a different word than it loaded
#pragma unroll
    for (int k = 0; k < ITERS; ++k) {
        const int w = (threadIdx.x & 31) * 1 + (threadIdx.x >> 5) * 32 + ((k + (threadIdx.x >> 5)) % ITERS) * 256;
        d[(size_t)blockIdx.x * (256 * ITERS) + w] = tile[w];
    }
}
__global__ void __launch_bounds__(256) stage16(const ulonglong2 *s, ulonglong2 *d, size_t n) {
    __shared__ ulonglong2 tile[1024];
    size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
    unsigned base = (unsigned)__cvta_generic_to_shared(tile);
#pragma unroll
    for (int k = 0; k < 4; ++k)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (threadIdx.x + k * 256) * 16), "l"(s + i + k * 256) : "memory");
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int w = threadIdx.x + ((k + (threadIdx.x >> 5)) & 3) * 256;
        d[(size_t)blockIdx.x * 1024 + w] = tile[w];
    }
}
int main() {
    const size_t words = 1ull << 27;   // 1 GiB of u64
    u64 *s, *d;
    cudaMalloc(&s, words * 8); cudaMalloc(&d, words * 8);
    cudaMemset(s, 1, words * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %7.1f GB/s  (%s)\n", name, 2.0 * words * 8 * 10 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    run("(a) 16-byte ld/st copy", [&] { copy16<<<(unsigned)(words / 2 / 1024), 256>>>((const ulonglong2 *)s, (ulonglong2 *)d, words / 2); });
    run("(b) 8-byte ld/st copy, 8 per thread", [&] { copy8<<<(unsigned)(words / 2048), 256>>>(s, d, words); });
    run("(c) 8-byte cp.async -> smem -> 8-byte st, 8/thr", [&] { stage8<8><<<(unsigned)(words / 2048), 256>>>(s, d, words); });
    run("(e) same, 16 per thread", [&] { stage8<16><<<(unsigned)(words / 4096), 256>>>(s, d, words); });
    run("(f) same, 4 per thread", [&] { stage8<4><<<(unsigned)(words / 1024), 256>>>(s, d, words); });
    run("(d) 16-byte cp.async -> smem -> 16-byte st", [&] { stage16<<<(unsigned)(words / 2 / 1024), 256>>>((const ulonglong2 *)s, (ulonglong2 *)d, words / 2); });
    cudaMemcpy(d, s, words * 8, cudaMemcpyDeviceToDevice);
    run("(m) cudaMemcpy D2D", [&] { cudaMemcpyAsync(d, s, words * 8, cudaMemcpyDeviceToDevice); });
    return 0;
}
