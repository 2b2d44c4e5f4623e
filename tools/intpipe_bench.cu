// intpipe_bench.cu -- measures the B200 integer issue ceilings the NTT kernels are bounded by:
// thread-instructions per clock per SM for IMAD.WIDE.U32 (fma pipe), IADD3 (alu pipe), a 1:1 mix,
// and the 64x64->hi multiply the Shoup butterfly uses.  MEASURED_PEAKS.json has no integer figure;
// DESIGN.md quotes the numbers this prints.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int ITERS = 4096, ILP = 8;

__global__ void k_imad_wide(u64 *out, u32 a, u32 b) {
    u64 acc[ILP];
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + i), "r"(b));
    }
    u64 s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad_lo(u32 *out, u32 a, u32 b) {
    u32 acc[ILP];
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a + i), "r"(b));
    }
    u32 s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(u32 *out, u32 a, u32 b) {
    u32 acc[ILP];
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; xor.b32 %0, t, %2; }" : "+r"(acc[i]) : "r"(a + i), "r"(b));
    }
    u32 s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix(u64 *out, u32 a, u32 b) {
    u64 acc[ILP];
    u32 x[ILP];
    for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x + i; x[i] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a + i), "r"(b));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
        }
    }
    u64 s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mulhi64(u64 *out, u64 a, u64 b) {
    u64 acc[ILP];
    for (int i = 0; i < ILP; ++i) acc[i] = a + threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = __umul64hi(acc[i], b) + a;
    }
    u64 s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
double run(F launch, int blocks, int threads, double ops_per_thread) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return 5.0 * blocks * threads * ops_per_thread / (ms * 1e-3);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    void *buf;
    cudaMalloc(&buf, (size_t)blocks * threads * 8);
    const double per = (double)ITERS * ILP;
    const double hz = clk_khz * 1e3;
    auto rep = [&](const char *name, double ops, double instr_per_op) {
        printf("%-28s %8.2f Tops/s  = %6.1f ops/clk/SM (at %d MHz nominal)  ~%.0f thread-instr/clk/SM\n", name,
               ops / 1e12, ops / hz / sms, clk_khz / 1000, ops / hz / sms * instr_per_op);
    };
    printf("%s, %d SMs\n", p.name, sms);
    rep("IMAD.WIDE.U32", run([&] { k_imad_wide<<<blocks, threads>>>((u64 *)buf, 3, 5); }, blocks, threads, per), 1);
    rep("IMAD (lo)", run([&] { k_imad_lo<<<blocks, threads>>>((u32 *)buf, 3, 5); }, blocks, threads, per), 1);
    rep("IADD3+LOP3 pair", run([&] { k_iadd3<<<blocks, threads>>>((u32 *)buf, 3, 5); }, blocks, threads, per), 2);
    rep("IMAD.WIDE + IADD3 mix", run([&] { k_mix<<<blocks, threads>>>((u64 *)buf, 3, 5); }, blocks, threads, per), 2);
    rep("mul.hi.u64 (+add)", run([&] { k_mulhi64<<<blocks, threads>>>((u64 *)buf, 12345, 0x9e3779b97f4a7c15ull); }, blocks, threads, per), 1);
    return 0;
}
