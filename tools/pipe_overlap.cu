// pipe_overlap.cu -- can the IMAD pipe and the ALU pipe of one SM sub-partition run concurrently when
// the two instruction streams come from DIFFERENT warps?  Half the warps run a multiply-only loop,
// the other half an add / conditional-subtract loop; compare with each half running alone.
#include <cstdio>
#include <cuda_runtime.h>
#include "../aloha_b200/csrc/modarith.cuh"
using namespace alb;

// mode bit0: mul warps active, bit1: alu warps active.  Warp w runs MUL if (w & 1) == 0 else ALU.
__global__ void __launch_bounds__(256, 3) k(u64 *io, u64 q, u64 wp, int iters, int mode) {
    u64 x[16];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < 16; ++i) x[i] = io[t + i * gridDim.x * blockDim.x];
    const u64 q2 = 2 * q, q8 = 8 * q;
    const u64 wwp = wp + threadIdx.x;
    const bool mulwarp = ((threadIdx.x >> 5) & 1) == 0;
    if (mulwarp) {
        if (!(mode & 1)) return;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                x[e + 8] = __umul64hi(x[e + 8], wwp) + x[e];
                x[e] = __umul64hi(x[e], wwp) + x[e + 8];
            }
        }
    } else {
        if (!(mode & 2)) return;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const u64 xp = x[e] + x[e + 8];
                x[e + 8] = csub_s((x[e] + x[e] + q2) - xp, q8);
                x[e] = csub_s(xp, q8);
            }
        }
    }
    for (int i = 0; i < 16; ++i) io[t + i * gridDim.x * blockDim.x] = x[i];
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 3, iters = 4000;
    u64 *io; cudaMalloc(&io, (size_t)blocks * 256 * 16 * 8); cudaMemset(io, 1, (size_t)blocks * 256 * 16 * 8);
    const u64 q = (1ull << 60) - (1ull << 17) * 7 + 1;
    const char *names[] = {"", "mul warps only", "alu warps only", "both (12 mul + 12 alu warps per SM)"};
    float t[4] = {0, 0, 0, 0};
    for (int m = 1; m <= 3; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<blocks, 256>>>(io, q, 67890, iters, m); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<blocks, 256>>>(io, q, 67890, iters, m); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&t[m], e0, e1);
        printf("%-40s %.3f ms\n", names[m], t[m]);
    }
    printf("sum of the two alone = %.3f ms; together = %.3f ms  => overlap factor %.2f (1.0 = none, %.2f = perfect)\n",
           t[1] + t[2], t[3], (t[1] + t[2]) / t[3], (t[1] + t[2]) / (t[1] > t[2] ? t[1] : t[2]));
    return 0;
}
