// bf_parts.cu -- where do the butterfly's issue cycles go?  Times its three parts separately
// (exact 64x64 mulhi | 6-IMAD lo chain | add/sub + csub) and together, 16 independent values/thread.
#include <cstdio>
#include <cuda_runtime.h>
#include "../aloha_b200/csrc/modarith.cuh"
using namespace alb;

template <int MODE>
__global__ void __launch_bounds__(256, 3) k(u64 *io, u64 q, u64 w, u64 wp, int iters) {
    u64 x[16];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < 16; ++i) x[i] = io[t + i * gridDim.x * blockDim.x];
    const u64 q2 = 2 * q, q8 = 8 * q, nq = 0 - q;
    u64 ww = w + threadIdx.x, wwp = wp + threadIdx.x;   // per-thread (register) twiddle, as in the kernels
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            u64 &a = x[e], &b = x[e + 8];
            if (MODE == 0) {                 // mulhi only
                b = __umul64hi(b, wwp) + a;
            } else if (MODE == 1) {          // lo chain only (qh := a)
                const u32 yl = (u32)b, yh = (u32)(b >> 32), wl = (u32)ww, wh = (u32)(ww >> 32);
                const u32 ql = (u32)a, qhh = (u32)(a >> 32), nl = (u32)nq, nh = (u32)(nq >> 32);
                u64 r = a;
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(yl), "r"(wl));
                asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(ql), "r"(nl));
                u32 hi = (u32)(r >> 32);
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(yl), "r"(wh));
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(yh), "r"(wl));
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(ql), "r"(nh));
                asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(qhh), "r"(nl));
                b = ((u64)hi << 32) | (u32)r;
            } else if (MODE == 2) {          // add/sub part + csub on both
                const u64 xp = a + b;
                b = csub_s((a + a + q2) - xp, q8);
                a = csub_s(xp, q8);
            } else {                          // whole butterfly + csub on both
                const u64 xp = shoup_mac(a, b, ww, wwp, nq);
                b = csub_s((a + a + q2) - xp, q8);
                a = csub_s(xp, q8);
            }
        }
    }
    for (int i = 0; i < 16; ++i) io[t + i * gridDim.x * blockDim.x] = x[i];
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 3, iters = 4000;
    u64 *io; cudaMalloc(&io, (size_t)blocks * 256 * 16 * 8); cudaMemset(io, 1, (size_t)blocks * 256 * 16 * 8);
    const u64 q = (1ull << 60) - (1ull << 17) * 7 + 1;
    const char *names[] = {"mulhi64 (4 IMAD.WIDE + carries)", "lo chain (2 WIDE + 4 IMAD)", "add/sub + 2 csub", "whole butterfly + 2 csub"};
    for (int m = 0; m < 4; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto launch = [&] {
            switch (m) {
            case 0: k<0><<<blocks, 256>>>(io, q, 12345, 67890, iters); break;
            case 1: k<1><<<blocks, 256>>>(io, q, 12345, 67890, iters); break;
            case 2: k<2><<<blocks, 256>>>(io, q, 12345, 67890, iters); break;
            default: k<3><<<blocks, 256>>>(io, q, 12345, 67890, iters); break;
            }
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double units = (double)blocks * 256 * 8 * iters;
        const double per_clk_sm = units / (ms * 1e-3) / (clk_khz * 1e3) / sms;
        printf("%-36s %.2f /clk/SM  = %.1f issue-cycles per warp-op per SMSP\n", names[m], per_clk_sm, 128.0 / per_clk_sm);
    }
    return 0;
}
