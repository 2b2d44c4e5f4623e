// bf_bench.cu -- arithmetic-only ceiling of the NTT butterfly code shape: 16 coefficients per thread
// in registers, radix-16 blocks (4 stages x 8 butterflies) repeated, no global/shared traffic in the
// loop.  Reports butterflies per clock per SM for several occupancies; the NTT kernels cannot exceed
// this, so (kernel rate / this rate) separates arithmetic cost from memory/exchange cost.
#include <cstdio>
#include <cuda_runtime.h>
#include "../aloha_b200/csrc/modarith.cuh"
using namespace alb;
struct Tw { u64 w, wp; };

__device__ __forceinline__ void ct_bf(u64 &x, u64 &y, const Tw &t, u64 nq, u64 q2) {
    const u64 xp = shoup_mac(x, y, t.w, t.wp, nq);
    y = (x + x + q2) - xp;
    x = xp;
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) k(u64 *io, const Tw *tw, u64 q, int iters) {
    u64 x[16];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < 16; ++i) x[i] = io[t + i * gridDim.x * blockDim.x];
    const u64 q2 = 2 * q, q8 = 8 * q, nq = 0 - q;
    Tw w[15];
    for (int i = 0; i < 15; ++i) w[i] = tw[(threadIdx.x & 15) * 16 + i];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int half = 8 >> v;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (e & half) continue;
                ct_bf(x[e], x[e + half], w[(1 << v) - 1 + (e >> (4 - v))], nq, q2);
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = csub_s(x[i], q8);
    }
    for (int i = 0; i < 16; ++i) io[t + i * gridDim.x * blockDim.x] = x[i];
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, iters = 2000;
    u64 *io; Tw *tw;
    cudaMalloc(&io, (size_t)sms * 8 * 256 * 16 * 8);
    cudaMalloc(&tw, 4096 * sizeof(Tw));
    cudaMemset(io, 1, (size_t)sms * 8 * 256 * 16 * 8);
    cudaMemset(tw, 3, 4096 * sizeof(Tw));
    const u64 q = (1ull << 60) - (1ull << 17) * 7 + 1;
    for (int ctas = 1; ctas <= 4; ++ctas) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto launch = [&] {
            if (ctas <= 2) k<2><<<sms * ctas, 256>>>(io, tw, q, iters);
            else if (ctas == 3) k<3><<<sms * ctas, 256>>>(io, tw, q, iters);
            else k<4><<<sms * ctas, 256>>>(io, tw, q, iters);
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bf = (double)sms * ctas * 256 * 32.0 * iters;
        printf("%d CTA/SM (%2d warps): %.2f butterflies/clk/SM  (%.2f Gbf/s; one N=2^16 limb-NTT = 524288 bf -> %.2f M limb-NTT/s ceiling)\n",
               ctas, ctas * 8, bf / (ms * 1e-3) / (clk_khz * 1e3) / sms, bf / ms / 1e6, bf / (ms * 1e-3) / 524288 / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
