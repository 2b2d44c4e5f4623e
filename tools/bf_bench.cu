// bf_bench.cu -- arithmetic-only ceiling of the NTT butterfly code shape: E coefficients per thread
// in registers, radix-E blocks (log2 E stages) repeated, twiddles read from shared memory per group
// (as the kernels read them from L1/L2), no global traffic in the loop.  Reports butterflies per
// clock per SM; the NTT kernels cannot exceed this, so (kernel rate / this rate) separates arithmetic
// cost from memory / exchange cost.
#include <cstdio>
#include <cuda_runtime.h>
#include "../aloha_b200/csrc/modarith.cuh"
using namespace alb;
struct __align__(16) Tw { u64 w, wp; };

__device__ __forceinline__ void ct_bf(u64 &x, u64 &y, const Tw &t, u64 nq, u64 q2) {
    const u64 xp = shoup_mac(x, y, t.w, t.wp, nq);
    y = (x + x + q2) - xp;
    x = xp;
}

// pseudo-Mersenne butterfly (q = 2^60 - d): 5 IMAD.WIDE, bounds +3q per stage
__device__ __forceinline__ void ct_bf_pm(u64 &x, u64 &y, const Tw &t, u32 d2, u64 q3) {
    const u64 m = mul_pm(y, t.w, t.wp, d2);
    y = x + q3 - m;
    x = x + m;
}

__device__ __forceinline__ void ct_bf_pm2(u64 &x, u64 &y, const Tw &t, u32 d2, u64 q3) {
    u64 P, L;
    mul_pm_parts(y, t.w, t.wp, d2, P, L);
    y = (x + q3 - P) - L;
    x = x + P + L;
}
__device__ __forceinline__ void ct_bf_pm3(u64 &x, u64 &y, const Tw &t, u32 d2, u64 q3) {
    u64 P, L;
    mul_pm_parts(y, t.w, t.wp, d2, P, L);
    const u64 xn = x + P + L;
    y = (x + x + q3) - xn;
    x = xn;
}

template <int LOGE, int MINB, int PM = 0>
__global__ void __launch_bounds__(256, MINB) k(u64 *io, const Tw *tw, u64 q, u64 q3, int iters) {
    constexpr int E = 1 << LOGE;
    __shared__ Tw stw[256];
    u64 x[E];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    stw[threadIdx.x] = tw[threadIdx.x];
    __syncthreads();
    for (int i = 0; i < E; ++i) x[i] = io[t + i * gridDim.x * blockDim.x];
    const u64 q2 = 2 * q, q8 = 8 * q, nq = 0 - q;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int v = 0; v < LOGE; ++v) {
            const int half = E >> (v + 1);
            Tw w;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                if (e & half) continue;
                if ((e & (half - 1)) == 0) w = stw[(((threadIdx.x >> 5) + it) & 15) * 16 + (1 << v) + (e >> (LOGE - v))];   // warp-uniform: a broadcast, no bank conflicts
                if (PM == 2) ct_bf_pm2(x[e], x[e + half], w, (u32)(2 * nq), q3);
                else if (PM == 3) ct_bf_pm3(x[e], x[e + half], w, (u32)(2 * nq), q3);
                else if (PM) ct_bf_pm(x[e], x[e + half], w, (u32)(2 * nq), q3);   // nq = 2^64 - q = 16 d ... only timing matters here
                else ct_bf(x[e], x[e + half], w, nq, q2);
            }
        }
        if (PM) {
#pragma unroll
            for (int i = 0; i < E; i += 2) x[i] = fold_pm(x[i], (u32)(nq >> 4));     // upper inputs only, as the kernels do
        } else {
#pragma unroll
            for (int i = 0; i < E; ++i) x[i] = csub_s(x[i], q8);
        }
    }
    for (int i = 0; i < E; ++i) io[t + i * gridDim.x * blockDim.x] = x[i];
}

template <int LOGE, int MINB, int PM = 0>
void run(int sms, int clk_khz, u64 *io, Tw *tw, u64 q) {
    const int iters = 2000, E = 1 << LOGE;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<LOGE, MINB, PM>);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<LOGE, MINB, PM>, 256, 0);
    const int blocks = sms * occ;
    k<LOGE, MINB, PM><<<blocks, 256>>>(io, tw, q, 3 * q, iters);
    cudaDeviceSynchronize();
    cudaEventRecord(e0); k<LOGE, MINB, PM><<<blocks, 256>>>(io, tw, q, 3 * q, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bf = (double)blocks * 256 * (E / 2 * LOGE) * iters;
    printf("%s E=%2d minb=%d regs=%3d occ=%d CTA/SM (%2d warps): %.2f bf/clk/SM -> %.2f M limb-NTT/s ceiling\n", PM == 1 ? "pseudo-Mersenne" : PM == 2 ? "PM 3-input adds" : PM == 3 ? "PM x'=x+P+L,2x-" : "Shoup          ", E, MINB,
           fa.numRegs, occ, occ * 8, bf / (ms * 1e-3) / (clk_khz * 1e3) / sms, bf / (ms * 1e-3) / 524288 / 1e6);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    u64 *io; Tw *tw;
    cudaMalloc(&io, (size_t)sms * 8 * 256 * 16 * 8);
    cudaMalloc(&tw, 4096 * sizeof(Tw));
    cudaMemset(io, 1, (size_t)sms * 8 * 256 * 16 * 8);
    cudaMemset(tw, 3, 4096 * sizeof(Tw));
    const u64 q = (1ull << 60) - (1ull << 17) * 7 + 1;
    run<4, 1>(sms, clk_khz, io, tw, q);
    run<4, 2>(sms, clk_khz, io, tw, q);
    run<4, 3>(sms, clk_khz, io, tw, q);
    run<4, 4>(sms, clk_khz, io, tw, q);
    run<4, 1, 1>(sms, clk_khz, io, tw, q);
    run<4, 2, 1>(sms, clk_khz, io, tw, q);
    run<4, 3, 1>(sms, clk_khz, io, tw, q);
    run<4, 2, 2>(sms, clk_khz, io, tw, q);
    run<4, 3, 2>(sms, clk_khz, io, tw, q);
    run<4, 2, 3>(sms, clk_khz, io, tw, q);
    run<4, 3, 3>(sms, clk_khz, io, tw, q);
    run<4, 4, 1>(sms, clk_khz, io, tw, q);
    run<3, 4, 1>(sms, clk_khz, io, tw, q);
    run<3, 2>(sms, clk_khz, io, tw, q);
    run<3, 4>(sms, clk_khz, io, tw, q);
    run<3, 6>(sms, clk_khz, io, tw, q);
    run<2, 4>(sms, clk_khz, io, tw, q);
    run<2, 8>(sms, clk_khz, io, tw, q);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
