// intpipe_bench2.cu -- issue cost of the exact instruction forms the butterfly uses (all-register
// operands): IMAD.WIDE with 64-bit addend, IMAD lo, 64-bit add (IADD3 + IADD3.X), and mixes.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int ITERS = 2048, ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(u64 *io, int iters) {
    u64 acc[ILP];
    u32 a[ILP], b[ILP];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < ILP; ++i) { acc[i] = io[t + i * 1024]; a[i] = (u32)io[t + i * 37 + 5]; b[i] = (u32)(io[t + i * 91 + 3] >> 7); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) {          // IMAD.WIDE.U32 Rd64, Ra, Rb, Rd64
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));
            } else if (MODE == 1) {   // IMAD lo: Rd, Ra, Rb, Rd
                u32 lo = (u32)acc[i];
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(a[i]), "r"(b[i]));
                acc[i] = (acc[i] & 0xffffffff00000000ull) | lo;
            } else if (MODE == 2) {   // 64-bit add of two register pairs
                asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"(acc[(i + 1) % ILP]));
            } else if (MODE == 3) {   // wide mad + 64-bit add alternating
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[(i + 3) % ILP]) : "l"(acc[(i + 5) % ILP]));
            } else if (MODE == 4) {   // mul.wide (no addend)
                u64 r;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a[i]), "r"((u32)acc[i]));
                acc[i] = r;
            } else if (MODE == 5) {   // 32-bit 2-operand add
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (MODE == 6) {   // wide mad + TWO 64-bit adds (1 IMAD : 4 ALU)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(a[i]), "r"(b[i]));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[(i + 3) % ILP]) : "l"(acc[(i + 5) % ILP]));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[(i + 2) % ILP]) : "l"(acc[(i + 6) % ILP]));
            }
        }
    }
    u64 s = 0;
    for (int i = 0; i < ILP; ++i) s += acc[i] + a[i];
    io[t] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    u64 *io; cudaMalloc(&io, 1 << 26); cudaMemset(io, 5, 1 << 26);
    const char *names[] = {"IMAD.WIDE (64b addend, all regs)", "IMAD lo (all regs)", "add.u64 (IADD3+IADD3.X)",
                           "IMAD.WIDE + add.u64 (1:2)", "mul.wide (no addend)", "add.u32", "IMAD.WIDE + 2x add.u64 (1:4)"};
    const double instr[] = {1, 1, 2, 3, 1, 1, 5};
    for (int warps = 8; warps <= 32; warps *= 2) {
        for (int m = 0; m < 7; ++m) {
            const int blocks = sms * warps / 8;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            auto launch = [&] {
                switch (m) {
                case 0: k<0><<<blocks, 256>>>(io, ITERS); break; case 1: k<1><<<blocks, 256>>>(io, ITERS); break;
                case 2: k<2><<<blocks, 256>>>(io, ITERS); break; case 3: k<3><<<blocks, 256>>>(io, ITERS); break;
                case 4: k<4><<<blocks, 256>>>(io, ITERS); break; case 5: k<5><<<blocks, 256>>>(io, ITERS); break;
                case 6: k<6><<<blocks, 256>>>(io, ITERS); break;
                }
            };
            launch(); cudaDeviceSynchronize();
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double groups = (double)blocks * 256 * ITERS * ILP;
            const double ipc = groups * instr[m] / (ms * 1e-3) / (clk_khz * 1e3) / sms;
            printf("%2d warps/SM  %-36s %6.1f thread-instr/clk/SM  (IPC/SMSP %.2f)\n", warps, names[m], ipc, ipc / 128);
        }
    }
    return 0;
}
