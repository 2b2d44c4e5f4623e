// alu_costs.cu -- issue cost (cycles per warp instruction per SM sub-partition) of the individual ALU
// instruction forms the butterfly's add / subtract / conditional-subtract code compiles to.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(u32 *io, int iters) {
    u32 a[ILP], b[ILP], c[ILP];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < ILP; ++i) { a[i] = io[t + i * 64]; b[i] = io[t + i * 64 + 7]; c[i] = io[t + i * 64 + 13]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                       // IADD3 2-reg
            if (MODE == 1) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));   // IADD3 3-reg (fused)
            if (MODE == 2) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(c[(i + 1) % ILP]));   // 64-bit add
            if (MODE == 3) asm volatile("{ .reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));          // ISETP + SEL
            if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));  // LOP3 3-reg
            if (MODE == 5) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));              // SHF
            if (MODE == 6) asm volatile("{ .reg .pred p; setp.ge.u32 p, %0, %1; @p sub.u32 %0, %0, %1; }" : "+r"(a[i]) : "r"(b[i]));                       // ISETP + predicated IADD
            if (MODE == 7) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                        // IMNMX
            if (MODE == 8) { float f = __uint_as_float(a[i]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__uint_as_float(b[i])), "f"(__uint_as_float(c[i]))); a[i] = __float_as_uint(f); }   // FFMA 3-reg
        }
    }
    u32 s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i] + b[i];
    io[t] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 4, iters = 4096;
    u32 *io; cudaMalloc(&io, 1 << 24); cudaMemset(io, 5, 1 << 24);
    const char *names[] = {"IADD3 (2 regs)", "IADD3 (3 regs)", "add.cc + addc (64-bit add)", "ISETP + SEL", "LOP3 (3 regs)", "SHF",
                           "ISETP + @p IADD", "IMNMX.U32", "FFMA (3 regs)"};
    const double instr[] = {1, 1, 2, 2, 1, 1, 2, 1, 1};
    for (int m = 0; m < 9; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto launch = [&] {
            switch (m) {
            case 0: k<0><<<blocks, 256>>>(io, iters); break; case 1: k<1><<<blocks, 256>>>(io, iters); break;
            case 2: k<2><<<blocks, 256>>>(io, iters); break; case 3: k<3><<<blocks, 256>>>(io, iters); break;
            case 4: k<4><<<blocks, 256>>>(io, iters); break; case 5: k<5><<<blocks, 256>>>(io, iters); break;
            case 6: k<6><<<blocks, 256>>>(io, iters); break; case 7: k<7><<<blocks, 256>>>(io, iters); break;
            default: k<8><<<blocks, 256>>>(io, iters); break;
            }
        };
        launch(); cudaDeviceSynchronize();
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double groups = (double)blocks * 256 * iters * ILP;
        const double per_clk_sm = groups / (ms * 1e-3) / (clk_khz * 1e3) / sms;    // groups per clk per SM
        printf("%-30s %.2f cycles per warp-group per SMSP  (%.2f per instruction, %g instr)\n", names[m], 128.0 / per_clk_sm,
               128.0 / per_clk_sm / instr[m], instr[m]);
    }
    return 0;
}
