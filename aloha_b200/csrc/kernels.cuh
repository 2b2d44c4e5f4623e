// kernels.cuh -- host-visible launch interface of the sm_100a kernels (kernels.cu).
// All launchers are asynchronous on the given stream and take device-resident job tables, so a
// single launch serves a whole batch of independent limb-polynomials (different moduli, different
// SPM / vreg / KSK addresses).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace alb {

typedef unsigned long long u64;
typedef unsigned int u32;

// Twiddle entry: {w, floor(w * 2^64 / q)}.  Forward table index j holds psi^bitrev(j, logN)
// (reference order: sim/vp/tf_rom_generator/tf_rom_generator.sv:28-30,111), inverse table the same
// with psi^-1 (:61-63,147-148).
struct __align__(16) Tw { u64 w, wp; };

struct ModulusConsts {
    u64 q;
    u64 ninv, ninv_p;      // N^-1 mod q (Shoup pair)            -- inverse transform, last stage
    u64 wninv, wninv_p;    // itw[1] * N^-1 mod q (Shoup pair)
    u32 mest;              // floor(2^91 / q)
    u32 pre;               // element-wise op folded into the transform's load: 0 none, 1 VCPY, 2 VFQMOD
};
enum NttPre : u32 { PRE_NONE = 0, PRE_VCPY = 1, PRE_VFQMOD = 2 };

// One limb-polynomial transform.
struct NttJob {
    const u64 *src;
    u64 *dst;
    const Tw *tw;          // forward or inverse table for (modulus, N)
    ModulusConsts mc;
};

// Element-wise modalu op over n words.
struct EwJob {
    u64 *dst;
    const u64 *a;
    const u64 *b;          // second vector operand (vv forms) or nullptr
    u64 s;                 // scalar (already reduced once, as modalu.sv:46 does)
    u64 q, iq;
};

// dst[(i*k) mod n] = ((i*k) mod 2n >= n) ? q - src[i] : src[i]   (VAUT), or
// dst[(i - rot) mod n] = src[i]                                   (VROLI)
struct PermJob {
    u64 *dst;
    const u64 *src;
    u64 q;
    u64 k;                 // VAUT: Galois element as the RTL sees it (already truncated); VROLI: unused
    u64 kinv;              // VAUT: k^-1 mod n (gather form); VROLI: rot mod n
};

// One stage of the reference's constant-geometry schedule, word-exact (ALOHA_F_STRICT):
//   forward  : dst[2p] = add(r(x[p]), m), dst[2p+1] = sub(r(x[p]), m), m = barrett(r(x[p+N/2]), w), w = tw[2^s + (p mod 2^s)]
//   inverse  : dst[p] = half(add(x[2p], x[2p+1])), dst[p+N/2] = half(barrett(sub(x[2p], x[2p+1]), w)), w = itw[2^f + (p mod 2^f)], f = logN-1-s
struct PeaseJob {
    u64 *dst;
    const u64 *src;
    const Tw *tw;
    u64 q, iq;
};

struct CopyJob {
    u64 *dst;
    const u64 *src;
};

// Fused multiply-accumulate chain: dst = sum_{t<terms} a[t] * b[t]   (each product an exact RTL
// Barrett, sums RTL addmod in instruction order).  terms <= 4.
struct MacJob {
    u64 *dst;
    const u64 *a[4];
    const u64 *b[4];
    u64 q, iq;
};

// dst = c + a * b   (VFQMUL.vv followed by VFQADD.vv, fused)
struct MulAddJob {
    u64 *dst;
    const u64 *c, *a, *b;
    u64 q, iq;
};

// dst = ((a_0 b_0 + a_1 b_1) + a_2 b_2) + ...   -- a whole VFQMUL / VFQADD accumulation chain (the
// key-switch inner product over digits).  pairs[2t], pairs[2t+1] = a_t, b_t.
struct SopJob {
    u64 *dst;
    const u64 *const *pairs;
    u64 q, iq;
    u32 terms, pad;
};

// dst = c + aut_k(x) * p   (rotate-and-sum inner step: VAUT, VFQMUL.vv, VFQADD.vv fused)
struct AutMacJob {
    u64 *dst;
    const u64 *c;
    const u64 *x;
    const u64 *p;
    u64 q, iq, k, kinv;
};

cudaError_t launch_ntt_forward(const NttJob *jobs_dev, u32 njobs, u32 logn, cudaStream_t st);
cudaError_t launch_ntt_inverse(const NttJob *jobs_dev, u32 njobs, u32 logn, cudaStream_t st);
cudaError_t launch_ew(u32 alu_op, const EwJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_vaut(const PermJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_vroli(const PermJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_pease(const PeaseJob *jobs_dev, u32 njobs, u32 logn, u32 stage, bool inverse, cudaStream_t st);
cudaError_t launch_copy(const CopyJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_mac(const MacJob *jobs_dev, u32 njobs, u32 terms, u32 n, cudaStream_t st);
cudaError_t launch_autmac(const AutMacJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_sop(const SopJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_muladd(const MulAddJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);

// number of kernel launches issued by the launchers above since process start (bench accounting)
unsigned long long kernel_launch_count();

}  // namespace alb
