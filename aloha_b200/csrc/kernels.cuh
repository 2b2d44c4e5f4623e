// kernels.cuh -- host-visible launch interface of the sm_100a kernels (kernels.cu).
// All launchers are asynchronous on the given stream and take device-resident job tables, so a
// single launch serves a whole batch of independent limb-polynomials (different moduli, different
// SPM / vreg / KSK addresses).
#pragma once
#include <cuda.h>            // CUtensorMap (type only; the encoder is fetched from the driver at run time)
#include <cuda_runtime.h>
#include <cstdint>

#include "aut_plan.hpp"

namespace alb {

typedef unsigned long long u64;
typedef unsigned int u32;

// Twiddle entry {w, wp}.  Forward table index j holds w = psi^bitrev(j, logN) (reference order:
// sim/vp/tf_rom_generator/tf_rom_generator.sv:28-30,111), inverse table the same with psi^-1
// (:61-63,147-148).  The companion word depends on the modulus form (below):
//   FORM_GENERIC : wp = floor(w * 2^64 / q)   (Shoup quotient)
//   FORM_PM      : wp = w * 2^32 mod q        (second half of the split product, modarith.cuh mul_pm)
struct __align__(16) Tw { u64 w, wp; };

// How the transform kernels multiply modulo q.  Chosen per modulus when its tables are built; one launch
// only ever holds jobs of one form.
enum ModForm : u32 {
    FORM_GENERIC = 0,      // any 60-bit prime: Shoup / Harvey, 10 IMAD per product
    FORM_PM = 1            // q = 2^60 - d with d <= 2^27 (pseudo-Mersenne): split product + one fold, 5 IMAD
};

struct ModulusConsts {
    u64 q;
    u64 ninv, ninv_p;      // N^-1 mod q as a twiddle pair         -- inverse transform, last stage
    u64 wninv, wninv_p;    // itw[1] * N^-1 mod q as a twiddle pair
    u32 mest;              // floor(2^91 / q)                      (FORM_GENERIC)
    u32 pre;               // element-wise op folded into the transform's load: 0 none, 1 VCPY, 2 VFQMOD
    u32 form;              // ModForm
    u32 d;                 // 2^60 - q                             (FORM_PM)
    u64 q3;                // 3q, the butterfly's subtraction offset (FORM_PM).  Loaded, not computed: ptxas
                           // otherwise re-derives it from q inside every butterfly (IMAD.WIDE by 3 + negation)
};
enum NttPre : u32 { PRE_NONE = 0, PRE_VCPY = 1, PRE_VFQMOD = 2 };

// One limb-polynomial transform.
struct NttJob {
    const u64 *src;
    u64 *dst;
    const Tw *tw;          // forward or inverse table for (modulus, N)
    const Tw *rtw;         // the same twiddles in the row kernels' read order, 256 entries per row of the
                           // N/256 x 256 view: what one TMA copy stages for a tile (ntt_kernels.cu, row_slot)
    ModulusConsts mc;
};
// Slot of twiddle (level u, index j < 2^u) inside a row's 256-entry block of `rtw` (slot 0 unused), as the
// INVERSE row pass reads it (half-warp per row, 16 coefficients per thread; GS level lt = 7 - u).
// Levels 0..3 are warp-uniform reads; levels 4..7 are read by lane h = j >> (u-4) with m = j mod 2^(u-4)
// and are laid out [m][h] so that the 16 lanes of a half-warp read 16 consecutive entries.
__host__ __device__ constexpr u32 row_slot(u32 u, u32 j) {
    return u < 4 ? (1u << u) + j : (1u << u) + (j & ((1u << (u - 4)) - 1)) * 16 + (j >> (u - 4));
}

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

// VAUT through the tile decomposition of aut_plan.hpp (the default path; PermJob + vaut_kernel is the
// 8-byte gather kept for A/B measurements, ALOHA_F_AUT_GATHER).
struct AutJob {
    u64 *dst;
    const u64 *src;
    u64 q;
    u64 k;                 // Galois element as the RTL sees it: the sign comes from (i*k) mod 2n
    AutPlan plan;
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

// Fast basis extension (the key-switch's ModUp / ModDown inner loop, hks.py): per source limb a VCPY or VFQMOD
// into the target modulus, a VFQMUL.vs by the limb's constant and a VFQADD into the running sum, optionally a
// final VFQSUB.vs -- one pass over the operands instead of three per term.
//   dst = [ (((m_0 + m_1) + m_2) + ...) - post_s ],   m_t = barrett(r(pre_t(x_t)), s_t)
struct BextTerm {
    const u64 *x;
    u64 s;                 // scalar (already reduced once, as modalu.sv:46 does)
    u64 sp;                // floor(s * 2^64 / q): the scalar's Shoup companion (fast path)
    u64 pre;               // NttPre: 0 none, 1 VCPY, 2 VFQMOD
};
struct BextJob {
    u64 *dst;
    const BextTerm *terms;
    u64 q, iq;
    u64 post_s;            // VFQSUB.vs scalar (reduced once)
    u32 nterms, post;      // post: 0 none, 1 subtract post_s
    u32 fast, mest;        // fast: q is a 60-bit modulus, iq its Barrett constant and every scalar is below q, so the
                           // RTL chain equals exact modular arithmetic on every in-domain input and the kernel may
                           // use lazy Shoup products (out-of-domain words take the RTL chain element by element);
                           // mest = floor(2^91 / q)
};

// dst = c + aut_k(x) * p   (rotate-and-sum inner step: VAUT, VFQMUL.vv, VFQADD.vv fused)
struct AutMacJob {
    u64 *dst;
    const u64 *c;
    const u64 *x;
    const u64 *p;
    u64 q, iq, k, kinv;
    AutPlan plan;          // tile decomposition of (n, k) (aut_plan.hpp); unused by the gather variant
};

// The forward row pass's layout of the same 256-entry block (a whole warp per row, 8 coefficients per
// thread, phases of 3 + 3 + 2 levels): levels 0..2 warp-uniform; 3..5 read by lane group hi = lane / 4,
// laid out [m][hi]; 6..7 read by lane T, laid out [m][T].
__host__ __device__ constexpr u32 row_slot8(u32 u, u32 j) {
    return u < 4 ? (1u << u) + j
         : u == 4 ? 16 + (j & 1) * 8 + (j >> 1)
         : u == 5 ? 32 + (j & 3) * 8 + (j >> 2)
         : u == 6 ? 64 + (j & 1) * 32 + (j >> 1)
                  : 128 + (j & 3) * 32 + (j >> 2);
}

// 16 jobs of one direction that share a modulus (and load op), as the TMA-staged row passes want them:
// one record is one bulk copy.  `src` is what the row pass reads -- forward: the column pass's output
// (= dst), or the job's source when N = 256 and there is no column pass; inverse: the job's source, which
// the inverse pass fetches through a tensor map (`src_map` = which device buffer, `src_line` = offset of
// the polynomial inside it in 128-byte lines).
struct __align__(16) NttRowGroup {
    const u64 *src[16];
    u64 *dst[16];
    const Tw *rtw;
    u64 pad;
    ModulusConsts mc;
    u32 src_line[16];
    u32 src_map[16];
};

// Tensor maps over the device buffers a transform can read (SPM image, KSK image, renaming pool), each seen
// as [lines][16 words] with SWIZZLE_128B: a 16 x 16 box is one 2 KiB row of a polynomial, landing in shared
// memory with 16-byte chunk c of line l at chunk c ^ (l mod 8) -- the layout in which a thread can read
// its 16 contiguous coefficients without bank conflicts.
struct TmaMaps {
    CUtensorMap m[3];
};

// every job of one launch has mc.form == form.  The first 16 * ngroups jobs are also described by
// `groups_dev` (runs of 16 jobs sharing a modulus): their row pass runs as the TMA-staged persistent
// kernel, one tile = the same row of 16 polynomials, twiddles staged once per tile.
cudaError_t launch_ntt_forward(const NttJob *jobs_dev, u32 njobs, const NttRowGroup *groups_dev, u32 ngroups, u32 logn,
                               u32 form, cudaStream_t st);
cudaError_t launch_ntt_inverse(const NttJob *jobs_dev, u32 njobs, const NttRowGroup *groups_dev, u32 ngroups,
                               const TmaMaps *maps_host, u32 logn, u32 form, cudaStream_t st);
cudaError_t launch_ew(u32 alu_op, const EwJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_vaut(const PermJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);          // 8-byte gather
// tiled: max_tiles = the largest plan.ntiles among the launch's jobs
cudaError_t launch_vaut_tiled(const AutJob *jobs_dev, u32 njobs, u32 n, u32 max_tiles, cudaStream_t st);
cudaError_t launch_autmac_tiled(const AutMacJob *jobs_dev, u32 njobs, u32 n, u32 max_tiles, cudaStream_t st);
cudaError_t launch_vroli(const PermJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_pease(const PeaseJob *jobs_dev, u32 njobs, u32 logn, u32 stage, bool inverse, cudaStream_t st);
cudaError_t launch_copy(const CopyJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_mac(const MacJob *jobs_dev, u32 njobs, u32 terms, u32 n, cudaStream_t st);
cudaError_t launch_autmac(const AutMacJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_sop(const SopJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_muladd(const MulAddJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);
cudaError_t launch_bext(const BextJob *jobs_dev, u32 njobs, u32 n, cudaStream_t st);

// number of kernel launches issued by the launchers above since process start (bench accounting)
unsigned long long kernel_launch_count();

}  // namespace alb
