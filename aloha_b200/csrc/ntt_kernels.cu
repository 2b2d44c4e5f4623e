// ntt_kernels.cu -- batched negacyclic NTT / INTT for sm_100a, N = 2^8 .. 2^16.
//
// Replaces the reference's VNTT / VINTT datapath (src/vp/ntt/ntt_fsm.sv:49-81 address schedule,
// src/vp/vxu/modalu.sv:160-165,296-327 CT / GS butterflies, per-lane twiddle ROMs).  The RTL's
// constant-geometry schedule is a hardware artefact; its net effect is the in-place Harvey
// transform out[k] = a(psi^(2*bitrev(k)+1)) with canonical outputs, which
// is what these kernels compute -- bit-identical for every input word below 2q.
//
// Decomposition (4-step, no transposes): N = R x 256, element j = r*256 + c.
//   forward : column kernel  -- the first S1 = log2 R stages pair rows of one column; every
//             column uses the same R-1 twiddles tw[1..R-1];
//             row kernel     -- the last 8 stages stay inside one 2 KiB row; row r uses the 255
//             twiddles tw[2^u (R + r) + g], u = 0..7.
//   inverse : the mirror image (row kernel first, then column kernel, N^-1 folded into the last
//             stage's twiddles).
// Each thread keeps 16 coefficients in registers and runs radix-16 (4 stages) between exchanges
// through shared memory; butterflies are Harvey-lazy with Shoup twiddles on the IMAD pipe, values
// live in [0, 16q) (q < 2^60) and are reduced to canonical form once, at the very end.
// Inputs are taken as they are (no pre-reduction): the RTL conditionally subtracts q once, which
// maps [q, 2q) onto the same residue class, and the final canonical reduction makes the stored word
// identical for every input below 2q.  Forward bounds: +2q per stage from 2q; only the UPPER input of a
// butterfly is ever reduced (8q conditional subtract), and only when the stage would pass 16q.
//
// Every kernel exists in two arithmetic forms (kernels.cuh ModForm), selected per modulus on the host:
// FORM_GENERIC is the Shoup / Harvey arithmetic above for any 60-bit prime; FORM_PM serves q = 2^60 - d
// (d <= 2^27), where a product costs 5 IMAD.WIDE instead of 10 IMAD (modarith.cuh mul_pm): products are
// below 3q, bounds grow +3q per stage, and the reduction of an upper input is a fold at 2^60 (3
// instructions, result below 2q) instead of a conditional subtract.  Same transform, same canonical
// output words.
//
// The row passes exist twice: plain kernels (ntt_fwd_rows / ntt_inv_rows: rows loaded straight into
// registers, twiddles through L1; any batch) and persistent TMA-staged ones (ntt_fwd_rows_tma /
// ntt_inv_rows_tma) for batches that hold runs of 16 polynomials under one modulus -- rows, the row's
// twiddle block and the group record are staged in shared memory by cp.async.bulk / cp.async.bulk.tensor
// behind mbarriers, one tile ahead of the butterflies.
#include <atomic>
#include <cstdlib>

#include "kernels.cuh"
#include "modarith.cuh"

#ifndef ROWS_MINB
#define ROWS_MINB 2
#endif
#ifndef COLS_MINB
#define COLS_MINB 3
#endif
#ifndef ROWS_TMA_MINB
#define ROWS_TMA_MINB 4      /* staged inverse row pass: CTAs of 16 * ROWS_TMA_TILE threads per SM */
#endif
#ifndef ROWS_TMA_STAGES
#define ROWS_TMA_STAGES 2
#endif
#ifndef ROWS_TMA_TILE
#define ROWS_TMA_TILE 8      /* staged inverse row pass: rows (= half-warps) per tile and CTA, 8 or 16 */
#endif

namespace alb {

namespace {

__device__ __forceinline__ Tw ldtw(const Tw *p) {
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(p));
    Tw t;
    t.w = v.x;
    t.wp = v.y;
    return t;
}

template <int FORM> struct Arith;

// Any 60-bit prime.  CT butterfly, lazy: (x, y) -> (x + w y, x - w y + 2q); x seeds the multiply-add
// chain, so x' costs no add; y' = 2x + 2q - x'.
template <> struct Arith<FORM_GENERIC> {
    static constexpr int MULB = 2;    // a lazy product is below MULB q
    static constexpr int GROW = 2;    // forward stage: both outputs below (B_x + GROW) q
    static constexpr int REDB = 8;    // red(): [0, 16q) -> [0, REDB q)
    u64 q, q2, q4, q8, nq;
    u32 mest;
    __device__ __forceinline__ explicit Arith(const ModulusConsts &mc)
        : q(mc.q), q2(2 * mc.q), q4(4 * mc.q), q8(8 * mc.q), nq(0 - mc.q), mest(mc.mest) {}
    __device__ __forceinline__ u64 mul(u64 y, u64 w, u64 wp) const { return mul_shoup(y, w, wp, nq); }
    __device__ __forceinline__ void ct(u64 &x, u64 &y, const Tw &t) const {
        const u64 xp = shoup_mac(x, y, t.w, t.wp, nq);
        y = (x + x + q2) - xp;
        x = xp;
    }
    __device__ __forceinline__ u64 red(u64 x) const { return csub_s(x, q8); }
    __device__ __forceinline__ u64 norm(u64 x, int b) const {       // below b q (b <= 16) -> below 2q
        if (b > 8) x = csub_s(x, q8);
        if (b > 4) x = csub_s(x, q4);
        if (b > 2) x = csub_s(x, q2);
        return x;
    }
    __device__ __forceinline__ u64 canon(u64 x) const { return reduce_full(x, q, nq, mest); }   // any word
    __device__ __forceinline__ u64 canon_lazy(u64 x) const { return reduce_lazy(x, nq, mest); } // any word -> below 2q
    __device__ __forceinline__ u64 canon_mul(u64 x) const { return csub_s(x, q); }              // a product
    __device__ __forceinline__ u64 off(int b) const { return b <= 2 ? q2 : b <= 4 ? q4 : q8; }
};

// q = 2^60 - d.  (x, y) -> (x + t, x - t + 3q) with t = w y mod q below 3q.
template <> struct Arith<FORM_PM> {
    static constexpr int MULB = 3, GROW = 3, REDB = 2;
    u64 q, q2, q3, q4, q8;
    u32 d, d2;
    __device__ __forceinline__ explicit Arith(const ModulusConsts &mc)
        : q(mc.q), q2(2 * mc.q), q3(mc.q3), q4(4 * mc.q), q8(8 * mc.q), d(mc.d), d2(2 * mc.d) {}
    __device__ __forceinline__ u64 mul(u64 y, u64 w, u64 wp) const { return mul_pm(y, w, wp, d2); }
    __device__ __forceinline__ void ct(u64 &x, u64 &y, const Tw &t) const {
        u64 P, L;
        mul_pm_parts(y, t.w, t.wp, d2, P, L);
        y = (x + q3 - P) - L;
        x = x + P + L;
    }
    __device__ __forceinline__ u64 red(u64 x) const { return fold_pm(x, d); }
    __device__ __forceinline__ u64 norm(u64 x, int b) const { return b > 2 ? fold_pm(x, d) : x; }
    __device__ __forceinline__ u64 canon(u64 x) const { return canon_pm(x, q, d); }
    __device__ __forceinline__ u64 canon_lazy(u64 x) const { return fold_pm(x, d); }
    __device__ __forceinline__ u64 canon_mul(u64 x) const { return canon_pm(x, q, d); }
    __device__ __forceinline__ u64 off(int b) const { return b <= 2 ? q2 : b <= 4 ? q4 : q8; }
};

// Base-extension primitive folded into the load of a forward transform, word-exact:
//   VCPY   = addmod(r(x), 0)  -> two compare-based conditional subtracts (expander.v:396-417);
//   VFQMOD = barrett(r(x), 1) -> x mod q for EVERY 64-bit x (the RTL's quotient estimate is within one
//            of floor(x/q) for 60-bit q), i.e. exactly what canon() computes.
template <int FORM>
__device__ __forceinline__ u64 apply_pre(u64 x, u32 pre, const Arith<FORM> &A) {
    if (pre == PRE_VCPY) {
        x = x >= A.q ? x - A.q : x;
        x = x >= A.q ? x - A.q : x;
    } else if (pre == PRE_VFQMOD) {
        x = A.canon(x);
    }
    return x;
}

// Bound (units of q) of what the forward column pass stores for S1 column stages: start at 2, one GROW
// per stage, an upper-input reduction to REDB whenever the next stage would pass 16.
template <int FORM>
__host__ __device__ constexpr int cols_out_bound(int s1) {
    int b = 2;
    for (int s = 0; s < s1; ++s) b = (b + Arith<FORM>::GROW > 16 ? Arith<FORM>::REDB : b) + Arith<FORM>::GROW;
    return b;
}

}  // namespace

// ============================================================================ forward: columns
// S1 = number of column stages (R = 2^S1 rows).  LA = min(S1,4) stages in phase A on rows
// r = h + H k (H = R / 2^LA), LB = S1 - LA stages in phase B on rows 16 G + e.
//
// Bound bookkeeping (units of q, resolved at compile time after unrolling): every value entering a
// stage is < B q; both outputs are < (B_x + 2) q where B_x bounds the UPPER input only -- the lower
// input goes through the Shoup multiply, which accepts any 64-bit word.  So only the upper inputs are
// ever reduced (A.red: to REDB q), and only when B + GROW would pass 16.
#define ALOHA_LDTW ldtw
#define ALOHA_CT_STAGE(NELEM, HALF, TWIDX)                                           \
    {                                                                                \
        const bool red_ = B + AR::GROW > 16;                                         \
        Tw w_;                                                                       \
        _Pragma("unroll") for (int e = 0; e < (NELEM); ++e) {                        \
            if (e & (HALF)) continue;                                                \
            if ((e & ((HALF)-1)) == 0) {                                             \
                const int g0 = e & ~(2 * (HALF)-1);                                  \
                w_ = ALOHA_LDTW(tw + (TWIDX));                                       \
            }                                                                        \
            if (red_) x[e] = A.red(x[e]);                                            \
            A.ct(x[e], x[e + (HALF)], w_);                                           \
        }                                                                            \
        B = (red_ ? AR::REDB : B) + AR::GROW;                                        \
    }

template <int S1, int FORM>
__global__ void __launch_bounds__(256, COLS_MINB) ntt_fwd_cols(const NttJob *__restrict__ jobs) {
    typedef Arith<FORM> AR;
    constexpr int LA = S1 < 4 ? S1 : 4, LB = S1 - LA, E = 1 << LA, R = 1 << S1;
    constexpr int H = R / E;             // threads per column in phase A
    constexpr int W = 256 / H;           // tile width in columns
    constexpr int TILES = H;             // tiles per polynomial
    extern __shared__ u64 smem[];        // [R][W] when LB > 0

    const NttJob &job = jobs[blockIdx.x / TILES];
    const int c0 = (blockIdx.x % TILES) * W;
    const int t = threadIdx.x, c = t % W, hg = t / W;
    const AR A(job.mc);
    const u64 *src = job.src + c0 + c;
    u64 *dst = job.dst + c0 + c;
    const Tw *tw = job.tw;

    u64 x[E];
#pragma unroll
    for (int k = 0; k < E; ++k) x[k] = src[(size_t)(hg + H * k) * 256];   // < 2q (see header)
    if (job.mc.pre) {
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = apply_pre(x[k], job.mc.pre, A);
    }
    int B = 2;
    // phase A: stage v pairs k-bit (LA-1-v); idx = 2^v + (k >> (LA - v))
#pragma unroll
    for (int v = 0; v < LA; ++v) ALOHA_CT_STAGE(E, E >> (v + 1), (1 << v) + (g0 >> (LA - v)))
    if constexpr (LB == 0) {
#pragma unroll
        for (int k = 0; k < E; ++k) dst[(size_t)(hg + H * k) * 256] = x[k];
    } else {
        // exchange: rows h + H k  ->  rows 16 G + e
#pragma unroll
        for (int k = 0; k < E; ++k) smem[(hg + H * k) * W + c] = x[k];
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = smem[(16 * hg + e) * W + c];
        // phase B: stage s = 4 + v pairs e-bit (LB-1-v); idx = 2^s + ((16 G + e) >> (S1 - s))
#pragma unroll
        for (int v = 0; v < LB; ++v) ALOHA_CT_STAGE(E, 1 << (LB - 1 - v), (1 << (4 + v)) + ((16 * hg + g0) >> (S1 - 4 - v)))
#pragma unroll
        for (int e = 0; e < E; ++e) dst[(size_t)(16 * hg + e) * 256] = x[e];   // < cols_out_bound<FORM>(S1) q
    }
}

// ============================================================================ forward: rows
// One half-warp per 256-coefficient row.  Padded exchange buffer: word jj lives at jj + 2*(jj>>4)
// so that both the strided writes (h + 16k) and the contiguous 16-byte reads (16g + e) are
// bank-conflict-free.
constexpr int kRowPad = 288;  // 256 + 2 * 16

template <int S1, int FORM>
__global__ void __launch_bounds__(256, ROWS_MINB) ntt_fwd_rows(const NttJob *__restrict__ jobs, u32 total_rows) {
    typedef Arith<FORM> AR;
    constexpr int R = 1 << S1;
    __shared__ u64 smem[16 * kRowPad];
    const int t = threadIdx.x, hw = t >> 4, h = t & 15;
    const u32 grow = blockIdx.x * 16 + hw;           // global row id = job * R + r
    if (grow >= total_rows) return;                   // whole half-warp exits together
    const NttJob &job = jobs[grow / R];
    const u32 r = grow % R;
    const AR A(job.mc);
    // the column pass (if any) has already moved the polynomial to job.dst
    const u64 *src = (S1 == 0 ? job.src : job.dst) + (size_t)r * 256;
    u64 *dst = job.dst + (size_t)r * 256;
    const Tw *tw = job.tw;
    u64 *buf = smem + hw * kRowPad;
    const u32 rr = R + r;

    u64 x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = src[h + 16 * k];
    if (S1 == 0 && job.mc.pre) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = apply_pre(x[k], job.mc.pre, A);
    }
    int B = cols_out_bound<FORM>(S1);          // what the column pass left (2 for a bare 256-point transform)
    // phase A: u = 0..3 pairs k-bit (3-u); idx = 2^u (R + r) + (k >> (4 - u))
#pragma unroll
    for (int u = 0; u < 4; ++u) ALOHA_CT_STAGE(16, 8 >> u, (rr << u) + (g0 >> (4 - u)))
    // exchange h + 16k -> 16g + e
#pragma unroll
    for (int k = 0; k < 16; ++k) buf[h + 18 * k] = x[k];
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(buf + 18 * h + e);
        x[e] = v.x;
        x[e + 1] = v.y;
    }
    // phase B: u = 4..7 pairs e-bit (7-u); idx = 2^u (R + r) + ((16 g + e) >> (8 - u))
#pragma unroll
    for (int u = 4; u < 8; ++u) ALOHA_CT_STAGE(16, 128 >> u, (rr << u) + ((16 * h + g0) >> (8 - u)))
    // < 16q -> canonical; store the thread's 128 contiguous bytes
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        ulonglong2 v;
        v.x = A.canon(x[e]);
        v.y = A.canon(x[e + 1]);
        *reinterpret_cast<ulonglong2 *>(dst + 16 * h + e) = v;
    }
}

// ============================================================================ TMA staging helpers
// The persistent row passes stage their tiles in shared memory with cp.async.bulk / cp.async.bulk.tensor
// (TMA) behind mbarriers; kStages tiles are in flight per CTA.
constexpr int kStages = ROWS_TMA_STAGES;
constexpr int kTileRows = ROWS_TMA_TILE;          // inverse pass: rows (= half-warps) per tile
constexpr int kTilesPerGroupRow = 16 / kTileRows;

__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (TMA, SASS UBLKCP); bytes a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ Tw ldtw_s(const Tw *p) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(p);
    Tw t;
    t.w = v.x;
    t.wp = v.y;
    return t;
}

// ============================================================================ forward: rows, TMA-staged
// Persistent variant of the row pass for launches that hold many polynomials per modulus (the batched
// configs).  One tile = row r of kTileRows8 polynomials that share a modulus (a quarter of an
// NttRowGroup): the rows (2 KiB each, from different polynomials), the row's 256 twiddles in read order
// (4 KiB, NttJob::rtw, row_slot8) and the group record are staged in shared memory by cp.async.bulk (TMA)
// into a ring of kStages stages, each armed with an mbarrier, so the loads of tile i+kStages run under
// the butterflies of the tiles before it.  There is no block-wide barrier: a warp that is done with a
// stage's shared memory counts itself out, and the last one out refills the stage.
// A whole warp per row, 8 coefficients per thread: levels 0..2 on elements lane + 32 k, exchange, levels
// 3..5 on elements hi*32 + m*4 + lo (lane = hi*4 + lo), exchange, levels 6..7 on elements 8 lane + e.
// Both exchanges run in place in the row's staged slot, padded to 288 words (word jj at jj + 2 (jj >> 4)):
// conflict-free for all four access patterns.  64 registers: 8 CTAs of 4 warps per SM.  (The previous
// version -- a half-warp per row, 16 coefficients per thread, 96 registers, one exchange -- had the same
// instruction count per butterfly but half the resident warps, and this pass is limited by warps waiting
// on each other's stage and on dependent integer latency, not by issue: 1.71 -> 1.77 M limb-NTT/s.
// Also measured, none faster: 16- and 4-row tiles of the old shape, a third stage, an XOR-swizzled 2 KiB
// slot, refilling the rows half a tile before the twiddles, prefetching the refill's source addresses.)
constexpr int kTileRows8 = 4;
constexpr int kTilesPerGroupRow8 = 16 / kTileRows8;
struct FwdRowsSmem {
    u64 data[kStages][kTileRows8][kRowPad];
    Tw tw[kStages][256];
    NttRowGroup grp[kStages];
    u64 full[kStages];
    u32 done[kStages];
};
static_assert((sizeof(FwdRowsSmem) + 1024) * 8 <= 233472, "eight CTAs per SM");
#undef ALOHA_LDTW
#define ALOHA_LDTW ldtw_s
template <int S1, int FORM>
__global__ void __launch_bounds__(32 * kTileRows8, 8) ntt_fwd_rows_tma(const NttRowGroup *__restrict__ groups, u32 ntiles) {
    typedef Arith<FORM> AR;
    constexpr int R = 1 << S1;
    constexpr u32 kStageBytes = kTileRows8 * 2048 + 256 * sizeof(Tw) + sizeof(NttRowGroup);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdRowsSmem &S = *reinterpret_cast<FwdRowsSmem *>(smem_raw);
    const int t = threadIdx.x, lane = t & 31, wr = t >> 5, hi = lane >> 2, lo = lane & 3;
    if (t == 0) {
        for (int b = 0; b < kStages; ++b) {
            mbar_init(&S.full[b], 1);
            S.done[b] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto stage_in = [&](u32 tile, int b) {
        const u32 sub = tile % kTilesPerGroupRow8, gr = tile / kTilesPerGroupRow8;
        const NttRowGroup *G = groups + gr / R;
        const size_t off = (size_t)(gr % R) * 256;
        if (lane == 0) mbar_arrive_expect_tx(&S.full[b], kStageBytes);
        __syncwarp();
        if (lane < kTileRows8) bulk_g2s(&S.data[b][lane][0], G->src[sub * kTileRows8 + lane] + off, 2048, &S.full[b]);
        else if (lane == 16) bulk_g2s(&S.tw[b][0], G->rtw + off, 256 * sizeof(Tw), &S.full[b]);
        else if (lane == 17) bulk_g2s(&S.grp[b], G, sizeof(NttRowGroup), &S.full[b]);
    };
    const u32 stride = gridDim.x;
    u32 tile = blockIdx.x;
    if (t < 32) {
        for (int b = 0; b < kStages; ++b)
            if (tile + b * stride < ntiles) stage_in(tile + b * stride, b);
    }
    for (u32 i = 0, b = 0, parity = 0; tile < ntiles; ++i, tile += stride) {
        mbar_wait(&S.full[b], parity);
        const NttRowGroup &G = S.grp[b];
        const AR A(G.mc);
        u64 *dst = G.dst[(tile % kTilesPerGroupRow8) * kTileRows8 + wr] + (size_t)(tile / kTilesPerGroupRow8 % R) * 256;
        const Tw *tw = S.tw[b];
        u64 *row = &S.data[b][wr][0];

        u64 x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = row[lane + 32 * k];
        __syncwarp();                      // the row is in registers: its slot is the exchange buffer now
        if (S1 == 0 && G.mc.pre) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = apply_pre(x[k], G.mc.pre, A);
        }
        int B = cols_out_bound<FORM>(S1);
        // phase A: level u pairs k-bit (2-u); twiddle j = k >> (3 - u), warp-uniform
#pragma unroll
        for (int u = 0; u < 3; ++u) ALOHA_CT_STAGE(8, 4 >> u, row_slot8(u, g0 >> (3 - u)))
        // exchange: lane + 32 k  ->  hi*32 + m*4 + lo
        u64 *wa = row + lane + 2 * (lane >> 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) wa[36 * k] = x[k];
        __syncwarp();
        u64 *rb = row + 36 * hi + lo;
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = rb[4 * m + 2 * (m >> 2)];
        // phase B: level u = 3 + v pairs m-bit (2-v); twiddle j = (hi << v) + (m >> (3 - v))
#pragma unroll
        for (int v = 0; v < 3; ++v) ALOHA_CT_STAGE(8, 4 >> v, row_slot8(3 + v, (hi << v) + (g0 >> (3 - v))))
        // exchange: hi*32 + m*4 + lo  ->  8 lane + e   (every lane has read its phase-B inputs: same words)
#pragma unroll
        for (int m = 0; m < 8; ++m) rb[4 * m + 2 * (m >> 2)] = x[m];
        __syncwarp();
        u64 *rc = row + 8 * lane + 2 * (lane >> 1);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const ulonglong2 v2 = *reinterpret_cast<const ulonglong2 *>(rc + e);
            x[e] = v2.x;
            x[e + 1] = v2.y;
        }
        // phase C: level u = 6 + v pairs e-bit (1-v); twiddle j = (lane << (v + 1)) + (e >> (2 - v))
#pragma unroll
        for (int v = 0; v < 2; ++v) ALOHA_CT_STAGE(8, 2 >> v, row_slot8(6 + v, (lane << (v + 1)) + (g0 >> (2 - v))))
        // leave the stage, then canonical values and 16-byte stores (see ntt_fwd_rows_tma)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        u32 last = 0;
        if (lane == 0) {
            last = atomicAdd(&S.done[b], 1u) == kTileRows8 - 1;
            if (last) S.done[b] = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && tile + kStages * stride < ntiles) stage_in(tile + kStages * stride, b);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
            const u64 a0 = A.canon_lazy(x[e]), a1 = A.canon_lazy(x[e + 1]);     // below 2q
            const u64 b0 = a0 - A.q, b1 = a1 - A.q;
            const bool n0 = (long long)b0 < 0, n1 = (long long)b1 < 0;
            const u32 w0 = n0 ? (u32)a0 : (u32)b0, w1 = n0 ? (u32)(a0 >> 32) : (u32)(b0 >> 32);
            const u32 w2 = n1 ? (u32)a1 : (u32)b1, w3 = n1 ? (u32)(a1 >> 32) : (u32)(b1 >> 32);
            asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 8 * lane + e), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
        }
        if (++b == kStages) { b = 0; parity ^= 1; }
    }
}
#undef ALOHA_LDTW
#define ALOHA_LDTW ldtw

// ============================================================================ inverse: rows
// GS stages lt = 0..7 (gap 2^lt).  idx = (N >> (lt+1)) + (j >> (lt+1)),  j = r*256 + jj.
//
// Lazy sums: out_x = x + y (bound b_x + b_y), out_y = (x - y + off(b_y)) * w (bound MULB).  bnd[] carries
// the per-register bounds (units of q) through the unrolled stages at compile time; an input is reduced
// (A.red) only when b_x + b_y would pass 16, and a block ends by bringing every register back under 2q
// -- 16 conditional subtracts per 32 butterflies instead of Harvey's 32 (FORM_GENERIC), 16 folds (FORM_PM).
#define ALOHA_GS_REDUCE_PAIR(I, J)                                                                  \
    if (bnd[I] + bnd[J] > 16) {                                                                     \
        if (bnd[I] > AR::REDB) { x[I] = A.red(x[I]); bnd[I] = AR::REDB; }                           \
        if (bnd[J] > AR::REDB) { x[J] = A.red(x[J]); bnd[J] = AR::REDB; }                           \
    }                                                                                               \
    const int by_ = bnd[J];                                                                         \
    if (bnd[I] + (by_ <= 2 ? 2 : by_ <= 4 ? 4 : 8) > 16) __trap();   /* folds away: never true */
#define ALOHA_GS_STAGE(NELEM, HALF, TWIDX)                                                          \
    {                                                                                               \
        Tw w_;                                                                                      \
        _Pragma("unroll") for (int e = 0; e < (NELEM); ++e) {                                       \
            if (e & (HALF)) continue;                                                               \
            if ((e & ((HALF)-1)) == 0) {                                                            \
                const int g0 = e & ~(2 * (HALF)-1);                                                 \
                w_ = ALOHA_LDTW(tw + (TWIDX));                                                      \
            }                                                                                       \
            ALOHA_GS_REDUCE_PAIR(e, e + (HALF))                                                     \
            const u64 d_ = x[e] - x[e + (HALF)] + A.off(by_);                                       \
            x[e] = x[e] + x[e + (HALF)];                                                            \
            x[e + (HALF)] = A.mul(d_, w_.w, w_.wp);                                                 \
            bnd[e] += by_;                                                                          \
            bnd[e + (HALF)] = AR::MULB;                                                             \
        }                                                                                           \
    }
// bring every register back under 2q
#define ALOHA_GS_NORMALISE(NELEM)                                                                   \
    _Pragma("unroll") for (int e = 0; e < (NELEM); ++e) {                                           \
        x[e] = A.norm(x[e], bnd[e]);                                                                \
        bnd[e] = 2;                                                                                 \
    }
// last stage of the whole transform: N^-1 folded in, both outputs multiplied, canonical results
#define ALOHA_GS_LAST(NELEM)                                                                        \
    _Pragma("unroll") for (int i = 0; i < (NELEM) / 2; ++i) {                                       \
        ALOHA_GS_REDUCE_PAIR(i, i + (NELEM) / 2)                                                    \
        const u64 s_ = x[i] + x[i + (NELEM) / 2], d_ = x[i] - x[i + (NELEM) / 2] + A.off(by_);      \
        x[i] = A.canon_mul(A.mul(s_, job.mc.ninv, job.mc.ninv_p));                                  \
        x[i + (NELEM) / 2] = A.canon_mul(A.mul(d_, job.mc.wninv, job.mc.wninv_p));                  \
    }

template <int S1, int FORM>
__global__ void __launch_bounds__(256, ROWS_MINB) ntt_inv_rows(const NttJob *__restrict__ jobs, u32 total_rows) {
    typedef Arith<FORM> AR;
    constexpr int R = 1 << S1;
    constexpr int LOGN = S1 + 8;
    __shared__ u64 smem[16 * kRowPad];
    const int t = threadIdx.x, hw = t >> 4, h = t & 15;
    const u32 grow = blockIdx.x * 16 + hw;
    if (grow >= total_rows) return;
    const NttJob &job = jobs[grow / R];
    const u32 r = grow % R;
    const AR A(job.mc);
    const u64 *src = job.src + (size_t)r * 256;
    u64 *dst = job.dst + (size_t)r * 256;
    const Tw *tw = job.tw;
    u64 *buf = smem + hw * kRowPad;

    u64 x[16];
    int bnd[16];
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + 16 * h + e);
        x[e] = v.x;          // < 2q (see header)
        x[e + 1] = v.y;
        bnd[e] = bnd[e + 1] = 2;
    }
    // lt = 0..3 pair e-bit lt
#pragma unroll
    for (int lt = 0; lt < 4; ++lt)
        ALOHA_GS_STAGE(16, 1 << lt, (1u << (LOGN - 1 - lt)) + (r << (7 - lt)) + ((16 * h + g0) >> (lt + 1)))
    ALOHA_GS_NORMALISE(16)
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        ulonglong2 v;
        v.x = x[e];
        v.y = x[e + 1];
        *reinterpret_cast<ulonglong2 *>(buf + 18 * h + e) = v;
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = buf[h + 18 * k];
    // lt = 4..7 pair k-bit (lt-4); idx = base + (k >> (lt - 3))
#pragma unroll
    for (int lt = 4; lt < 8; ++lt) {
        if (S1 == 0 && lt == 7) {
            ALOHA_GS_LAST(16)
        } else {
            ALOHA_GS_STAGE(16, 1 << (lt - 4), (1u << (LOGN - 1 - lt)) + (r << (7 - lt)) + (g0 >> (lt - 3)))
        }
    }
    if (S1 != 0) ALOHA_GS_NORMALISE(16)
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[h + 16 * k] = x[k];  // < 2q (canonical when S1 == 0)
}

// ============================================================================ inverse: rows, TMA-staged
// The mirror image of ntt_fwd_rows_tma.  A thread starts from 16 CONTIGUOUS coefficients (line h of the
// row), which a linear copy would deliver with 8-way bank conflicts, so the rows come in through a
// tensor map with SWIZZLE_128B (kernels.cuh TmaMaps): chunk c of line l sits at chunk c ^ (l mod 8), the 8
// lanes of an LDS.128 phase read 8 different chunks, and the exchange runs in place in the same layout.
// Twiddles: the row's 255 inverse twiddles, level lt stored at row_slot(7 - lt, .).
struct InvRowsSmem {
    u64 data[kStages][kTileRows][256];     // first member: the swizzle pattern wants 1 KiB alignment
    Tw tw[kStages][256];
    NttRowGroup grp[kStages];
    u64 full[kStages];
    u32 done[kStages];
};
static_assert(sizeof(InvRowsSmem) + 1024 <= (233472 - ROWS_TMA_MINB * 1024) / ROWS_TMA_MINB, "ROWS_TMA_MINB CTAs per SM");

__device__ __forceinline__ void tensor_g2s_2d(void *dst, const CUtensorMap *map, u32 c0, u32 c1, u64 *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_addr(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar)) : "memory");
}

#undef ALOHA_LDTW
#define ALOHA_LDTW ldtw_s
template <int S1, int FORM>
__global__ void __launch_bounds__(16 * kTileRows, ROWS_TMA_MINB)
ntt_inv_rows_tma(const NttRowGroup *__restrict__ groups, u32 ntiles, const __grid_constant__ TmaMaps maps) {
    typedef Arith<FORM> AR;
    constexpr int R = 1 << S1;
    constexpr u32 kStageBytes = kTileRows * 2048 + 256 * sizeof(Tw) + sizeof(NttRowGroup);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // the swizzle pattern is a function of the shared-memory address: round the window up to 1 KiB (the
    // launch allocates 1 KiB of slack)
    InvRowsSmem &S = *reinterpret_cast<InvRowsSmem *>(smem_raw + ((1024 - (smem_addr(smem_raw) & 1023)) & 1023));
    const int t = threadIdx.x, lane = t & 31, hw = t >> 4, h = t & 15;
    if (t == 0) {
        for (int b = 0; b < kStages; ++b) {
            mbar_init(&S.full[b], 1);
            S.done[b] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto stage_in = [&](u32 tile, int b) {
        const u32 sub = tile % kTilesPerGroupRow, gr = tile / kTilesPerGroupRow;
        const NttRowGroup *G = groups + gr / R;
        const u32 r = gr % R;
        if (lane == 0) mbar_arrive_expect_tx(&S.full[b], kStageBytes);
        __syncwarp();
        if (lane < kTileRows) {
            const int k = sub * kTileRows + lane;
            tensor_g2s_2d(&S.data[b][lane][0], &maps.m[G->src_map[k]], 0, G->src_line[k] + 16 * r, &S.full[b]);
        } else if (lane == 16) {
            bulk_g2s(&S.tw[b][0], G->rtw + (size_t)r * 256, 256 * sizeof(Tw), &S.full[b]);
        } else if (lane == 17) {
            bulk_g2s(&S.grp[b], G, sizeof(NttRowGroup), &S.full[b]);
        }
    };
    const u32 stride = gridDim.x;
    u32 tile = blockIdx.x;
    if (t < 32) {
        for (int b = 0; b < kStages; ++b)
            if (tile + b * stride < ntiles) stage_in(tile + b * stride, b);
    }
    for (u32 i = 0, b = 0, parity = 0; tile < ntiles; ++i, tile += stride) {
        mbar_wait(&S.full[b], parity);
        const NttRowGroup &job = S.grp[b];
        const AR A(job.mc);
        u64 *dst = job.dst[(tile % kTilesPerGroupRow) * kTileRows + hw] + (size_t)(tile / kTilesPerGroupRow % R) * 256;
        const Tw *tw = S.tw[b];
        u64 *row = &S.data[b][hw][0];
        u64 *mine = row + 16 * h;          // line h: this thread's 16 contiguous coefficients, chunks XOR (h mod 8)

        u64 x[16];
        int bnd[16];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(mine + (((e >> 1) ^ (h & 7)) << 1));
            x[e] = v.x;          // < 2q (see header)
            x[e + 1] = v.y;
            bnd[e] = bnd[e + 1] = 2;
        }
        // lt = 0..3 pair e-bit lt; twiddle j = (16 h + e) >> (lt + 1) of level lt
#pragma unroll
        for (int lt = 0; lt < 4; ++lt) ALOHA_GS_STAGE(16, 1 << lt, row_slot(7 - lt, (16 * h + g0) >> (lt + 1)))
        ALOHA_GS_NORMALISE(16)
        // exchange 16 h + e -> h + 16 k, in place: a thread overwrites exactly the words it read
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            ulonglong2 v;
            v.x = x[e];
            v.y = x[e + 1];
            *reinterpret_cast<ulonglong2 *>(mine + (((e >> 1) ^ (h & 7)) << 1)) = v;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = row[16 * k + ((((h >> 1) ^ (k & 7)) << 1) | (h & 1))];
        // lt = 4..7 pair k-bit (lt-4); twiddle j = k >> (lt - 3)
#pragma unroll
        for (int lt = 4; lt < 8; ++lt) {
            if (S1 == 0 && lt == 7) {
                ALOHA_GS_LAST(16)
            } else {
                ALOHA_GS_STAGE(16, 1 << (lt - 4), row_slot(7 - lt, g0 >> (lt - 3)))
            }
        }
        if (S1 != 0) ALOHA_GS_NORMALISE(16)
        // leave the stage (see ntt_fwd_rows_tma), then store
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        u32 last = 0;
        if (lane == 0) {
            last = atomicAdd(&S.done[b], 1u) == kTileRows / 2 - 1;
            if (last) S.done[b] = 0;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last && tile + kStages * stride < ntiles) stage_in(tile + kStages * stride, b);
#pragma unroll
        for (int k = 0; k < 16; ++k) asm volatile("st.global.u64 [%0], %1;" ::"l"(dst + h + 16 * k), "l"(x[k]) : "memory");
        if (++b == kStages) { b = 0; parity ^= 1; }
    }
}
#undef ALOHA_LDTW
#define ALOHA_LDTW ldtw

// ============================================================================ inverse: columns
// GS stages lt = 8 .. 8+S1-1, row-distance bit b = lt - 8.  m = 2^(S1-1-b), idx = m + (r >> (b+1)).
template <int S1, int FORM>
__global__ void __launch_bounds__(256, COLS_MINB) ntt_inv_cols(const NttJob *__restrict__ jobs) {
    typedef Arith<FORM> AR;
    constexpr int LA = S1 < 4 ? S1 : 4, LB = S1 - LA, E = 1 << LA, R = 1 << S1;
    constexpr int H = R / E, W = 256 / H, TILES = H;
    extern __shared__ u64 smem[];

    const NttJob &job = jobs[blockIdx.x / TILES];
    const int c0 = (blockIdx.x % TILES) * W;
    const int t = threadIdx.x, c = t % W, hg = t / W;
    const AR A(job.mc);
    const u64 *src = job.dst + c0 + c;   // the row pass has already moved the polynomial to job.dst
    u64 *dst = job.dst + c0 + c;
    const Tw *tw = job.tw;

    u64 x[E];
    int bnd[E];
#pragma unroll
    for (int e = 0; e < E; ++e) bnd[e] = 2;
    if constexpr (LB > 0) {
        // low LB stages on contiguous rows 16 G + e
#pragma unroll
        for (int e = 0; e < E; ++e) x[e] = src[(size_t)(16 * hg + e) * 256];
#pragma unroll
        for (int b = 0; b < LB; ++b)
            ALOHA_GS_STAGE(E, 1 << b, (1 << (S1 - 1 - b)) + ((16 * hg + g0) >> (b + 1)))
        ALOHA_GS_NORMALISE(E)
#pragma unroll
        for (int e = 0; e < E; ++e) smem[(16 * hg + e) * W + c] = x[e];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = smem[(hg + H * k) * W + c];
    } else {
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = src[(size_t)(hg + H * k) * 256];
    }
    // high LA stages on rows h + H k: b = LB .. S1-1 pairs k-bit (b - LB); idx = m + (k >> (b+1-LB))
#pragma unroll
    for (int b = LB; b < S1; ++b) {
        if (b == S1 - 1) {
            ALOHA_GS_LAST(E)
        } else {
            ALOHA_GS_STAGE(E, 1 << (b - LB), (1 << (S1 - 1 - b)) + (g0 >> (b + 1 - LB)))
        }
    }
#pragma unroll
    for (int k = 0; k < E; ++k) dst[(size_t)(hg + H * k) * 256] = x[k];
}

// ============================================================================ launchers
std::atomic<unsigned long long> g_launches{0};
unsigned long long kernel_launch_count() { return g_launches.load(std::memory_order_relaxed); }
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// SM count of the CURRENT device (engines on different GPUs may share this process)
static int sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

template <int S1, int FORM>
static cudaError_t fwd_impl(const NttJob *jobs, u32 njobs, const NttRowGroup *groups, u32 ngroups, cudaStream_t st) {
    constexpr int R = 1 << S1;
    if constexpr (S1 > 0) {
        constexpr int LA = S1 < 4 ? S1 : 4, H = R >> LA;
        const size_t smem = S1 > 4 ? (size_t)4096 * 8 : 0;
        ntt_fwd_cols<S1, FORM><<<njobs * H, 256, smem, st>>>(jobs);
        count_launch();
    }
    if (ngroups) {
        static int resident_dev[64] = {};   // CTAs of the persistent row pass that fit on one SM (function
        int dev = 0;                        // attributes are per device: set them on each one the process uses)
        cudaGetDevice(&dev);
        int &resident = resident_dev[dev & 63];
        if (!resident) {
            cudaError_t e = cudaFuncSetAttribute(ntt_fwd_rows_tma<S1, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FwdRowsSmem));
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, ntt_fwd_rows_tma<S1, FORM>, 32 * kTileRows8, sizeof(FwdRowsSmem));
            if (e != cudaSuccess) return e;
            if (resident < 1) return cudaErrorLaunchOutOfResources;
        }
        const u32 tiles = ngroups * R * kTilesPerGroupRow8, ctas = (u32)(resident * sm_count());
        ntt_fwd_rows_tma<S1, FORM><<<tiles < ctas ? tiles : ctas, 32 * kTileRows8, sizeof(FwdRowsSmem), st>>>(groups, tiles);
        count_launch();
    }
    if (njobs > 16 * ngroups) {
        const u32 rows = (njobs - 16 * ngroups) * R;
        ntt_fwd_rows<S1, FORM><<<(rows + 15) / 16, 256, 0, st>>>(jobs + 16 * ngroups, rows);
        count_launch();
    }
    return cudaGetLastError();
}

template <int S1, int FORM>
static cudaError_t inv_impl(const NttJob *jobs, u32 njobs, const NttRowGroup *groups, u32 ngroups, const TmaMaps *maps,
                            cudaStream_t st) {
    constexpr int R = 1 << S1;
    if (ngroups) {
        constexpr size_t kSmem = sizeof(InvRowsSmem) + 1024;
        static int resident_dev[64] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        int &resident = resident_dev[dev & 63];
        if (!resident) {
            cudaError_t e = cudaFuncSetAttribute(ntt_inv_rows_tma<S1, FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, ntt_inv_rows_tma<S1, FORM>, 16 * kTileRows, kSmem);
            if (e != cudaSuccess) return e;
            if (resident < 1) return cudaErrorLaunchOutOfResources;
        }
        const u32 tiles = ngroups * R * kTilesPerGroupRow, ctas = (u32)(resident * sm_count());
        ntt_inv_rows_tma<S1, FORM><<<tiles < ctas ? tiles : ctas, 16 * kTileRows, kSmem, st>>>(groups, tiles, *maps);
        count_launch();
    }
    if (njobs > 16 * ngroups) {
        const u32 rows = (njobs - 16 * ngroups) * R;
        ntt_inv_rows<S1, FORM><<<(rows + 15) / 16, 256, 0, st>>>(jobs + 16 * ngroups, rows);
        count_launch();
    }
    if constexpr (S1 > 0) {
        constexpr int LA = S1 < 4 ? S1 : 4, H = R >> LA;
        const size_t smem = S1 > 4 ? (size_t)4096 * 8 : 0;
        ntt_inv_cols<S1, FORM><<<njobs * H, 256, smem, st>>>(jobs);
        count_launch();
    }
    return cudaGetLastError();
}

// A forward job runs columns src->dst then rows dst->dst; an inverse job rows src->dst then columns
// dst->dst.  src == dst (exactly) is allowed: every CTA / half-warp reads its whole tile before it
// writes it.  Partially overlapping src / dst is the caller's bug.
#define ALOHA_NTT_DISPATCH(IMPL, ...)                                                          \
    if (form > FORM_PM) return cudaErrorInvalidValue;                                     \
    switch (logn * 2 + form) {                                                            \
    case 16: return IMPL<0, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 17: return IMPL<0, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 18: return IMPL<1, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 19: return IMPL<1, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 20: return IMPL<2, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 21: return IMPL<2, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 22: return IMPL<3, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 23: return IMPL<3, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 24: return IMPL<4, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 25: return IMPL<4, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 26: return IMPL<5, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 27: return IMPL<5, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 28: return IMPL<6, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 29: return IMPL<6, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 30: return IMPL<7, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 31: return IMPL<7, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    case 32: return IMPL<8, FORM_GENERIC>(jobs, njobs, groups, ngroups, __VA_ARGS__); case 33: return IMPL<8, FORM_PM>(jobs, njobs, groups, ngroups, __VA_ARGS__); \
    default: return cudaErrorInvalidValue;                                                \
    }
cudaError_t launch_ntt_forward(const NttJob *jobs, u32 njobs, const NttRowGroup *groups, u32 ngroups, u32 logn, u32 form,
                               cudaStream_t st) {
    if (16 * ngroups > njobs) return cudaErrorInvalidValue;
    ALOHA_NTT_DISPATCH(fwd_impl, st)
}
cudaError_t launch_ntt_inverse(const NttJob *jobs, u32 njobs, const NttRowGroup *groups, u32 ngroups, const TmaMaps *maps,
                               u32 logn, u32 form, cudaStream_t st) {
    if (16 * ngroups > njobs || (ngroups && !maps)) return cudaErrorInvalidValue;
    ALOHA_NTT_DISPATCH(inv_impl, maps, st)
}

}  // namespace alb
