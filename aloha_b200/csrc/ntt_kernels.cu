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
#include <cstdlib>

#include "kernels.cuh"
#include "modarith.cuh"

#ifndef ROWS_MINB
#define ROWS_MINB 2
#endif
#ifndef COLS_MINB
#define COLS_MINB 3
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

// CT butterfly, lazy: (x, y) -> (x + w y, x - w y + 2q).  Bound grows by 2q.  x seeds the
// multiply-add chain, so x' costs no add; y' = 2x + 2q - x'.
__device__ __forceinline__ void ct_bf(u64 &x, u64 &y, const Tw &t, u64 nq, u64 q2) {
    const u64 xp = shoup_mac(x, y, t.w, t.wp, nq);
    y = (x + x + q2) - xp;
    x = xp;
}


// Base-extension primitive folded into the load of a forward transform, word-exact:
//   VCPY   = addmod(r(x), 0)  -> two compare-based conditional subtracts (expander.v:396-417);
//   VFQMOD = barrett(r(x), 1) -> x mod q for EVERY 64-bit x (the RTL's quotient estimate is within one
//            of floor(x/q) for 60-bit q), i.e. exactly what reduce_full computes.
__device__ __forceinline__ u64 apply_pre(u64 x, u32 pre, u64 q, u64 nq, u32 mest) {
    if (pre == PRE_VCPY) {
        x = x >= q ? x - q : x;
        x = x >= q ? x - q : x;
    } else if (pre == PRE_VFQMOD) {
        x = reduce_full(x, q, nq, mest);
    }
    return x;
}

// Bound (units of q) of what the forward column pass stores for S1 column stages: start at 2, +2 per
// stage, an upper-input reduction to 8 whenever the next stage would pass 16.
__host__ __device__ constexpr int cols_out_bound(int s1) {
    int b = 2;
    for (int s = 0; s < s1; ++s) b = (b + 2 > 16 ? 8 : b) + 2;
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
// ever reduced (one 8q conditional subtract), and only when B + 2 would pass 16.
#define ALOHA_CT_STAGE(NELEM, HALF, TWIDX)                                           \
    {                                                                                \
        const bool red_ = B + 2 > 16;                                                \
        Tw w_;                                                                       \
        _Pragma("unroll") for (int e = 0; e < (NELEM); ++e) {                        \
            if (e & (HALF)) continue;                                                \
            if ((e & ((HALF)-1)) == 0) {                                             \
                const int g0 = e & ~(2 * (HALF)-1);                                  \
                w_ = ldtw(tw + (TWIDX));                                             \
            }                                                                        \
            if (red_) x[e] = csub_s(x[e], q8);                                       \
            ct_bf(x[e], x[e + (HALF)], w_, nq, q2);                                  \
        }                                                                            \
        B = (red_ ? 8 : B) + 2;                                                      \
    }

template <int S1>
__global__ void __launch_bounds__(256, COLS_MINB) ntt_fwd_cols(const NttJob *__restrict__ jobs) {
    constexpr int LA = S1 < 4 ? S1 : 4, LB = S1 - LA, E = 1 << LA, R = 1 << S1;
    constexpr int H = R / E;             // threads per column in phase A
    constexpr int W = 256 / H;           // tile width in columns
    constexpr int TILES = H;             // tiles per polynomial
    extern __shared__ u64 smem[];        // [R][W] when LB > 0

    const NttJob &job = jobs[blockIdx.x / TILES];
    const int c0 = (blockIdx.x % TILES) * W;
    const int t = threadIdx.x, c = t % W, hg = t / W;
    const u64 q = job.mc.q, q2 = 2 * q, q8 = 8 * q, nq = 0 - q;
    const u64 *src = job.src + c0 + c;
    u64 *dst = job.dst + c0 + c;
    const Tw *tw = job.tw;

    u64 x[E];
#pragma unroll
    for (int k = 0; k < E; ++k) x[k] = src[(size_t)(hg + H * k) * 256];   // < 2q (see header)
    if (job.mc.pre) {
#pragma unroll
        for (int k = 0; k < E; ++k) x[k] = apply_pre(x[k], job.mc.pre, q, nq, job.mc.mest);
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
        for (int e = 0; e < E; ++e) dst[(size_t)(16 * hg + e) * 256] = x[e];   // < cols_out_bound(S1) q
    }
}

// ============================================================================ forward: rows
// One half-warp per 256-coefficient row.  Padded exchange buffer: word jj lives at jj + 2*(jj>>4)
// so that both the strided writes (h + 16k) and the contiguous 16-byte reads (16g + e) are
// bank-conflict-free.
constexpr int kRowPad = 288;  // 256 + 2 * 16

template <int S1>
__global__ void __launch_bounds__(256, ROWS_MINB) ntt_fwd_rows(const NttJob *__restrict__ jobs, u32 total_rows) {
    constexpr int R = 1 << S1;
    __shared__ u64 smem[16 * kRowPad];
    const int t = threadIdx.x, hw = t >> 4, h = t & 15;
    const u32 grow = blockIdx.x * 16 + hw;           // global row id = job * R + r
    if (grow >= total_rows) return;                   // whole half-warp exits together
    const NttJob &job = jobs[grow / R];
    const u32 r = grow % R;
    const u64 q = job.mc.q, q2 = 2 * q, q8 = 8 * q, nq = 0 - q;
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
        for (int k = 0; k < 16; ++k) x[k] = apply_pre(x[k], job.mc.pre, q, nq, job.mc.mest);
    }
    int B = cols_out_bound(S1);          // what the column pass left (2 for a bare 256-point transform)
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
    const u32 mest = job.mc.mest;
#pragma unroll
    for (int e = 0; e < 16; e += 2) {
        ulonglong2 v;
        v.x = reduce_full(x[e], q, nq, mest);
        v.y = reduce_full(x[e + 1], q, nq, mest);
        *reinterpret_cast<ulonglong2 *>(dst + 16 * h + e) = v;
    }
}

// ============================================================================ inverse: rows
// GS stages lt = 0..7 (gap 2^lt).  idx = (N >> (lt+1)) + (j >> (lt+1)),  j = r*256 + jj.
//
// Lazy sums: out_x = x + y (bound b_x + b_y), out_y = (x - y + b_y q) * w (bound 2).  bnd[] carries the
// per-register bounds (units of q) through the unrolled stages at compile time; an input is reduced
// (8q conditional subtract) only when b_x + b_y would pass 16, and a block ends by bringing every
// register back under 2q -- 16 conditional subtracts per 32 butterflies instead of Harvey's 32.
#define ALOHA_GS_STAGE(NELEM, HALF, TWIDX)                                                          \
    {                                                                                               \
        Tw w_;                                                                                      \
        _Pragma("unroll") for (int e = 0; e < (NELEM); ++e) {                                       \
            if (e & (HALF)) continue;                                                               \
            if ((e & ((HALF)-1)) == 0) {                                                            \
                const int g0 = e & ~(2 * (HALF)-1);                                                 \
                w_ = ldtw(tw + (TWIDX));                                                            \
            }                                                                                       \
            if (bnd[e] + bnd[e + (HALF)] > 16) {                                                    \
                if (bnd[e] > 8) { x[e] = csub_s(x[e], q8); bnd[e] = 8; }                            \
                if (bnd[e + (HALF)] > 8) { x[e + (HALF)] = csub_s(x[e + (HALF)], q8); bnd[e + (HALF)] = 8; } \
            }                                                                                       \
            const int by_ = bnd[e + (HALF)];                                                        \
            const u64 off_ = by_ <= 2 ? q2 : by_ <= 4 ? q4 : q8;                                    \
            const u64 d_ = x[e] - x[e + (HALF)] + off_;                                             \
            x[e] = x[e] + x[e + (HALF)];                                                            \
            x[e + (HALF)] = mul_shoup(d_, w_.w, w_.wp, nq);                                         \
            bnd[e] += by_;                                                                          \
            bnd[e + (HALF)] = 2;                                                                    \
        }                                                                                           \
    }
// bring every register back under 2q
#define ALOHA_GS_NORMALISE(NELEM)                                                                   \
    _Pragma("unroll") for (int e = 0; e < (NELEM); ++e) {                                           \
        if (bnd[e] > 8) x[e] = csub_s(x[e], q8);                                                    \
        if (bnd[e] > 4) x[e] = csub_s(x[e], q4);                                                    \
        if (bnd[e] > 2) x[e] = csub_s(x[e], q2);                                                    \
        bnd[e] = 2;                                                                                 \
    }
// last stage of the whole transform: N^-1 folded in, both outputs multiplied, canonical results
#define ALOHA_GS_LAST(NELEM)                                                                        \
    _Pragma("unroll") for (int i = 0; i < (NELEM) / 2; ++i) {                                       \
        if (bnd[i] + bnd[i + (NELEM) / 2] > 16) {                                                   \
            if (bnd[i] > 8) { x[i] = csub_s(x[i], q8); bnd[i] = 8; }                                \
            if (bnd[i + (NELEM) / 2] > 8) { x[i + (NELEM) / 2] = csub_s(x[i + (NELEM) / 2], q8); bnd[i + (NELEM) / 2] = 8; } \
        }                                                                                           \
        const int by_ = bnd[i + (NELEM) / 2];                                                       \
        const u64 off_ = by_ <= 2 ? q2 : by_ <= 4 ? q4 : q8;                                        \
        const u64 s_ = x[i] + x[i + (NELEM) / 2], d_ = x[i] - x[i + (NELEM) / 2] + off_;            \
        x[i] = csub_s(mul_shoup(s_, job.mc.ninv, job.mc.ninv_p, nq), q);                            \
        x[i + (NELEM) / 2] = csub_s(mul_shoup(d_, job.mc.wninv, job.mc.wninv_p, nq), q);            \
    }

template <int S1>
__global__ void __launch_bounds__(256, ROWS_MINB) ntt_inv_rows(const NttJob *__restrict__ jobs, u32 total_rows) {
    constexpr int R = 1 << S1;
    constexpr int LOGN = S1 + 8;
    __shared__ u64 smem[16 * kRowPad];
    const int t = threadIdx.x, hw = t >> 4, h = t & 15;
    const u32 grow = blockIdx.x * 16 + hw;
    if (grow >= total_rows) return;
    const NttJob &job = jobs[grow / R];
    const u32 r = grow % R;
    const u64 q = job.mc.q, q2 = 2 * q, q4 = 4 * q, q8 = 8 * q, nq = 0 - q;
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

// ============================================================================ inverse: columns
// GS stages lt = 8 .. 8+S1-1, row-distance bit b = lt - 8.  m = 2^(S1-1-b), idx = m + (r >> (b+1)).
template <int S1>
__global__ void __launch_bounds__(256, COLS_MINB) ntt_inv_cols(const NttJob *__restrict__ jobs) {
    constexpr int LA = S1 < 4 ? S1 : 4, LB = S1 - LA, E = 1 << LA, R = 1 << S1;
    constexpr int H = R / E, W = 256 / H, TILES = H;
    extern __shared__ u64 smem[];

    const NttJob &job = jobs[blockIdx.x / TILES];
    const int c0 = (blockIdx.x % TILES) * W;
    const int t = threadIdx.x, c = t % W, hg = t / W;
    const u64 q = job.mc.q, q2 = 2 * q, q4 = 4 * q, q8 = 8 * q, nq = 0 - q;
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
unsigned long long g_launches = 0;
unsigned long long kernel_launch_count() { return g_launches; }
static inline void count_launch() { ++g_launches; }

template <int S1>
static cudaError_t fwd_impl(const NttJob *jobs, u32 njobs, cudaStream_t st) {
    constexpr int R = 1 << S1;
    if constexpr (S1 > 0) {
        constexpr int LA = S1 < 4 ? S1 : 4, H = R >> LA;
        const size_t smem = S1 > 4 ? (size_t)4096 * 8 : 0;
        ntt_fwd_cols<S1><<<njobs * H, 256, smem, st>>>(jobs);
        count_launch();
    }
    const u32 rows = njobs * R;
    ntt_fwd_rows<S1><<<(rows + 15) / 16, 256, 0, st>>>(jobs, rows);
    count_launch();
    return cudaGetLastError();
}

template <int S1>
static cudaError_t inv_impl(const NttJob *jobs, u32 njobs, cudaStream_t st) {
    constexpr int R = 1 << S1;
    const u32 rows = njobs * R;
    ntt_inv_rows<S1><<<(rows + 15) / 16, 256, 0, st>>>(jobs, rows);
    count_launch();
    if constexpr (S1 > 0) {
        constexpr int LA = S1 < 4 ? S1 : 4, H = R >> LA;
        const size_t smem = S1 > 4 ? (size_t)4096 * 8 : 0;
        ntt_inv_cols<S1><<<njobs * H, 256, smem, st>>>(jobs);
        count_launch();
    }
    return cudaGetLastError();
}

// A forward job runs columns src->dst then rows dst->dst; an inverse job rows src->dst then columns
// dst->dst.  src == dst (exactly) is allowed: every CTA / half-warp reads its whole tile before it
// writes it.  Partially overlapping src / dst is the caller's bug.
cudaError_t launch_ntt_forward(const NttJob *jobs, u32 njobs, u32 logn, cudaStream_t st) {
    switch (logn) {
    case 8: return fwd_impl<0>(jobs, njobs, st);
    case 9: return fwd_impl<1>(jobs, njobs, st);
    case 10: return fwd_impl<2>(jobs, njobs, st);
    case 11: return fwd_impl<3>(jobs, njobs, st);
    case 12: return fwd_impl<4>(jobs, njobs, st);
    case 13: return fwd_impl<5>(jobs, njobs, st);
    case 14: return fwd_impl<6>(jobs, njobs, st);
    case 15: return fwd_impl<7>(jobs, njobs, st);
    case 16: return fwd_impl<8>(jobs, njobs, st);
    default: return cudaErrorInvalidValue;
    }
}
cudaError_t launch_ntt_inverse(const NttJob *jobs, u32 njobs, u32 logn, cudaStream_t st) {
    switch (logn) {
    case 8: return inv_impl<0>(jobs, njobs, st);
    case 9: return inv_impl<1>(jobs, njobs, st);
    case 10: return inv_impl<2>(jobs, njobs, st);
    case 11: return inv_impl<3>(jobs, njobs, st);
    case 12: return inv_impl<4>(jobs, njobs, st);
    case 13: return inv_impl<5>(jobs, njobs, st);
    case 14: return inv_impl<6>(jobs, njobs, st);
    case 15: return inv_impl<7>(jobs, njobs, st);
    case 16: return inv_impl<8>(jobs, njobs, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace alb
