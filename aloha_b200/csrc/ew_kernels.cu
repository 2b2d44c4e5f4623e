// ew_kernels.cu -- element-wise RNS ALU, automorphism / rotation and copy kernels for sm_100a.
//
// Reference functions replaced:
//   modalu (src/vp/vxu/modalu.sv:22-46,152-249,351-379)  -> ew_kernel<OP>, mac_kernel
//   VAUT / VROLI address+sign logic (src/vp/vxu/vxu_lane.sv:594-599) + the Benes-style lane
//   interconnect (src/vp/iconn/iconn_top.sv:72-138)       -> perm kernels (gather form)
//   VLE / VSE (src/vp/vmu/*, spm.sv)                      -> copy_kernel (only when the batcher
//                                                            cannot alias the access away)
// All of them are HBM-bound streams: 16-byte vector accesses, one job table per launch
// (blockIdx.y = job), grids sized so every SM holds several CTAs.
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "kernels.cuh"
#include "modarith.cuh"

namespace alb {

extern std::atomic<unsigned long long> g_launches;

namespace {

constexpr int kThreads = 256;
constexpr int kVec = 2;                       // u64 per 16-byte access
constexpr int kPerThread = 4;                 // two 16-byte accesses in flight per operand
constexpr int kPerBlock = kThreads * kPerThread;

__device__ __forceinline__ ulonglong2 ld2(const u64 *p) {
    return *reinterpret_cast<const ulonglong2 *>(p);
}
__device__ __forceinline__ void st2(u64 *p, u64 a, u64 b) {
    ulonglong2 v;
    v.x = a;
    v.y = b;
    *reinterpret_cast<ulonglong2 *>(p) = v;
}

template <u32 OP>
__global__ void __launch_bounds__(kThreads) ew_kernel(const EwJob *__restrict__ jobs, u32 n) {
    const EwJob job = jobs[blockIdx.y];
    constexpr bool kVV = OP == ALU_MUL_VV || OP == ALU_ADD_VV || OP == ALU_SUB_VV;
    const u64 q = job.q, iq = job.iq, s = job.s;
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 i = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (i >= n) return;
        const ulonglong2 a = ld2(job.a + i);
        ulonglong2 b = a;
        if (kVV) b = ld2(job.b + i);
        st2(job.dst + i, rtl_alu<OP>(a.x, b.x, s, q, iq), rtl_alu<OP>(a.y, b.y, s, q, iq));
    }
}

// dst = sum_t a[t] * b[t], products and sums exactly as the RTL's VFQMUL.vv / VFQADD.vv chain
template <int TERMS>
__global__ void __launch_bounds__(kThreads) mac_kernel(const MacJob *__restrict__ jobs, u32 n) {
    const MacJob &job = jobs[blockIdx.y];
    const u64 q = job.q, iq = job.iq;
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 i = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (i >= n) return;
        u64 acc0 = 0, acc1 = 0;
#pragma unroll
        for (int t = 0; t < TERMS; ++t) {
            const ulonglong2 a = ld2(job.a[t] + i), b = ld2(job.b[t] + i);
            const u64 p0 = rtl_alu<ALU_MUL_VV>(a.x, b.x, 0, q, iq);
            const u64 p1 = rtl_alu<ALU_MUL_VV>(a.y, b.y, 0, q, iq);
            if (t == 0) { acc0 = p0; acc1 = p1; }
            else { acc0 = rtl_add(acc0, p0, q); acc1 = rtl_add(acc1, p1, q); }  // both < q already
        }
        st2(job.dst + i, acc0, acc1);
    }
}

// Gather form of VAUT: destination d takes source i = d * k^-1 mod n; sign from (i*k) mod 2n,
// which equals d or d + n.  Writes are coalesced 16-byte stores; the strided 8-byte reads hit L2
// (one limb-polynomial is at most 512 KiB).
__device__ __forceinline__ u64 aut_gather(const u64 *__restrict__ src, u32 d, u32 n, u64 kinv, u64 k,
                                          u64 q) {
    const u32 i = (u32)((u64)d * kinv) & (n - 1);
    const u64 x = __ldg(src + i);
    const bool neg = (((u64)i * k) & (2ull * n - 1)) >= n;
    return neg ? q - x : x;
}

__global__ void __launch_bounds__(kThreads) vaut_kernel(const PermJob *__restrict__ jobs, u32 n) {
    const PermJob job = jobs[blockIdx.y];
    const u64 k = job.k;
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 d = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (d >= n) return;
        st2(job.dst + d, aut_gather(job.src, d, n, job.kinv, k, job.q),
            aut_gather(job.src, d + 1, n, job.kinv, k, job.q));
    }
}

// dst[d] = src[(d + rot) mod n]
__global__ void __launch_bounds__(kThreads) vroli_kernel(const PermJob *__restrict__ jobs, u32 n) {
    const PermJob job = jobs[blockIdx.y];
    const u32 rot = (u32)job.kinv;
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 d = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (d >= n) return;
        st2(job.dst + d, __ldg(job.src + ((d + rot) & (n - 1))),
            __ldg(job.src + ((d + 1 + rot) & (n - 1))));
    }
}

__global__ void __launch_bounds__(kThreads) copy_kernel(const CopyJob *__restrict__ jobs, u32 n) {
    const CopyJob job = jobs[blockIdx.y];
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 i = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (i >= n) return;
        const ulonglong2 v = ld2(job.src + i);
        st2(job.dst + i, v.x, v.y);
    }
}

// dst = c + aut_k(x) * p : VAUT (raw q - x), VFQMUL.vv, VFQADD.vv in one pass.
__global__ void __launch_bounds__(kThreads) autmac_kernel(const AutMacJob *__restrict__ jobs, u32 n) {
    const AutMacJob job = jobs[blockIdx.y];
    const u64 q = job.q, iq = job.iq, k = job.k;
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 d = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (d >= n) return;
        const u64 x0 = aut_gather(job.x, d, n, job.kinv, k, q);
        const u64 x1 = aut_gather(job.x, d + 1, n, job.kinv, k, q);
        const ulonglong2 p = ld2(job.p + d), c = ld2(job.c + d);
        const u64 m0 = rtl_alu<ALU_MUL_VV>(x0, p.x, 0, q, iq);
        const u64 m1 = rtl_alu<ALU_MUL_VV>(x1, p.y, 0, q, iq);
        st2(job.dst + d, rtl_alu<ALU_ADD_VV>(c.x, m0, 0, q, iq), rtl_alu<ALU_ADD_VV>(c.y, m1, 0, q, iq));
    }
}

// ---- VAUT through shared-memory tiles (aut_plan.hpp) -----------------------------------------------
// One CTA = one tile of the (point j, offset f) cover of Z_n.  Load phase: lanes run along j, so a warp
// reads runs of consecutive SOURCE words; store phase: lanes run along f, so a warp writes runs of
// consecutive DESTINATION words (and, fused, reads p and c along the same runs).  Both sides of the
// permutation move in 128-byte lines; the transposition is the padded shared-memory tile.
// Same per-element function as the gather kernels: raw q - x on the negated half ((i*k) mod 2n >= n).
constexpr int kAutIters = kAutTile / kAutThreads;   // slots per thread

// Index arithmetic is incremental (aut_plan.hpp AutLoadWalk / AutStoreWalk, the same code the CPU model in
// tests/native replays): per-CTA constant steps, aut_src / aut_dst evaluated once per thread.
template <bool MAC, class Job>
__device__ __forceinline__ void aut_tile_body(const Job &job, u32 n, u64 *tile, u32 tile_index) {
    const AutPlan &P = job.plan;
    const AutTile T = aut_tile(P, tile_index);
    const u64 *__restrict__ src;
    if constexpr (MAC) src = job.x; else src = job.src;
    u64 *__restrict__ dst = job.dst;
    const u64 q = job.q;
    const u32 k2 = (u32)job.k & (2 * n - 1);
    // ---- load: cp.async (LDGSTS) straight into the tile, lanes along the source
    auto load = [&](auto wide) {
        AutLoadWalk<decltype(wide)::value> w(P, T, threadIdx.x);
        if (w.idle()) return;
        const u32 tile_addr = (u32)__cvta_generic_to_shared(tile);
#pragma unroll
        for (int it = 0; it < kAutIters; ++it) {
            if (w.valid())
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(tile_addr + w.sm * 8), "l"(src + w.i) : "memory");
            w.next();
        }
    };
    if (T.log_jb > kAutThreadsLog) load(std::true_type{}); else load(std::false_type{});
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_all;" ::: "memory");
    __syncthreads();
    // ---- store: lanes along the destination
    auto store = [&](auto wide) {
        AutStoreWalk<decltype(wide)::value> w(P, T, threadIdx.x);
        if (w.idle()) return;
        [[maybe_unused]] const u64 *__restrict__ pp = nullptr, *__restrict__ cc = nullptr;
        [[maybe_unused]] u64 iq = 0;
        if constexpr (MAC) { pp = job.p; cc = job.c; iq = job.iq; }
#pragma unroll
        for (int it = 0; it < kAutIters; ++it) {
            if (w.valid()) {
                const u64 x = tile[w.sm];
                const u64 y = aut_negated(w.i, k2, n) ? q - x : x;
                if constexpr (MAC) {
                    const u64 m = rtl_alu<ALU_MUL_VV>(y, pp[w.d], 0, q, iq);
                    dst[w.d] = rtl_alu<ALU_ADD_VV>(cc[w.d], m, 0, q, iq);
                } else {
                    dst[w.d] = y;
                }
            }
            w.next();
        }
    };
    if (T.log_fb > kAutThreadsLog) store(std::true_type{}); else store(std::false_type{});
}

// The permuted side of the transfer reaches DRAM as 256-byte pieces scattered over the polynomial, which costs
// row-buffer locality (both aut kernels top out near 0.64 of the streaming copy rate).  So every CTA first asks
// L2 for a CONTIGUOUS slice of the source of a job kAutPrefetchJobs further down the launch: by the time that
// job's CTAs run, DRAM has delivered their polynomial in long bursts and their scattered runs hit L2.
#ifndef ALOHA_AUT_PREFETCH_JOBS
#define ALOHA_AUT_PREFETCH_JOBS 32
#endif
constexpr u32 kAutPrefetchJobs = ALOHA_AUT_PREFETCH_JOBS;

template <class Job>
__device__ __forceinline__ void aut_prefetch_ahead(const Job *__restrict__ jobs, u32 n, const u64 *Job::*field) {
    if (kAutPrefetchJobs == 0 || blockIdx.y + kAutPrefetchJobs >= gridDim.y) return;
    const Job &ahead = jobs[blockIdx.y + kAutPrefetchJobs];
    const u32 lines = n / 16, per = (lines + gridDim.x - 1) / gridDim.x, l0 = blockIdx.x * per;
    const u64 *base = ahead.*field;
    for (u32 l = l0 + threadIdx.x; l < l0 + per && l < lines; l += kAutThreads)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + 16 * (size_t)l));
}

// A CTA may walk kAutTilesPerCta consecutive tiles of its job one after the other (the fetch of the job record
// that every data load depends on is then paid once per CTA instead of once per tile).  Measured on B200 with
// 1, 2, 4, 8, 16: one tile per CTA is the fastest (0.65 of the HBM copy rate; 4: 0.53-0.65, 16: 0.41-0.65) --
// fewer, longer CTAs per polynomial lose more to imbalance than the head costs.  Kept as a build-time knob.
#ifndef ALOHA_AUT_TILES_PER_CTA
#define ALOHA_AUT_TILES_PER_CTA 1
#endif
constexpr u32 kAutTilesPerCta = ALOHA_AUT_TILES_PER_CTA;

template <bool MAC, class Job>
__device__ __forceinline__ void aut_cta(const Job &job, u32 n, u64 *tile) {
    const u32 first = blockIdx.x * kAutTilesPerCta, ntiles = job.plan.ntiles;
    for (u32 t = first; t < first + kAutTilesPerCta && t < ntiles; ++t) {
        if (t != first) __syncthreads();                 // the previous tile's store phase still reads the buffer
        aut_tile_body<MAC>(job, n, tile, t);
    }
}

__global__ void __launch_bounds__(kAutThreads) vaut_tiled_kernel(const AutJob *__restrict__ jobs, u32 n) {
    __shared__ u64 tile[kAutSmemWords];
    aut_prefetch_ahead(jobs, n, &AutJob::src);
    aut_cta<false>(jobs[blockIdx.y], n, tile);
}
__global__ void __launch_bounds__(kAutThreads) autmac_tiled_kernel(const AutMacJob *__restrict__ jobs, u32 n) {
    __shared__ u64 tile[kAutSmemWords];
    aut_cta<true>(jobs[blockIdx.y], n, tile);
}

// dst = c + a*b : product and sum exactly as VFQMUL.vv then VFQADD.vv would store them
__global__ void __launch_bounds__(kThreads) muladd_kernel(const MulAddJob *__restrict__ jobs, u32 n) {
    const MulAddJob job = jobs[blockIdx.y];
    const u64 q = job.q, iq = job.iq;
#pragma unroll
    for (int it = 0; it < kPerThread / kVec; ++it) {
        const u32 i = blockIdx.x * kPerBlock + it * kThreads * kVec + threadIdx.x * kVec;
        if (i >= n) return;
        const ulonglong2 a = ld2(job.a + i), b = ld2(job.b + i), c = ld2(job.c + i);
        const u64 m0 = rtl_alu<ALU_MUL_VV>(a.x, b.x, 0, q, iq);
        const u64 m1 = rtl_alu<ALU_MUL_VV>(a.y, b.y, 0, q, iq);
        st2(job.dst + i, rtl_alu<ALU_ADD_VV>(c.x, m0, 0, q, iq), rtl_alu<ALU_ADD_VV>(c.y, m1, 0, q, iq));
    }
}

// Sum of products over `terms` operand pairs, accumulated in instruction order with the RTL's
// multiply and add.  One 16-byte access per operand per thread; the pointer table is read through
// the read-only path and is identical for every thread of the block.
__global__ void __launch_bounds__(kThreads) sop_kernel(const SopJob *__restrict__ jobs, u32 n) {
    const SopJob job = jobs[blockIdx.y];
    const u64 q = job.q, iq = job.iq;
    const u32 i = (blockIdx.x * kThreads + threadIdx.x) * kVec;
    if (i >= n) return;
    const ulonglong2 *pairs = reinterpret_cast<const ulonglong2 *>(job.pairs);
    u64 acc0 = 0, acc1 = 0;
    constexpr u32 kU = 4;                       // operand pairs in flight per thread
    u32 t = 0;
    for (; t + kU <= job.terms; t += kU) {
        ulonglong2 a[kU], b[kU];
#pragma unroll
        for (u32 u = 0; u < kU; ++u) {
            const ulonglong2 pp = __ldg(pairs + t + u);
            a[u] = ld2(reinterpret_cast<const u64 *>(pp.x) + i);
            b[u] = ld2(reinterpret_cast<const u64 *>(pp.y) + i);
        }
#pragma unroll
        for (u32 u = 0; u < kU; ++u) {
            const u64 m0 = rtl_alu<ALU_MUL_VV>(a[u].x, b[u].x, 0, q, iq);
            const u64 m1 = rtl_alu<ALU_MUL_VV>(a[u].y, b[u].y, 0, q, iq);
            if (t + u == 0) { acc0 = m0; acc1 = m1; }
            else { acc0 = rtl_alu<ALU_ADD_VV>(acc0, m0, 0, q, iq); acc1 = rtl_alu<ALU_ADD_VV>(acc1, m1, 0, q, iq); }
        }
    }
    for (; t < job.terms; ++t) {
        const ulonglong2 pp = __ldg(pairs + t);
        const ulonglong2 a = ld2(reinterpret_cast<const u64 *>(pp.x) + i), b = ld2(reinterpret_cast<const u64 *>(pp.y) + i);
        const u64 m0 = rtl_alu<ALU_MUL_VV>(a.x, b.x, 0, q, iq);
        const u64 m1 = rtl_alu<ALU_MUL_VV>(a.y, b.y, 0, q, iq);
        if (t == 0) { acc0 = m0; acc1 = m1; }
        else { acc0 = rtl_alu<ALU_ADD_VV>(acc0, m0, 0, q, iq); acc1 = rtl_alu<ALU_ADD_VV>(acc1, m1, 0, q, iq); }
    }
    st2(job.dst + i, acc0, acc1);
}

// Fast basis extension: sum over source limbs of (x_t reduced into this modulus) * constant_t, then an
// optional VFQSUB.vs.  Reference semantics = the RTL's VCPY / VFQMOD, VFQMUL.vs and VFQADD.vv in instruction
// order (bext_rtl below, word for word).
//
// When iq is q's Barrett constant and every operand reaching a multiply is canonical, rtl_barrett(a, s) is
// exactly a * s mod q and the chain of RTL adds is exactly the modular sum, so ANY exact evaluation stores the
// same words.  The fast path uses one: lazy Shoup products (in [0, 2q), for any 64-bit operand) summed
// without reduction, one canonical reduction at the end.  An input word is in the fast domain when the RTL's
// own conditional subtracts would have made it canonical: x < 4q behind a VCPY (two subtracts there, one in
// the multiply), any word behind a VFQMOD, x < 2q with no pre-op.  Anything else (never produced by the
// generated streams) takes the RTL chain for that element.
__device__ __forceinline__ u64 bext_rtl(const BextJob &job, u32 i, u32 lane_word) {
    const u64 q = job.q, iq = job.iq;
    u64 acc = 0;
    for (u32 t = 0; t < job.nterms; ++t) {
        const BextTerm tm = job.terms[t];
        u64 a = tm.x[i + lane_word];
        if (tm.pre == PRE_VCPY) a = rtl_alu<ALU_ADD_VS>(a, 0, 0, q, iq);
        else if (tm.pre == PRE_VFQMOD) a = rtl_alu<ALU_MOD>(a, 0, 0, q, iq);
        const u64 m = rtl_alu<ALU_MUL_VS>(a, 0, tm.s, q, iq);
        acc = t == 0 ? m : rtl_alu<ALU_ADD_VV>(acc, m, 0, q, iq);
    }
    if (job.post == 1) acc = rtl_alu<ALU_SUB_VS>(acc, 0, job.post_s, q, iq);
    return acc;
}

__global__ void __launch_bounds__(kThreads) bext_kernel(const BextJob *__restrict__ jobs, u32 n) {
    const BextJob job = jobs[blockIdx.y];
    const u64 q = job.q;
    const u32 i = (blockIdx.x * kThreads + threadIdx.x) * kVec;
    if (i >= n) return;
    if (!job.fast) {
        st2(job.dst + i, bext_rtl(job, i, 0), bext_rtl(job, i, 1));
        return;
    }
    const u64 nq = 0 - q, q2 = 2 * q, q4 = 4 * q;
    u64 acc0 = 0, acc1 = 0;
    bool out_of_domain = false;
    constexpr u32 kU = 4;                       // operands in flight per thread
    auto term = [&](const BextTerm &tm, const ulonglong2 &v, u32 count) {
        u64 a0 = v.x, a1 = v.y;
        // VFQMOD is x mod q for every word, and the Shoup product below takes any 64-bit word: nothing to do for it
        if (tm.pre == PRE_VCPY) out_of_domain |= (a0 >= q4) | (a1 >= q4);
        else if (tm.pre != PRE_VFQMOD) out_of_domain |= (a0 >= q2) | (a1 >= q2);
        // every 7 lazy products (each below 2q) the running sums (then below 16q) are brought back below 2q
        if (count && count % 7 == 0) { acc0 = csub_s(csub_s(csub_s(acc0, 8 * q), q4), q2); acc1 = csub_s(csub_s(csub_s(acc1, 8 * q), q4), q2); }
        acc0 = shoup_mac(acc0, a0, tm.s, tm.sp, nq);
        acc1 = shoup_mac(acc1, a1, tm.s, tm.sp, nq);
    };
    u32 t = 0;
    for (; t + kU <= job.nterms; t += kU) {
        BextTerm tm[kU];
        ulonglong2 v[kU];
#pragma unroll
        for (u32 u = 0; u < kU; ++u) {
            tm[u] = job.terms[t + u];
            v[u] = ld2(tm[u].x + i);
        }
#pragma unroll
        for (u32 u = 0; u < kU; ++u) term(tm[u], v[u], t + u);
    }
    for (; t < job.nterms; ++t) {
        const BextTerm tm = job.terms[t];
        term(tm, ld2(tm.x + i), t);
    }
    if (out_of_domain) {                        // not taken by generated streams: RTL chain for these two words
        st2(job.dst + i, bext_rtl(job, i, 0), bext_rtl(job, i, 1));
        return;
    }
    // below 16q -> canonical
    acc0 = csub_s(csub_s(csub_s(csub_s(acc0, 8 * q), q4), q2), q);
    acc1 = csub_s(csub_s(csub_s(csub_s(acc1, 8 * q), q4), q2), q);
    if (job.post == 1) { acc0 = rtl_sub(acc0, job.post_s, q); acc1 = rtl_sub(acc1, job.post_s, q); }
    st2(job.dst + i, acc0, acc1);
}

// ALOHA_F_STRICT transforms: one launch per stage of the RTL's constant-geometry schedule
// (src/vp/ntt/ntt_fsm.sv:49-81; net effect per stage in SURVEY 3.3), every butterfly evaluated with the
// RTL ALU's CT / GS opcodes (modalu.sv:160-165, 296-327) -- so the result, and the ping-pong
// intermediate the RTL leaves in the source register, are word-exact for ANY input word.
template <bool INV>
__global__ void __launch_bounds__(kThreads) pease_kernel(const PeaseJob *__restrict__ jobs, u32 logn, u32 stage) {
    const PeaseJob job = jobs[blockIdx.y];
    const u32 p = blockIdx.x * kThreads + threadIdx.x, h = 1u << (logn - 1);
    if (p >= h) return;
    const u64 q = job.q, iq = job.iq;
    if (!INV) {
        const u32 m = 1u << stage;
        const u64 w = prered(job.tw[m + (p & (m - 1))].w, q);
        const u64 a = prered(job.src[p], q), b = prered(job.src[p + h], q);
        const u64 t = rtl_barrett(b, w, q, iq);
        st2(job.dst + 2 * p, rtl_add(a, t, q), rtl_sub(a, t, q));
    } else {
        const u32 m = 1u << (logn - 1 - stage);
        const u64 w = prered(job.tw[m + (p & (m - 1))].w, q);
        const ulonglong2 v = ld2(job.src + 2 * p);
        const u64 a = prered(v.x, q), b = prered(v.y, q);
        job.dst[p] = rtl_half(rtl_add(a, b, q), q);
        job.dst[p + h] = rtl_half(rtl_barrett(rtl_sub(a, b, q), w, q, iq), q);
    }
}

inline dim3 grid_for(u32 n, u32 njobs) { return dim3((n + kPerBlock - 1) / kPerBlock, njobs); }

}  // namespace

cudaError_t launch_ew(u32 op, const EwJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    const dim3 g = grid_for(n, njobs);
    switch (op) {
    case ALU_MUL_VV: ew_kernel<ALU_MUL_VV><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_MUL_VS: ew_kernel<ALU_MUL_VS><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_ADD_VV: ew_kernel<ALU_ADD_VV><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_ADD_VS: ew_kernel<ALU_ADD_VS><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_SUB_VV: ew_kernel<ALU_SUB_VV><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_SUB_VS: ew_kernel<ALU_SUB_VS><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_SUB_SV: ew_kernel<ALU_SUB_SV><<<g, kThreads, 0, st>>>(jobs, n); break;
    case ALU_MOD: ew_kernel<ALU_MOD><<<g, kThreads, 0, st>>>(jobs, n); break;
    default: return cudaErrorInvalidValue;
    }
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_vaut(const PermJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    vaut_kernel<<<grid_for(n, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_vaut_tiled(const AutJob *jobs, u32 njobs, u32 n, u32 max_tiles, cudaStream_t st) {
    vaut_tiled_kernel<<<dim3((max_tiles + kAutTilesPerCta - 1) / kAutTilesPerCta, njobs), kAutThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_autmac_tiled(const AutMacJob *jobs, u32 njobs, u32 n, u32 max_tiles, cudaStream_t st) {
    autmac_tiled_kernel<<<dim3((max_tiles + kAutTilesPerCta - 1) / kAutTilesPerCta, njobs), kAutThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_vroli(const PermJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    vroli_kernel<<<grid_for(n, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_pease(const PeaseJob *jobs, u32 njobs, u32 logn, u32 stage, bool inverse, cudaStream_t st) {
    const dim3 g(((1u << (logn - 1)) + kThreads - 1) / kThreads, njobs);
    if (inverse) pease_kernel<true><<<g, kThreads, 0, st>>>(jobs, logn, stage);
    else pease_kernel<false><<<g, kThreads, 0, st>>>(jobs, logn, stage);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_copy(const CopyJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    copy_kernel<<<grid_for(n, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_mac(const MacJob *jobs, u32 njobs, u32 terms, u32 n, cudaStream_t st) {
    const dim3 g = grid_for(n, njobs);
    switch (terms) {
    case 1: mac_kernel<1><<<g, kThreads, 0, st>>>(jobs, n); break;
    case 2: mac_kernel<2><<<g, kThreads, 0, st>>>(jobs, n); break;
    case 3: mac_kernel<3><<<g, kThreads, 0, st>>>(jobs, n); break;
    case 4: mac_kernel<4><<<g, kThreads, 0, st>>>(jobs, n); break;
    default: return cudaErrorInvalidValue;
    }
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_sop(const SopJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    sop_kernel<<<dim3((n / kVec + kThreads - 1) / kThreads, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_bext(const BextJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    bext_kernel<<<dim3((n / kVec + kThreads - 1) / kThreads, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_muladd(const MulAddJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    muladd_kernel<<<grid_for(n, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}
cudaError_t launch_autmac(const AutMacJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    autmac_kernel<<<grid_for(n, njobs), kThreads, 0, st>>>(jobs, n);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace alb
