// aut_plan.hpp -- tile decomposition of the Galois permutation d = i*k mod n (VAUT), shared by the host
// planner, the CUDA kernels (ew_kernels.cu) and the CPU test that replays the kernels' index arithmetic
// (tests/native/aut_plan_model.cpp).
//
// Replaces the address generation of src/vp/vxu/vxu_lane.sv:594-599 (dst address = (i*k) mod N, sign from
// (i*k) mod 2N) and the lane interconnect that carries it out 128 lanes per cycle
// (src/vp/iconn/iconn_top.sv:72-138).
//
// The permutation destroys locality at word granularity: neighbours in the source (i, i+1) land k apart,
// neighbours in the destination (d, d+1) come from sources k^-1 apart.  A tile that is contiguous on both
// sides is a parallelogram of the lattice {(j, f) -> d = j*k + f}:  +1 in j is +1 in the SOURCE, +1 in f is
// +1 in the DESTINATION.  (Boxes [0,J) x [0,F) never tile Z_n exactly for odd k, so the cover is built from
// the three-distance structure instead.)
//
//   Take the J points j*k mod n, j < J, on the circle Z_n.  For J = u + v, where u*k mod n = alpha is the
//   smallest positive residue among j < J and n - v*k mod n = beta the smallest negative one, the points cut
//   the circle into exactly two arc lengths: point j < v is followed by point j + u (arc alpha), point
//   j >= v by point j - v (arc beta);  v*alpha + u*beta = n.  Every destination d lies on exactly one arc:
//       d = (j*k + f) mod n,  0 <= f < arc(j)      <->      source i = (j + f*kinv) mod n.
//   So Z_n is covered exactly once by two rectangles of (j, f) pairs: class A = [0,v) x [0,alpha) and
//   class B = [v,J) x [0,beta).  All (u, alpha, v, beta) configurations come out of the subtractive
//   Euclidean algorithm on (k mod n, n); the planner walks them, replays one tile of each class and keeps
//   the configuration that touches the fewest 32-byte sectors per element on both sides.
//
//   A tile is JB consecutive j times FB consecutive f of one class (JB * FB <= 2048, chosen per class).
//   Loading it reads, for each f, the JB consecutive source words starting at j0 + f*kinv; storing it
//   writes, for each j, the FB consecutive destination words starting at j*k + f0.  The transposition
//   happens in shared memory.  Thin rectangles stay coalesced because the lattice is self-dual: when a class
//   has few points, consecutive f rows are (nearly) adjacent in the source, and vice versa.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ALOHA_HD __host__ __device__ __forceinline__
#else
#define ALOHA_HD inline
#endif

namespace alb {

#ifndef ALOHA_AUT_TILE_LOG
#define ALOHA_AUT_TILE_LOG 11
#endif
constexpr unsigned kAutTileLog = ALOHA_AUT_TILE_LOG;  // 2048 words = 16 KiB per tile
constexpr unsigned kAutTile = 1u << kAutTileLog;
constexpr unsigned kAutSmemWords = kAutTile + 1024;  // worst-case padding: FB rows * 1 word (FB <= 1024)
#ifndef ALOHA_AUT_THREADS
#define ALOHA_AUT_THREADS 256
#endif
constexpr unsigned kAutThreads = ALOHA_AUT_THREADS;    // threads per CTA of the tiled kernels (a thread's slots are this far apart)
constexpr unsigned kAutThreadsLog = kAutThreads == 128 ? 7 : kAutThreads == 256 ? 8 : kAutThreads == 512 ? 9 : 10;
static_assert((1u << kAutThreadsLog) == kAutThreads, "128, 256, 512 or 1024 threads");

struct AutClass {
    uint32_t j_begin, j_end;   // points of this class
    uint32_t gap;              // arc length
    uint32_t log_jb, log_fb;   // tile = 2^log_jb points x 2^log_fb offsets
    uint32_t fblocks;          // f-blocks per j-block
    uint32_t fblocks_magic;    // ceil(2^32 / fblocks): tile / fblocks = umulhi(tile, magic) for tile < 2^16
    uint32_t stride;           // shared-memory row stride in words (row = one f)
    uint32_t tile_begin;       // first tile id of this class
};

struct AutPlan {
    uint32_t mask;             // n - 1
    uint32_t kmod, kinv;       // k mod n, k^-1 mod n
    uint32_t ntiles;
    AutClass cls[2];           // A then B
};

struct AutTile {
    uint32_t j0, jcount, f0, fcount;
    uint32_t log_jb, log_fb, stride;
};

ALOHA_HD AutTile aut_tile(const AutPlan &p, uint32_t tile) {
    const AutClass &c = p.cls[tile >= p.cls[1].tile_begin ? 1 : 0];
    tile -= c.tile_begin;
    const uint32_t jb = c.fblocks == 1 ? tile : (uint32_t)(((uint64_t)tile * c.fblocks_magic) >> 32), fb = tile - jb * c.fblocks;
    AutTile t;
    t.log_jb = c.log_jb; t.log_fb = c.log_fb; t.stride = c.stride;
    t.j0 = c.j_begin + (jb << c.log_jb);
    t.f0 = fb << c.log_fb;
    const uint32_t jb_size = 1u << c.log_jb, fb_size = 1u << c.log_fb;
    t.jcount = c.j_end - t.j0 < jb_size ? c.j_end - t.j0 : jb_size;
    t.fcount = c.gap - t.f0 < fb_size ? c.gap - t.f0 : fb_size;
    return t;
}
// source / destination index of element (jl, fl) of a tile
ALOHA_HD uint32_t aut_src(const AutPlan &p, const AutTile &t, uint32_t jl, uint32_t fl) {
    return (t.j0 + jl + (t.f0 + fl) * p.kinv) & p.mask;
}
ALOHA_HD uint32_t aut_dst(const AutPlan &p, const AutTile &t, uint32_t jl, uint32_t fl) {
    return ((t.j0 + jl) * p.kmod + t.f0 + fl) & p.mask;
}
// which (jl, fl) a thread's slot s addresses in the load phase (lanes along j) / store phase (lanes along f)
ALOHA_HD void aut_load_slot(const AutTile &t, uint32_t s, uint32_t *jl, uint32_t *fl) {
    *fl = s >> t.log_jb; *jl = s & ((1u << t.log_jb) - 1);
}
ALOHA_HD void aut_store_slot(const AutTile &t, uint32_t s, uint32_t *jl, uint32_t *fl) {
    *jl = s >> t.log_fb; *fl = s & ((1u << t.log_fb) - 1);
}

// ---- how a thread walks its slots ----------------------------------------------------------------------
// A thread's slots are kAutThreads apart (slot = it * 256 + tid).  From one slot to the next either the fast
// coordinate advances by 256 (block side > 256: WIDE; it wraps into the slow one every side / 256 steps, for
// the whole CTA at once) or the slow coordinate advances by 256 / side.  Source index, destination index and
// shared-memory word therefore advance by per-CTA constants; aut_src / aut_dst are evaluated once per thread.
template <bool WIDE>
struct AutWalk {                 // coordinates (fast, slow) of a block whose fast side is 2^log_side
    uint32_t fast, slow, side, step_slow;
    ALOHA_HD AutWalk(uint32_t tid, uint32_t log_side)
        : fast(tid & ((1u << log_side) - 1)), slow(tid >> log_side), side(1u << log_side), step_slow(kAutThreads >> log_side) {}
    // advance by 256 slots; true when the fast coordinate wrapped (WIDE only; uniform over the CTA)
    ALOHA_HD bool next() {
        if (!WIDE) { slow += step_slow; return false; }
        fast += kAutThreads;
        if (fast >= side) { fast -= side; ++slow; return true; }
        return false;
    }
};

// load phase: fast = jl (lanes along the source), slow = fl
template <bool WIDE>
struct AutLoadWalk {
    AutWalk<WIDE> w;
    uint32_t i, sm;              // source index, shared-memory word
    uint32_t di, di_wrap, dsm, dsm_wrap, mask, jcount, fcount;
    ALOHA_HD AutLoadWalk(const AutPlan &p, const AutTile &t, uint32_t tid)
        : w(tid, t.log_jb), i(aut_src(p, t, w.fast, w.slow)), sm(w.slow * t.stride + w.fast),
          di(WIDE ? kAutThreads : w.step_slow * p.kinv), di_wrap(p.kinv - w.side),
          dsm(WIDE ? kAutThreads : w.step_slow * t.stride), dsm_wrap(t.stride - w.side), mask(p.mask),
          jcount(t.jcount), fcount(t.fcount) {}
    ALOHA_HD bool idle() const { return !WIDE && w.fast >= jcount; }     // narrow block: fast never changes
    ALOHA_HD bool valid() const { return (!WIDE || w.fast < jcount) && w.slow < fcount; }
    ALOHA_HD void next() {
        i += di; sm += dsm;
        if (w.next()) { i += di_wrap; sm += dsm_wrap; }
        i &= mask;
    }
};

// store phase: fast = fl (lanes along the destination), slow = jl
template <bool WIDE>
struct AutStoreWalk {
    AutWalk<WIDE> w;
    uint32_t i, d, sm;           // source index (for the sign), destination index, shared-memory word
    uint32_t di, di_wrap, dd, dd_wrap, dsm, dsm_wrap, mask, jcount, fcount;
    ALOHA_HD AutStoreWalk(const AutPlan &p, const AutTile &t, uint32_t tid)
        : w(tid, t.log_fb), i(aut_src(p, t, w.slow, w.fast)), d(aut_dst(p, t, w.slow, w.fast)), sm(w.fast * t.stride + w.slow),
          di(WIDE ? kAutThreads * p.kinv : w.step_slow), di_wrap(1u - w.side * p.kinv),
          dd(WIDE ? kAutThreads : w.step_slow * p.kmod), dd_wrap(p.kmod - w.side),
          dsm(WIDE ? kAutThreads * t.stride : w.step_slow), dsm_wrap(1u - w.side * t.stride), mask(p.mask),
          jcount(t.jcount), fcount(t.fcount) {}
    ALOHA_HD bool idle() const { return !WIDE && w.fast >= fcount; }
    ALOHA_HD bool valid() const { return w.slow < jcount && (!WIDE || w.fast < fcount); }
    ALOHA_HD void next() {
        i += di; d += dd; sm += dsm;
        if (w.next()) { i += di_wrap; d += dd_wrap; sm += dsm_wrap; }
        i &= mask; d &= mask;
    }
};
// (i * k) mod 2n >= n  <=>  bit log2(n) of i * k: 32-bit arithmetic is enough (the bit sits below bit 32)
ALOHA_HD bool aut_negated(uint32_t i, uint32_t k2, uint32_t n) { return ((i * k2) & n) != 0; }

namespace autdetail {

inline uint32_t clog2(uint64_t x) { uint32_t l = 0; while ((1ull << l) < x) ++l; return l; }

// shape a class's tiles as 2^log_jb points x 2^log_fb offsets (clamped to what the class has) and pad the rows
inline void shape_class(AutClass &c, uint32_t want_log_fb) {
    const uint32_t count = c.j_end - c.j_begin, lj = clog2(count ? count : 1), lf = clog2(c.gap ? c.gap : 1);
    uint32_t log_fb = want_log_fb < lf ? want_log_fb : lf;
    uint32_t log_jb = kAutTileLog - log_fb < lj ? kAutTileLog - log_fb : lj;
    if (log_jb + log_fb < kAutTileLog)                   // too few points: lengthen the f side
        log_fb = kAutTileLog - log_jb < lf ? kAutTileLog - log_jb : lf;
    if (log_fb > 10) log_fb = 10;                        // shared-memory padding budget (kAutSmemWords)
    c.log_jb = log_jb;
    c.log_fb = log_fb;
    const uint32_t JB = 1u << log_jb, FB = 1u << log_fb;
    c.fblocks = c.gap ? (c.gap + FB - 1) / FB : 0;
    c.fblocks_magic = c.fblocks > 1 ? (uint32_t)(((1ull << 32) + c.fblocks - 1) / c.fblocks) : 0;
    // row stride: the store phase reads element (jl, fl) with fl fastest; 16 consecutive lanes must fall on
    // 16 different 8-byte banks.  FB >= 16: an odd stride; 1 < FB < 16: stride = 16 / FB (mod 32 / FB).
    if (FB >= 16) c.stride = JB | 1;
    else if (FB == 1) c.stride = JB;
    else {
        const uint32_t m = 32 / FB, r = 16 / FB;
        c.stride = JB + (r + m - JB % m) % m;
    }
}

inline uint32_t class_tiles(const AutClass &c) {
    const uint32_t count = c.j_end - c.j_begin;
    return count && c.gap ? ((count + (1u << c.log_jb) - 1) >> c.log_jb) * c.fblocks : 0;
}

inline void number_tiles(AutPlan &p) {
    p.cls[0].tile_begin = 0;
    p.cls[1].tile_begin = class_tiles(p.cls[0]);
    p.ntiles = p.cls[1].tile_begin + class_tiles(p.cls[1]);
}

// What one tile costs, in SM cycles (rough, but it ranks shapes the way the kernels behave): every warp
// that has work in a phase issues about 25 instructions per slot step, and every 32-byte sector touched on
// either side is memory-system work.  Replays the kernel's own walks.
inline double tile_cost(const AutPlan &p, uint32_t tile, uint64_t *elements) {
    const AutTile t = aut_tile(p, tile);
    uint64_t sectors = 0, warp_steps = 0;
    const uint32_t slots = 1u << (t.log_jb + t.log_fb);
    for (int phase = 0; phase < 2; ++phase)
        for (uint32_t base = 0; base < slots; base += 32) {
            uint32_t sec[32], ns = 0;
            for (uint32_t lane = 0; lane < 32 && base + lane < slots; ++lane) {
                uint32_t jl, fl;
                if (phase == 0) aut_load_slot(t, base + lane, &jl, &fl); else aut_store_slot(t, base + lane, &jl, &fl);
                if (jl >= t.jcount || fl >= t.fcount) continue;
                const uint32_t a = (phase == 0 ? aut_src(p, t, jl, fl) : aut_dst(p, t, jl, fl)) >> 2;
                bool dup = false;
                for (uint32_t e = 0; e < ns && !dup; ++e) dup = sec[e] == a;
                if (!dup) sec[ns++] = a;
                if (phase == 0) ++*elements;
            }
            sectors += ns;
            warp_steps += ns ? 1 : 0;
        }
    return 12.0 * (double)warp_steps + 1.4 * (double)sectors;
}

// cost per element of a class with its current shape: a full tile and the last (partial) f-block, weighted
inline double class_cost(const AutPlan &p, int c) {
    const AutClass &C = p.cls[c];
    if (C.j_end == C.j_begin || !C.gap) return 0;
    uint64_t el_full = 0, el_last = 0;
    const double full = tile_cost(p, C.tile_begin, &el_full);
    if (C.fblocks == 1) return full / (double)el_full;
    const double last = tile_cost(p, C.tile_begin + C.fblocks - 1, &el_last);
    return (full * (C.fblocks - 1) + last) / ((double)el_full * (C.fblocks - 1) + (double)el_last);
}

// choose each class's tile shape (destination-run length 8 .. 1024 words) by that cost
inline double finish(AutPlan &p) {
    double total = 0;
    for (int c = 0; c < 2; ++c) {
        AutClass best = p.cls[c];
        double best_cost = -1;
        uint32_t last_fb = ~0u;
#ifndef ALOHA_AUT_MIN_LOG_FB
#define ALOHA_AUT_MIN_LOG_FB 3
#endif
        for (uint32_t want = ALOHA_AUT_MIN_LOG_FB; want <= 10; ++want) {
            shape_class(p.cls[c], want);
            if (p.cls[c].log_fb == last_fb) continue;          // clamped to the same shape as before
            last_fb = p.cls[c].log_fb;
            number_tiles(p);
            const double cost = class_cost(p, c);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = p.cls[c]; }
        }
        p.cls[c] = best;
        number_tiles(p);
        const double share = (double)(best.j_end - best.j_begin) * best.gap / ((double)p.mask + 1);
        total += share * (best_cost < 0 ? 0 : best_cost);
    }
    number_tiles(p);
    return total;
}

}  // namespace autdetail

// Host: choose the configuration and the tile shapes for (n, k).  k odd, n a power of two >= 4.
inline AutPlan make_aut_plan(uint32_t n, uint64_t k) {
    AutPlan base{};
    base.mask = n - 1;
    base.kmod = (uint32_t)(k & (n - 1));
    uint64_t inv = k;                                   // Newton: k^-1 mod 2^64
    for (int i = 0; i < 6; ++i) inv *= 2 - k * inv;
    base.kinv = (uint32_t)(inv & (n - 1));
    double built_cost = 0;
    auto build = [&](uint64_t u, uint64_t alpha, uint64_t v, uint64_t beta) {
        AutPlan p = base;
        p.cls[0].j_begin = 0; p.cls[0].j_end = (uint32_t)v; p.cls[0].gap = (uint32_t)alpha;
        p.cls[1].j_begin = (uint32_t)v; p.cls[1].j_end = (uint32_t)(u + v); p.cls[1].gap = (uint32_t)beta;
        built_cost = autdetail::finish(p);       // per-class tile shapes chosen by cost; expected cycles per element
        return p;
    };
    auto cost = [&](const AutPlan &) { return built_cost; };
    // Subtractive Euclid over (u, alpha), (v, beta).  Inside a run of equal steps the states change linearly,
    // so they are sampled geometrically (1, 2, 4, ... steps into the run, and its last two states).
    uint64_t u = 1, alpha = base.kmod, v = 1, beta = n - base.kmod;
    AutPlan best = build(u, alpha, v, beta);
    double best_cost = cost(best);
    auto consider = [&](uint64_t cu, uint64_t ca, uint64_t cv, uint64_t cb) {
        const AutPlan p = build(cu, ca, cv, cb);
        const double c = cost(p);
        if (c < best_cost) { best_cost = c; best = p; }
    };
    while (!(alpha == 1 && beta == 1)) {
        if (alpha > beta) {
            const uint64_t steps = (alpha - 1) / beta;           // alpha stays >= 1
            for (uint64_t t = 1;; t *= 2) {
                const uint64_t s = t < steps ? t : steps;
                consider(u + s * v, alpha - s * beta, v, beta);
                if (s == steps) break;
            }
            if (steps > 1) consider(u + (steps - 1) * v, alpha - (steps - 1) * beta, v, beta);
            u += steps * v; alpha -= steps * beta;
        } else {
            const uint64_t steps = (beta - 1) / alpha;
            for (uint64_t t = 1;; t *= 2) {
                const uint64_t s = t < steps ? t : steps;
                consider(u, alpha, v + s * u, beta - s * alpha);
                if (s == steps) break;
            }
            if (steps > 1) consider(u, alpha, v + (steps - 1) * u, beta - (steps - 1) * alpha);
            v += steps * u; beta -= steps * alpha;
        }
    }
    return best;
}

}  // namespace alb
