// aut_plan_model.cpp -- CPU replay of the tiled VAUT kernels' index arithmetic (test infrastructure).
// Includes the SAME header the CUDA kernels use (aloha_b200/csrc/aut_plan.hpp) and walks the launch
// exactly as ew_kernels.cu does: grid.x = tile, 256 threads, 8 slots per thread, load phase with lanes
// along j, store phase with lanes along f.  Reports coverage and, per warp instruction, how many 32-byte
// sectors each side touches (the coalescing the design claims).
#include <cstdint>
#include <cstring>
#include <set>
#include <type_traits>
#include <vector>

#include "../../aloha_b200/csrc/aut_plan.hpp"

using namespace alb;
typedef unsigned long long u64;

extern "C" {

// out: mask, kmod, kinv, ntiles, then per class: j_begin, j_end, gap, log_jb, log_fb, fblocks, stride, tile_begin
int aut_model_plan(uint32_t n, u64 k, uint32_t out[22]) {
    const AutPlan p = make_aut_plan(n, k);
    static_assert(sizeof(AutPlan) == 22 * sizeof(uint32_t), "flat layout");
    std::memcpy(out, &p, sizeof p);
    return 0;
}

// dst = aut_k(src) through the tile walk.  stats: [0] elements written, [1] destinations written twice,
// [2] source sectors touched summed over warp load instructions, [3] destination sectors summed over warp
// store instructions, [4] max shared-memory words used, [5] worst 8-byte bank conflict degree in the load
// phase's shared stores, [6] the same for the store phase's shared loads, [7] active warp instructions
int aut_model_apply(uint32_t n, u64 k, u64 q, const u64 *src, u64 *dst, u64 stats[8]) {
    const AutPlan P = make_aut_plan(n, k);
    std::vector<uint8_t> seen(n, 0);
    std::memset(stats, 0, 8 * sizeof(u64));
    const u64 k2 = k & (2ull * n - 1);
    for (uint32_t tile = 0; tile < P.ntiles; ++tile) {
        const AutTile T = aut_tile(P, tile);
        const uint32_t slots = 1u << (T.log_jb + T.log_fb);
        if (slots > kAutTile) return -1;
        std::vector<u64> smem(kAutSmemWords, ~0ull);
        // per warp instruction: the lanes' accesses.  phase 0 = load (AutLoadWalk), 1 = store (AutStoreWalk),
        // each thread walking its 8 slots exactly as the kernel does (incremental indices).
        struct Access { bool valid; uint32_t i, d, sm; };
        for (int phase = 0; phase < 2; ++phase) {
            std::vector<std::vector<Access>> acc(kAutThreads, std::vector<Access>(kAutTile / kAutThreads));
            for (uint32_t tid = 0; tid < kAutThreads; ++tid) {
                auto walk = [&](auto w) {
                    for (int it = 0; it < (int)(kAutTile / kAutThreads); ++it) {
                        Access a{!w.idle() && w.valid(), w.i, 0, w.sm};
                        if constexpr (std::is_same_v<decltype(w), AutStoreWalk<true>> || std::is_same_v<decltype(w), AutStoreWalk<false>>) a.d = w.d;
                        acc[tid][it] = a;
                        w.next();
                    }
                };
                if (phase == 0) { if (T.log_jb > kAutThreadsLog) walk(AutLoadWalk<true>(P, T, tid)); else walk(AutLoadWalk<false>(P, T, tid)); }
                else { if (T.log_fb > kAutThreadsLog) walk(AutStoreWalk<true>(P, T, tid)); else walk(AutStoreWalk<false>(P, T, tid)); }
            }
            for (int it = 0; it < (int)(kAutTile / kAutThreads); ++it)
                for (uint32_t warp = 0; warp < kAutThreads / 32; ++warp) {
                    std::set<uint32_t> sectors;
                    uint32_t bank[2][16] = {};
                    bool active = false;
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const Access &a = acc[warp * 32 + lane][it];
                        if (!a.valid) continue;
                        active = true;
                        if (a.sm >= kAutSmemWords) return -2;
                        if (a.sm + 1 > stats[4]) stats[4] = a.sm + 1;
                        ++bank[lane >> 4][a.sm & 15];
                        if (phase == 0) {
                            smem[a.sm] = src[a.i];
                            sectors.insert(a.i >> 2);
                        } else {
                            dst[a.d] = aut_negated(a.i, (uint32_t)k2, n) ? q - smem[a.sm] : smem[a.sm];
                            if (seen[a.d]) ++stats[1];
                            seen[a.d] = 1;
                            ++stats[0];
                            sectors.insert(a.d >> 2);
                        }
                    }
                    if (!active) continue;
                    ++stats[7];
                    stats[2 + phase] += sectors.size();
                    for (int h = 0; h < 2; ++h)
                        for (int bnk = 0; bnk < 16; ++bnk)
                            if (bank[h][bnk] > stats[5 + phase]) stats[5 + phase] = bank[h][bnk];
                }
        }
    }
    return 0;
}
}
