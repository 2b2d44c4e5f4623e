// aut_plan_model.cpp -- CPU replay of the tiled VAUT kernels' index arithmetic (test infrastructure).
// Includes the SAME header the CUDA kernels use (aloha_b200/csrc/aut_plan.hpp) and walks the launch
// exactly as ew_kernels.cu does: grid.x = tile, 256 threads, 8 slots per thread, load phase with lanes
// along j, store phase with lanes along f.  Reports coverage and, per warp instruction, how many 32-byte
// sectors each side touches (the coalescing the design claims).
#include <cstdint>
#include <cstring>
#include <set>
#include <vector>

#include "../../aloha_b200/csrc/aut_plan.hpp"

using namespace alb;
typedef unsigned long long u64;

extern "C" {

// out: mask, kmod, kinv, ntiles, then per class: j_begin, j_end, gap, log_jb, log_fb, fblocks, stride, tile_begin
int aut_model_plan(uint32_t n, u64 k, uint32_t out[20]) {
    const AutPlan p = make_aut_plan(n, k);
    static_assert(sizeof(AutPlan) == 20 * sizeof(uint32_t), "flat layout");
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
        for (int phase = 0; phase < 2; ++phase)
            for (int it = 0; it < 8; ++it)
                for (uint32_t warp = 0; warp < 8; ++warp) {
                    std::set<uint32_t> sectors;
                    uint32_t bank[2][16] = {};
                    bool active = false;
                    for (uint32_t lane = 0; lane < 32; ++lane) {
                        const uint32_t s = it * 256 + warp * 32 + lane;
                        uint32_t jl, fl;
                        if (phase == 0) aut_load_slot(T, s, &jl, &fl); else aut_store_slot(T, s, &jl, &fl);
                        if (!(s < slots && jl < T.jcount && fl < T.fcount)) continue;
                        active = true;
                        const uint32_t i = aut_src(P, T, jl, fl), d = aut_dst(P, T, jl, fl), w = fl * T.stride + jl;
                        if (w >= kAutSmemWords) return -2;
                        if (w + 1 > stats[4]) stats[4] = w + 1;
                        ++bank[lane >> 4][w & 15];
                        if (phase == 0) {
                            smem[w] = src[i];
                            sectors.insert(i >> 2);
                        } else {
                            const bool neg = (((u64)i * k2) & (2ull * n - 1)) >= n;
                            dst[d] = neg ? q - smem[w] : smem[w];
                            if (seen[d]) ++stats[1];
                            seen[d] = 1;
                            ++stats[0];
                            sectors.insert(d >> 2);
                        }
                    }
                    if (!active) continue;
                    ++stats[7];
                    stats[2 + phase] += sectors.size();
                    for (int h = 0; h < 2; ++h)
                        for (int b = 0; b < 16; ++b)
                            if (bank[h][b] > stats[5 + phase]) stats[5 + phase] = bank[h][b];
                }
    }
    return 0;
}
}
