// engine.hpp -- host side of the engine: machine state, instruction-stream batcher, plan cache.
//
// Execution model.  aloha_run_vp(_batch) does not interpret instructions one by one on the GPU.
// The host *symbolically* executes the decoded stream (exactly the reference's architectural
// semantics: CSR context, bank-port operand routing, VLE/VSE addressing), producing a list of
// vector ops over device address ranges:
//   * every vector register is a *location* -- a renaming-pool buffer or an alias of an SPM / KSK
//     range.  VLE becomes an alias (no copy); VSE of a value produced in the same plan redirects
//     the producer to write SPM directly (store forwarding); anything that would overwrite an
//     aliased range first materialises the alias (copy-on-write), so architectural state is exact;
//   * ops are levelled by their true dependencies (RAW / WAW / WAR on address ranges); ops of one
//     level and kind -- typically the same instruction position of many RNS limbs, polynomials or
//     run_vp calls -- become ONE kernel launch with a device-resident job table;
//   * the resulting plan (launch list + job tables + exit state) is cached by (pc, CSR sets, entry
//     state), so a steady-state caller pays one hash lookup and a handful of launches.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <deque>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/aloha_b200.h"
#include "isa.hpp"
#include "kernels.cuh"

namespace alb {

constexpr u64 kLanes = 128;       // SYS_NUM_LANE  (vp_defines.vh:25) -- layout constant only
constexpr u64 kIramDepthDefault = 4096;  // IRAM_DEPTH    (vp_defines.vh:31)

enum Space : uint8_t { SP_UNDEF = 0, SP_POOL, SP_SPM, SP_KSK };

struct Loc {
    Space space = SP_UNDEF;
    u64 off = 0;   // word offset inside SPM / KSK, or pool buffer index
    u64 n = 0;     // valid words
};

enum OpKind : uint8_t { K_EW = 0, K_NTT, K_INTT, K_VAUT, K_VROLI, K_COPY, K_MULADD, K_AUTMAC, K_SOP, K_PEASE_F, K_PEASE_I, K_BEXT };

struct ExtTerm { const u64 *x; u64 s; u32 pre; };   // one summand of a fused base extension (kernels.cuh BextTerm)

struct VecOp {
    OpKind kind;
    u32 alu = 0;
    u32 n = 0;
    u64 *dst = nullptr;
    const u64 *a = nullptr, *b = nullptr, *c = nullptr;   // c: addend of the fused forms
    bool dead = false;                                     // removed by the fusion pass
    std::vector<std::pair<const u64 *, const u64 *>> terms;   // K_SOP: dst = sum_t a_t * b_t (in this order)
    std::vector<ExtTerm> ext;                              // K_BEXT: dst = sum_t pre_t(x_t) * s_t  [- post_s]
    u64 post_s = 0;
    u32 post = 0;
    u64 s = 0, q = 0, iq = 0, k = 0, kinv = 0;
    u32 pre = 0;                                           // K_NTT: base-extension op folded into the load
    int mod = -1;
    int level = 0;
};

struct Launch {
    OpKind kind;
    u32 alu, n, njobs;
    size_t table_off;   // byte offset of the job table inside the plan's device buffer
    u32 ngroups = 0;    // K_NTT: the first 16 * ngroups jobs are same-modulus runs of 16, described by the
    size_t group_off = 0;   //        NttRowGroup records at this offset (kernels.cuh launch_ntt_forward)
    u32 aux = 0;        // K_VAUT / K_AUTMAC (tiled): the largest tile count among the launch's jobs
};

struct TwTable {
    Tw *fwd = nullptr, *inv = nullptr;
    Tw *fwd_rows = nullptr, *inv_rows = nullptr;   // the same twiddles in the row passes' read order (kernels.cuh row_slot)
    ModulusConsts mc{};
};

struct Plan {
    std::vector<Launch> launches;
    void *d_tables = nullptr;
    size_t table_bytes = 0;
    // exit state
    u64 vl = 0, q = 0, iq = 0;
    int mod_idx = -1;
    Loc loc[32];
    std::vector<std::pair<u64, u64>> written;   // SPM word ranges stored to
    bool written_marked = false;                // the machine's written-flags already carry them (flags are only ever set)
    // entry conditions under which this plan may be replayed (everything else about the entry
    // state is irrelevant to it): registers it reads before writing must sit where they sat when it
    // was built; registers it overwrites may hold anything; the rest must not live in a pool buffer
    // the plan scribbles on, nor alias an SPM range it stores to.
    uint32_t live_in_mask = 0, killed_mask = 0;
    Loc live_in[32];
    std::vector<uint8_t> alloc_set;             // pool buffers written by the plan
    // accounting
    u64 instructions = 0, limb_ntts = 0, elided = 0, emitted = 0, kernel_launches = 0, fused = 0;
    cudaGraphExec_t graph = nullptr;
};

}  // namespace alb

struct aloha {
    aloha_cfg cfg{};
    alb::u64 nmax = 0;
    unsigned kbits = 0;
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t up_stream = nullptr, down_stream = nullptr;   // asynchronous DMA channels
    struct PendingDma { alb::u64 off, n; cudaEvent_t done; };
    std::vector<PendingDma> pending_down;                      // downloads not yet known complete
    std::vector<cudaEvent_t> event_pool;                       // recycled (timing-disabled) events
    std::vector<uint32_t> queued_pcs;                          // ALOHA_F_DEFER: run_vp calls not yet planned
    std::vector<aloha_vp_args> queued_args;
    alb::u64 *d_spm = nullptr, *d_ksk = nullptr, *d_pool = nullptr;
    alb::TmaMaps tma_maps{};      // swizzled tensor maps over d_spm / d_ksk / d_pool (inverse row pass)
    bool tma_maps_ok = false;
    alb::u64 spm_words = 0, ksk_words = 0;
    uint32_t pool_count = 0;
    alb::u64 iram_depth = alb::kIramDepthDefault;
    std::vector<uint8_t> isram, isram_valid;
    uint64_t isram_version = 0, tf_version = 0;
    std::vector<alb::u64> mod_q, mod_psi;
    std::map<std::pair<int, unsigned>, alb::TwTable> tw_tables;
    std::map<std::pair<uint32_t, uint64_t>, alb::AutPlan> aut_plans;   // (n, k) -> tile decomposition (aut_plan.hpp)
    // architectural state (persists across run_vp)
    alb::u64 vl = 0, q = 0, iq = 0;
    int mod_idx = -1;
    alb::Loc loc[32];
    std::vector<uint8_t> written;   // one flag per 64-byte beat of SPM
    std::unordered_map<std::string, std::vector<alb::Plan>> plans;   // key -> candidates (see Plan)
    aloha_stats stats{};
    std::string last_error;

    alb::u64 *ptr(const alb::Loc &l) const {
        switch (l.space) {
        case alb::SP_POOL: return d_pool + l.off * nmax;
        case alb::SP_SPM: return d_spm + l.off;
        case alb::SP_KSK: return d_ksk + l.off;
        default: return nullptr;
        }
    }
};
