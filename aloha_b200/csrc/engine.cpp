// engine.cpp -- machine state, batcher (symbolic execution -> levelled launch plan), executor, C-ABI.
// See engine.hpp for the execution model and include/aloha_b200.h for the boundary.
#include "engine.hpp"

#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <cstring>

using namespace alb;
typedef unsigned __int128 u128;

namespace {

#define CU(call)                                                                        \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            E->last_error = std::string(#call) + ": " + cudaGetErrorString(e_);         \
            return ALOHA_E_CUDA;                                                        \
        }                                                                               \
    } while (0)

// Every entry point runs with the engine's device current and puts the caller's device back on exit: a
// caller may drive several engines (one per GPU) from one thread, or switch devices between calls.
struct DeviceGuard {
    int prev = -1, dev;
    explicit DeviceGuard(int d) : dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

int fail(aloha *E, int code, const std::string &msg) {
    E->last_error = msg;
    return code;
}

unsigned ilog2(u64 x) { unsigned l = 0; while ((1ull << l) < x) ++l; return l; }

u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = (u64)((u128)r * a % q);
        a = (u64)((u128)a * a % q);
        e >>= 1;
    }
    return r;
}
u64 bitrev(u64 x, unsigned bits) {
    u64 r = 0;
    for (unsigned i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
// q = 2^60 - d with d <= 2^27: the transforms use the split-product arithmetic (modarith.cuh mul_pm)
u32 modulus_form(const aloha *E, u64 q) {
    if (E->cfg.flags & ALOHA_F_GENERIC_MODMUL) return FORM_GENERIC;
    return (q < (1ull << 60) && (1ull << 60) - q <= (1ull << 27)) ? FORM_PM : FORM_GENERIC;
}
// twiddle pair in the form's layout (kernels.cuh Tw)
Tw twiddle(u64 w, u64 q, u32 form) {
    Tw t;
    t.w = w;
    t.wp = form == FORM_PM ? (u64)(((u128)w << 32) % q) : (u64)(((u128)w << 64) / q);
    return t;
}

// Which tensor map covers p, and p's offset inside that buffer in 128-byte lines (kernels.cuh TmaMaps).
bool tensor_coords(const aloha *E, const u64 *p, u32 *map, u32 *line) {
    const struct { const u64 *base; u64 words; } bufs[3] = {
        {E->d_spm, E->spm_words}, {E->d_ksk, E->ksk_words}, {E->d_pool, (u64)E->pool_count * E->nmax}};
    for (u32 i = 0; i < 3; ++i)
        if (bufs[i].base && p >= bufs[i].base && p < bufs[i].base + bufs[i].words) {
            const u64 off = (u64)(p - bufs[i].base);
            if (off % 16) return false;
            *map = i;
            *line = (u32)(off / 16);
            return true;
        }
    return false;
}

// Swizzled tensor maps over the three buffers.  The encoder lives in the driver; without it (or if it
// refuses) the inverse transforms simply keep their plain row pass.
void build_tensor_maps(aloha *E) {
    typedef CUresult (*Encode)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn ||
        q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return;
    }
    const struct { u64 *base; u64 words; } bufs[3] = {
        {E->d_spm, E->spm_words}, {E->d_ksk, E->ksk_words}, {E->d_pool, (u64)E->pool_count * E->nmax}};
    for (int i = 0; i < 3; ++i) {
        u64 *base = bufs[i].base ? bufs[i].base : E->d_spm;            // an absent buffer is never selected
        const u64 words = bufs[i].base ? bufs[i].words : E->spm_words;
        const cuuint64_t dims[2] = {16, words / 16};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {16, 16}, estr[2] = {1, 1};
        if (words < 256 || dims[1] > (1ull << 32) ||
            ((Encode)fn)(&E->tma_maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return;
    }
    E->tma_maps_ok = true;
}

// ------------------------------------------------------------------------------ twiddle tables
// Index j holds root^bitrev(j, logN) -- the reference ROM's order
// (sim/vp/tf_rom_generator/tf_rom_generator.sv:28-30,61-63,111,147-148).
int get_tables(aloha *E, int mod, unsigned logn, const TwTable **out) {
    auto key = std::make_pair(mod, logn);
    auto it = E->tw_tables.find(key);
    if (it != E->tw_tables.end()) { *out = &it->second; return ALOHA_OK; }
    const u64 q = E->mod_q[mod], n = 1ull << logn;
    if (q <= (1ull << 59) || q >= (1ull << 60))
        return fail(E, ALOHA_E_STATE, "transform modulus must be a 60-bit prime (mod_width = 60, vxu_lane.sv:539)");
    if ((q - 1) % (2 * n)) return fail(E, ALOHA_E_STATE, "2N does not divide q-1");
    const u64 psi = powmod(E->mod_psi[mod], E->nmax / n, q);
    if (powmod(psi, n, q) != q - 1) return fail(E, ALOHA_E_STATE, "psi is not a primitive 2N-th root of unity");
    const u64 ipsi = powmod(psi, q - 2, q);
    const u32 form = modulus_form(E, q);
    std::vector<Tw> fwd(n), inv(n);
    u64 cf = 1, ci = 1;
    for (u64 e = 0; e < n; ++e) {
        const u64 j = bitrev(e, logn);
        fwd[j] = twiddle(cf, q, form);
        inv[j] = twiddle(ci, q, form);
        cf = (u64)((u128)cf * psi % q);
        ci = (u64)((u128)ci * ipsi % q);
    }
    // the forward row pass's view: row r of the N/256 x 256 layout uses tw[((R + r) << u) + j], u < 8
    const u64 R = n / 256;
    // inverse row pass: GS level lt < 8 of row r uses itw[2^(logN-1-lt) + (r << (7-lt)) + j], j < 2^(7-lt)
    std::vector<Tw> fwd_rows(n, Tw{0, 0}), inv_rows(n, Tw{0, 0});
    for (u64 r = 0; r < R; ++r)
        for (u32 u = 0; u < 8; ++u)
            for (u32 j = 0; j < (1u << u); ++j) {
                fwd_rows[r * 256 + row_slot8(u, j)] = fwd[((R + r) << u) + j];
                inv_rows[r * 256 + row_slot(u, j)] = inv[(1ull << (logn - 8 + u)) + (r << u) + j];
            }
    TwTable t;
    const struct { Tw **dev; const std::vector<Tw> *host; } parts[4] = {
        {&t.fwd, &fwd}, {&t.inv, &inv}, {&t.fwd_rows, &fwd_rows}, {&t.inv_rows, &inv_rows}};
    for (auto &pt : parts) {
        cudaError_t e = cudaMalloc(pt.dev, n * sizeof(Tw));
        if (e == cudaSuccess) e = cudaMemcpy(*pt.dev, pt.host->data(), n * sizeof(Tw), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {                 // nothing half-built stays behind
            cudaFree(t.fwd); cudaFree(t.inv); cudaFree(t.fwd_rows); cudaFree(t.inv_rows);
            return fail(E, ALOHA_E_CUDA, std::string("twiddle tables: ") + cudaGetErrorString(e));
        }
    }
    const u64 ninv = powmod(n % q, q - 2, q);
    const Tw a = twiddle(ninv, q, form), b = twiddle((u64)((u128)inv[1].w * ninv % q), q, form);
    t.mc.q = q;
    t.mc.ninv = a.w; t.mc.ninv_p = a.wp;
    t.mc.wninv = b.w; t.mc.wninv_p = b.wp;
    t.mc.mest = (u32)((((u128)1) << 91) / q);
    t.mc.pre = 0;
    t.mc.form = form;
    t.mc.d = form == FORM_PM ? (u32)((1ull << 60) - q) : 0;
    t.mc.q3 = 3 * q;
    *out = &(E->tw_tables[key] = t);
    return ALOHA_OK;
}

void free_tables(aloha *E) {
    for (auto &kv : E->tw_tables) { cudaFree(kv.second.fwd); cudaFree(kv.second.inv); cudaFree(kv.second.fwd_rows); cudaFree(kv.second.inv_rows); }
    E->tw_tables.clear();
}
void free_plans(aloha *E) {
    for (auto &kv : E->plans)
        for (auto &pl : kv.second) {
            if (pl.graph) cudaGraphExecDestroy(pl.graph);
            cudaFree(pl.d_tables);
        }
    E->plans.clear();
}

// Tile decomposition of the Galois permutation for (n, k), built once per pair.
const AutPlan &aut_plan_for(aloha *E, u32 n, u64 k) {
    auto key = std::make_pair((uint32_t)n, (uint64_t)k);
    auto it = E->aut_plans.find(key);
    if (it == E->aut_plans.end()) it = E->aut_plans.emplace(key, make_aut_plan(n, k)).first;
    return it->second;
}

inline bool overlap(const u64 *a, u64 an, const u64 *b, u64 bn) {
    return a && b && a < b + bn && b < a + an;
}

// ------------------------------------------------------------------------------ batcher
struct Builder {
    aloha *E;
    std::vector<VecOp> ops;
    Loc loc[32];
    u64 vl, q, iq;
    int mod_idx;
    std::deque<u32> free_bufs;
    std::vector<int> producer;          // pool buffer -> index of the op that wrote it in this plan
    std::vector<std::pair<u64, u64>> written;
    u64 instructions = 0, limb_ntts = 0, elided = 0, emitted = 0;
    uint32_t live_in_mask = 0, killed_mask = 0;
    Loc live_in[32];
    std::vector<uint8_t> alloc_set;

    // the plan's behaviour depends on where `reg` sat on entry
    void note_read(int reg) {
        if (reg < 0 || ((killed_mask | live_in_mask) >> reg) & 1) return;
        live_in_mask |= 1u << reg;
        live_in[reg] = E->loc[reg];
    }

    explicit Builder(aloha *e) : E(e), vl(e->vl), q(e->q), iq(e->iq), mod_idx(e->mod_idx) {
        std::vector<uint8_t> used(E->pool_count, 0);
        for (int r = 0; r < 32; ++r) {
            loc[r] = E->loc[r];
            if (loc[r].space == SP_POOL) used[loc[r].off] = 1;
        }
        for (u32 b = 0; b < E->pool_count; ++b) if (!used[b]) free_bufs.push_back(b);
        producer.assign(E->pool_count, -1);
        alloc_set.assign(E->pool_count, 0);
    }

    u64 *ptr(const Loc &l) const { return E->ptr(l); }

    int alloc(Loc *out, u64 n) {
        if (free_bufs.empty()) return fail(E, ALOHA_E_NOMEM, "renaming pool exhausted");
        out->space = SP_POOL;
        out->off = free_bufs.front();
        out->n = n;
        free_bufs.pop_front();
        producer[out->off] = -1;
        alloc_set[out->off] = 1;
        return ALOHA_OK;
    }
    void release(const Loc &l) {
        if (l.space == SP_POOL) { free_bufs.push_back((u32)l.off); producer[l.off] = -1; }
    }
    // vd gets a fresh value living at `nl`
    void define(int vd, const Loc &nl) {
        release(loc[vd]);
        loc[vd] = nl;
        killed_mask |= 1u << vd;
    }
    void emit_copy(u64 *dst, const u64 *src, u64 n) {
        VecOp o{};
        o.kind = K_COPY;
        o.n = (u32)n;
        o.dst = dst;
        o.a = src;
        ops.push_back(o);
        ++emitted;
    }
    // Something is about to overwrite words [off, off+n) of `space`: registers aliasing that range
    // keep their value by moving to a pool buffer first (copy-on-write).
    int cow(Space space, u64 off, u64 n, int except_reg, bool *moved = nullptr) {
        for (int r = 0; r < 32; ++r) {
            if (r == except_reg || loc[r].space != space) continue;
            if (!(loc[r].off < off + n && off < loc[r].off + loc[r].n)) continue;
            note_read(r);
            Loc nl;
            int rc = alloc(&nl, loc[r].n);
            if (rc) return rc;
            emit_copy(ptr(nl), ptr(loc[r]), loc[r].n);
            producer[nl.off] = (int)ops.size() - 1;
            loc[r] = nl;
            if (moved) *moved = true;
        }
        return ALOHA_OK;
    }

    int read_loc(int reg, u64 n, const u64 **out) {
        if (reg < 0) return fail(E, ALOHA_E_OPCODE, "operand port disabled");
        note_read(reg);
        const Loc &l = loc[reg];
        if (l.space == SP_UNDEF)
            return fail(E, ALOHA_E_UNDEFINED,
                        "v" + std::to_string(reg) + " read while undefined (never written, or clobbered by VNTT/VINTT)");
        if (l.n < n) return fail(E, ALOHA_E_UNDEFINED, "v" + std::to_string(reg) + " holds fewer words than vl");
        *out = ptr(l);
        return ALOHA_OK;
    }

    // vp_top_full.sv:105-117 + addr_gen.v:44
    int resolve_mem(u64 imm, bool is_load, const aloha_vp_args &a, u64 rows, Space *space, u64 *word_off) {
        const unsigned sel = (unsigned)(imm >> 48);
        const u64 off = (imm >> 10) & 0xffff;
        if (off + rows > 0x10000) return fail(E, ALOHA_E_RANGE, "VLE/VSE row offset wraps the 16-bit row field");
        const bool ksk = is_load && sel == 15;
        const u64 base = ksk ? a.ksk_ptr : sel == 0 ? a.src0 : sel == 1 ? a.src1 : sel == 2 ? a.rslt : 0;
        const u64 limit = ksk ? E->cfg.ksk_rows : E->cfg.spm_rows;
        if (base + off + rows > limit) return fail(E, ALOHA_E_RANGE, "VLE/VSE rows outside SPM/KSK memory");
        *space = ksk ? SP_KSK : SP_SPM;
        *word_off = (base + off) * kLanes;
        return ALOHA_OK;
    }

    int store(int vs, u64 word_off, u64 n) {
        if (vs < 0) return fail(E, ALOHA_E_OPCODE, "VSE source port disabled");
        note_read(vs);
        if (loc[vs].space == SP_UNDEF) return fail(E, ALOHA_E_UNDEFINED, "VSE of undefined v" + std::to_string(vs));
        if (loc[vs].n < n) return fail(E, ALOHA_E_UNDEFINED, "VSE of a register shorter than vl");
        u64 *M = E->d_spm + word_off;
        written.emplace_back(word_off, n);
        if (loc[vs].space == SP_SPM && loc[vs].off == word_off) { ++elided; return ALOHA_OK; }  // already there
        // the source itself may alias a partially overlapping SPM range: move it out first
        if (loc[vs].space == SP_SPM && loc[vs].off < word_off + n && word_off < loc[vs].off + loc[vs].n) {
            Loc nl;
            int rc = alloc(&nl, loc[vs].n);
            if (rc) return rc;
            emit_copy(ptr(nl), ptr(loc[vs]), loc[vs].n);
            producer[nl.off] = (int)ops.size() - 1;
            loc[vs] = nl;
        }
        // copy-on-write copies are appended AFTER the producer in program order, so a store that
        // needed one cannot be forwarded back into the producer
        bool moved = false;
        int rc = cow(SP_SPM, word_off, n, vs, &moved);
        if (rc) return rc;
        const bool alias_ok = !(E->cfg.flags & ALOHA_F_NO_ALIAS) && !moved;
        if (alias_ok && loc[vs].space == SP_POOL && loc[vs].n == n && producer[loc[vs].off] >= 0) {
            const int p = producer[loc[vs].off];
            const u64 *X = ptr(loc[vs]);
            VecOp &po = ops[p];
            bool ok = true;
            // the producer may read M only as an exact in-place operand of an index-preserving kernel
            const bool inplace_safe = po.kind == K_EW || po.kind == K_NTT || po.kind == K_INTT || po.kind == K_COPY;
            for (const u64 *s : {po.a, po.b})
                if (overlap(s, po.n, M, n) && !(inplace_safe && s == M)) ok = false;
            // nothing between the producer and here may touch M
            for (size_t j = p + 1; ok && j < ops.size(); ++j) {
                const VecOp &o = ops[j];
                if (overlap(o.dst, o.n, M, n) || overlap(o.a, o.n, M, n) || overlap(o.b, o.n, M, n)) ok = false;
            }
            if (ok) {
                for (size_t j = p + 1; j < ops.size(); ++j) {
                    if (ops[j].a == X) ops[j].a = M;
                    if (ops[j].b == X) ops[j].b = M;
                }
                po.dst = M;
                Loc nl;
                nl.space = SP_SPM;
                nl.off = word_off;
                nl.n = n;
                define(vs, nl);
                ++elided;
                return ALOHA_OK;
            }
        }
        emit_copy(M, ptr(loc[vs]), n);
        return ALOHA_OK;
    }

    // ALOHA_F_STRICT: the RTL's schedule, one op per stage, ping-ponging between the source register's
    // buffer A and the destination's buffer B (ntt_fsm.sv:80).  With an odd number of stages the
    // result lands in B and A keeps stage logN-2 (exactly the RTL); with an even number the
    // RTL would leave them the other way round -- like the oracle, vd is defined to hold the result.
    int strict_transform(const VecOp &proto, int src_reg, int wd, u64 n) {
        const unsigned logn = ilog2(n);
        Loc A;
        int rc = alloc(&A, n);                 // private, writable copy of vs1's value
        if (rc) return rc;
        emit_copy(ptr(A), proto.a, n);
        producer[A.off] = (int)ops.size() - 1;
        Loc Bl;
        rc = alloc(&Bl, n);
        if (rc) return rc;
        u64 *bufs[2] = {ptr(A), ptr(Bl)};
        for (unsigned s = 0; s < logn; ++s) {
            VecOp o = proto;
            o.kind = proto.kind == K_NTT ? K_PEASE_F : K_PEASE_I;
            o.alu = s;                          // stage number (also keeps stages in separate launches)
            o.a = bufs[s & 1];
            o.dst = bufs[(s + 1) & 1];
            ops.push_back(o);
        }
        const bool result_in_B = logn & 1;
        // the last stage wrote the result buffer, the one before it the other buffer
        producer[(result_in_B ? Bl : A).off] = (int)ops.size() - 1;
        producer[(result_in_B ? A : Bl).off] = (int)ops.size() - 2;
        define(wd, result_in_B ? Bl : A);
        define(src_reg, result_in_B ? A : Bl);
        return ALOHA_OK;
    }

    int step(const Inst &in, const aloha_vp_args &a, bool *brk) {
        ++instructions;
        const MicroOp m = expand(in, a.step);
        *brk = in.funct6 == F6_BREAK;
        if (m.cfg == 1) {  // seq_top.v:417-429
            const u64 n = m.scalar_cfg / 64;
            if (m.scalar_cfg % 64 || n < 256 || (n & (n - 1)) || n > E->nmax)
                return fail(E, ALOHA_E_STATE, "VSETVL: vl/64 must be a power of two in [256, vlmax/64]");
            vl = m.scalar_cfg;
            return ALOHA_OK;
        }
        if (m.cfg == 2) {  // vxu_top.sv:112-118
            q = m.scalar_cfg;
            mod_idx = -1;
            for (size_t i = 0; i < E->mod_q.size(); ++i)
                if (E->mod_q[i] == q) { mod_idx = (int)i; break; }
            return ALOHA_OK;
        }
        if (m.cfg == 3) { iq = m.scalar_cfg; return ALOHA_OK; }
        switch (in.funct6) {
        case F6_BREAK: case F6_NOP: return ALOHA_OK;
        case F6_FQMUL: case F6_FQADD: case F6_FQSUB: case F6_FQMOD: case F6_VCPY: case F6_VAUT:
        case F6_VROLI: case F6_NTT: case F6_INTT: case F6_VLE: case F6_VSE: break;
        default: return fail(E, ALOHA_E_OPCODE, "unknown funct6 " + std::to_string(in.funct6));
        }
        if ((in.funct6 == F6_FQMUL || in.funct6 == F6_FQADD) && in.funct3 > 1)
            return fail(E, ALOHA_E_OPCODE, "unsupported funct3");
        if (in.funct6 == F6_FQSUB && in.funct3 > 2) return fail(E, ALOHA_E_OPCODE, "unsupported funct3");
        if (!vl) return fail(E, ALOHA_E_STATE, "vector op before VSETVL");
        const u64 n = vl / 64, rows = n / kLanes;
        const int r0 = port_reg(m.b0r, 0), r1 = port_reg(m.b1r, 1);
        const int w0 = port_reg(m.b0w, 0), w1 = port_reg(m.b1w, 1);
        const int wd = w0 >= 0 ? w0 : w1;
        auto rd = [&](unsigned bit) { return ((m.muxo >> bit) & 1) ? r1 : r0; };

        if (m.ls == 1) {
            Space sp;
            u64 off;
            int rc = resolve_mem(m.scalar_ls, true, a, rows, &sp, &off);
            if (rc) return rc;
            if (wd < 0) return fail(E, ALOHA_E_OPCODE, "VLE without destination");
            Loc nl;
            if (E->cfg.flags & ALOHA_F_NO_ALIAS) {
                rc = alloc(&nl, n);
                if (rc) return rc;
                emit_copy(ptr(nl), (sp == SP_KSK ? E->d_ksk : E->d_spm) + off, n);
                producer[nl.off] = (int)ops.size() - 1;
            } else {
                nl.space = sp;
                nl.off = off;
                nl.n = n;
                ++elided;
            }
            define(wd, nl);
            return ALOHA_OK;
        }
        if (m.ls == 2) {
            Space sp;
            u64 off;
            int rc = resolve_mem(m.scalar_ls, false, a, rows, &sp, &off);
            if (rc) return rc;
            return store(rd(0), off, n);
        }
        if (!q) return fail(E, ALOHA_E_STATE, "ALU op before VSETQ");
        if (wd < 0) return fail(E, ALOHA_E_OPCODE, "no destination register");

        VecOp o{};
        o.n = (u32)n;
        o.q = q;
        o.iq = iq;
        int src_reg = -1;
        if (m.ntt == 2 || m.ntt == 3) {
            src_reg = m.ntt == 2 ? rd(1) : rd(3);
            if (src_reg == wd) return fail(E, ALOHA_E_ILLEGAL, "VNTT/VINTT with vd == vs1");
            int tf = mod_idx;
            if (tf < 0 && (E->cfg.flags & ALOHA_F_STRICT) && !E->mod_q.empty())
                tf = (int)E->mod_q.size() - 1;      // the RTL falls back to its last ROM (vxu_top.sv:115-116, Q6)
            if (tf < 0) return fail(E, ALOHA_E_STATE, "VNTT/VINTT under a modulus with no twiddle ROM (aloha_load_tf_rom)");
            if (n > 65536) return fail(E, ALOHA_E_STATE, "VNTT/VINTT support N <= 65536");
            o.kind = m.ntt == 2 ? K_NTT : K_INTT;
            o.mod = tf;
            int rc = read_loc(src_reg, n, &o.a);
            if (rc) return rc;
            const TwTable *t;
            rc = get_tables(E, tf, ilog2(n), &t);
            if (rc) return rc;
            o.alu = t->mc.form;                 // one launch holds one arithmetic form
            ++limb_ntts;
        } else if (m.iconn == 1 || m.iconn == 2) {
            src_reg = rd(1);
            if (src_reg == wd) return fail(E, ALOHA_E_ILLEGAL, "VAUT/VROLI with vd == vs1");
            int rc = read_loc(src_reg, n, &o.a);
            if (rc) return rc;
            if (m.iconn == 1) {
                o.kind = K_VAUT;
                o.k = m.scalar_iconn & ((1ull << E->kbits) - 1);   // vxu_lane.sv:594 truncation
                if (!(o.k & 1)) return fail(E, ALOHA_E_ILLEGAL, "VAUT with even k");
                u64 inv = o.k;                                      // Newton: k^-1 mod 2^64
                for (int i = 0; i < 6; ++i) inv *= 2 - o.k * inv;
                o.kinv = inv & (n - 1);
            } else {
                o.kind = K_VROLI;
                o.kinv = m.scalar_iconn & (n - 1);
            }
        } else {
            o.kind = K_EW;
            o.alu = (u32)m.alu;
            const bool vv = o.alu == A_MULVV || o.alu == A_ADDVV || o.alu == A_SUBVV;
            int rc = read_loc(rd(3), n, &o.a);
            if (rc) return rc;
            if (vv) {
                rc = read_loc(rd(2), n, &o.b);
                if (rc) return rc;
            }
            o.s = m.scalar_alu >= q ? m.scalar_alu - q : m.scalar_alu;   // modalu.sv:46
        }
        if ((o.kind == K_NTT || o.kind == K_INTT) && (E->cfg.flags & ALOHA_F_STRICT))
            return strict_transform(o, src_reg, wd, n);
        Loc nl;
        int rc = alloc(&nl, n);
        if (rc) return rc;
        o.dst = ptr(nl);
        ops.push_back(o);
        producer[nl.off] = (int)ops.size() - 1;
        define(wd, nl);
        if ((o.kind == K_NTT || o.kind == K_INTT) && src_reg >= 0 && src_reg != wd) {
            // The RTL ping-pongs through vs1 and leaves an intermediate stage there.
            // This engine does not reproduce that content: the register becomes undefined.
            release(loc[src_reg]);
            loc[src_reg] = Loc{};
        }
        return ALOHA_OK;
    }
};

// Peephole fusion over the op list (program order): a VFQMUL.vv whose product feeds exactly one
// VFQADD.vv and is dead afterwards becomes one multiply-add kernel; if one multiplicand is in turn the
// single-use, dead output of a VAUT, all three become the gather-multiply-add kernel.  "Dead" = the
// value is not what any vector register holds when the plan ends (architectural state stays exact).
// Every fused kernel evaluates the same RTL-exact per-element functions in the same order.
size_t fuse_ops(std::vector<VecOp> &ops, const Loc *final_loc, const aloha *E) {
    const size_t n = ops.size();
    std::vector<int> readers(n, 0), prod_a(n, -1), prod_b(n, -1);
    std::unordered_map<const u64 *, int> writer;   // pointer -> index of the op whose value it holds now
    for (size_t i = 0; i < n; ++i) {
        const VecOp &o = ops[i];
        auto src = [&](const u64 *p) {
            auto it = writer.find(p);
            if (p && it != writer.end()) { ++readers[it->second]; return it->second; }
            return -1;
        };
        prod_a[i] = src(o.a);
        prod_b[i] = src(o.b);
        writer[o.dst] = (int)i;
    }
    std::vector<uint8_t> live_out(n, 0);
    for (int r = 0; r < 32; ++r) {
        auto it = writer.find(E->ptr(final_loc[r]));
        if (final_loc[r].space != SP_UNDEF && it != writer.end()) live_out[it->second] = 1;
    }
    auto single_use_temp = [&](int p) {
        // the producer's destination must be a renaming-pool buffer: a value stored to SPM is visible
        const VecOp &o = ops[p];
        const bool in_pool = o.dst >= E->d_pool && o.dst < E->d_pool + (u64)E->pool_count * E->nmax;
        return in_pool && readers[p] == 1 && !live_out[p];
    };
    size_t fused = 0;
    for (size_t i = 0; i < n; ++i) {
        VecOp &add = ops[i];
        if (add.dead || add.kind != K_EW || add.alu != A_ADDVV) continue;
        for (int side = 0; side < 2; ++side) {
            const int pm = side == 0 ? prod_a[i] : prod_b[i];
            if (pm < 0 || ops[pm].dead || ops[pm].kind != K_EW || ops[pm].alu != A_MULVV) continue;
            VecOp &mul = ops[pm];
            if (!single_use_temp(pm) || mul.q != add.q || mul.iq != add.iq || mul.n != add.n) continue;
            // nothing between the multiply and the add may overwrite the multiply's inputs
            bool clobbered = false;
            for (size_t j = pm + 1; j < i && !clobbered; ++j)
                if (!ops[j].dead && (overlap(ops[j].dst, ops[j].n, mul.a, mul.n) || overlap(ops[j].dst, ops[j].n, mul.b, mul.n)))
                    clobbered = true;
            if (clobbered) continue;
            const u64 *addend = side == 0 ? add.b : add.a;
            // index-preserving fused kernels may run in place on an operand, but only on exactly the same
            // range: a partial overlap would read words another thread has already replaced
            auto partial = [&](const u64 *p) { return p != add.dst && overlap(add.dst, add.n, p, add.n); };
            if (partial(mul.a) || partial(mul.b) || partial(addend)) continue;
            // is one multiplicand a single-use, dead automorphism output?
            int pa = -1;
            const u64 *other = nullptr;
            for (int ms = 0; ms < 2 && pa < 0; ++ms) {
                const int cand = ms == 0 ? prod_a[pm] : prod_b[pm];
                if (cand >= 0 && !ops[cand].dead && ops[cand].kind == K_VAUT && ops[cand].q == mul.q && single_use_temp(cand)) {
                    bool bad = false;
                    for (size_t j = cand + 1; j < i && !bad; ++j)
                        if (!ops[j].dead && overlap(ops[j].dst, ops[j].n, ops[cand].a, ops[cand].n)) bad = true;
                    // the fused kernel gathers x while other threads store dst: they must not share words
                    // (store forwarding may already have pointed the add at the SPM range x lives in)
                    if (overlap(add.dst, add.n, ops[cand].a, ops[cand].n)) bad = true;
                    if (!bad) { pa = cand; other = ms == 0 ? mul.b : mul.a; }
                }
            }
            if (pa >= 0) {
                add.kind = K_AUTMAC;
                add.a = ops[pa].a;          // x (gathered)
                add.b = other;              // p
                add.k = ops[pa].k;
                add.kinv = ops[pa].kinv;
                ops[pa].dead = true;
                ++fused;
            } else {
                add.kind = K_MULADD;
                add.a = mul.a;
                add.b = mul.b;
            }
            add.c = addend;
            add.alu = 0;
            mul.dead = true;
            ++fused;
            break;
        }
    }
    // Fast basis extension (hks.py: digits of several limbs, several special primes): per source limb
    //   VCPY | VFQMOD  ->  VFQMUL.vs  ->  VFQADD.vv into a running sum   [ -> VFQSUB.vs ]
    // with every intermediate a single-use, dead temporary.  The whole sum becomes one K_BEXT op that
    // evaluates the same RTL functions in the same order.
    {
        auto intact = [&](size_t from, size_t to, const u64 *p, u64 len) {      // nobody in (from, to) writes [p, p+len)
            for (size_t j = from + 1; j < to; ++j)
                if (!ops[j].dead && overlap(ops[j].dst, ops[j].n, p, len)) return false;
            return true;
        };
        auto same_mod = [&](const VecOp &x, const VecOp &y) { return x.q == y.q && x.iq == y.iq && x.n == y.n; };
        // the summands op `m` stands for when consumed at `at` (m is a partial sum, or a VFQMUL.vs possibly fed
        // by a VCPY / VFQMOD), and the ops that disappear if the caller fuses them
        auto as_ext = [&](int m, size_t at, std::vector<ExtTerm> *out, std::vector<int> *kill) {
            VecOp &mo = ops[m];
            if (mo.dead || !single_use_temp(m)) return false;
            if (mo.kind == K_BEXT && !mo.post) {
                for (auto &t : mo.ext) if (!intact(m, at, t.x, mo.n)) return false;
                *out = mo.ext;
                kill->push_back(m);
                return true;
            }
            if (mo.kind != K_EW || mo.alu != A_MULVS) return false;
            const int pe = prod_a[m];
            if (pe >= 0 && !ops[pe].dead && ops[pe].kind == K_EW && same_mod(ops[pe], mo) && single_use_temp(pe)) {
                const VecOp &e = ops[pe];
                const u32 pre = (e.alu == A_ADDVS && e.s == 0) ? (u32)PRE_VCPY : e.alu == A_MOD ? (u32)PRE_VFQMOD : 0u;
                if (pre && intact(pe, at, e.a, e.n)) {
                    out->push_back(ExtTerm{e.a, mo.s, pre});
                    kill->push_back(pe);
                    kill->push_back(m);
                    return true;
                }
            }
            if (!intact(m, at, mo.a, mo.n)) return false;
            out->push_back(ExtTerm{mo.a, mo.s, 0});
            kill->push_back(m);
            return true;
        };
        for (size_t i = 0; i < n; ++i) {
            VecOp &o = ops[i];
            if (o.dead || o.kind != K_EW) continue;
            if (o.alu == A_ADDVV) {
                const int pa = prod_a[i], pb = prod_b[i];
                if (pa < 0 || pb < 0 || pa == pb || !same_mod(ops[pa], o) || !same_mod(ops[pb], o)) continue;
                std::vector<ExtTerm> ta, tb;
                std::vector<int> kill;
                if (!as_ext(pa, i, &ta, &kill) || !as_ext(pb, i, &tb, &kill) || ta.size() + tb.size() > 64) continue;
                bool alias = false;             // the fused kernel reads every x while it writes dst
                for (auto *v : {&ta, &tb}) for (auto &t : *v) if (t.x != o.dst && overlap(o.dst, o.n, t.x, o.n)) alias = true;
                if (alias) continue;
                for (int d : kill) { ops[d].dead = true; ++fused; }
                ta.insert(ta.end(), tb.begin(), tb.end());
                o.kind = K_BEXT;
                o.ext = std::move(ta);
                o.a = o.b = nullptr;
                o.alu = 0;
            } else if (o.alu == A_SUBVS) {
                const int pa = prod_a[i];
                if (pa < 0 || ops[pa].dead || ops[pa].kind != K_BEXT || ops[pa].post || !same_mod(ops[pa], o) || !single_use_temp(pa)) continue;
                bool ok = true;
                for (auto &t : ops[pa].ext) if (!intact(pa, i, t.x, o.n) || (t.x != o.dst && overlap(o.dst, o.n, t.x, o.n))) ok = false;
                if (!ok) continue;
                o.kind = K_BEXT;
                o.ext = ops[pa].ext;
                o.post = 1;
                o.post_s = o.s;
                o.a = nullptr;
                o.alu = 0;
                ops[pa].dead = true;
                ++fused;
            }
        }
    }
    // VCPY / VFQMOD feeding exactly one forward transform under the same modulus (the key-switch base
    // extension): the transform applies the op while loading, the intermediate never exists.
    for (size_t i = 0; i < n; ++i) {
        VecOp &t = ops[i];
        if (t.dead || t.kind != K_NTT || t.pre) continue;
        const int pe = prod_a[i];
        if (pe < 0 || ops[pe].dead || ops[pe].kind != K_EW || ops[pe].q != t.q || ops[pe].n != t.n) continue;
        const VecOp &e = ops[pe];
        u32 pre = (e.alu == A_ADDVS && e.s == 0) ? (u32)PRE_VCPY : e.alu == A_MOD ? (u32)PRE_VFQMOD : 0u;
        // the transform's load stage computes VFQMOD as an exact x mod q, which is what barrett(r(x), 1) returns
        // only under q's own Barrett constant: with any other VSETIQ the op stays a separate RTL-exact kernel
        if (pre == PRE_VFQMOD && e.iq != (u64)((((u128)1) << 121) / e.q)) pre = 0;
        if (!pre || !single_use_temp(pe)) continue;
        bool clobbered = false;
        for (size_t j = pe + 1; j < i && !clobbered; ++j)
            if (!ops[j].dead && overlap(ops[j].dst, ops[j].n, e.a, e.n)) clobbered = true;
        if (clobbered) continue;
        t.pre = pre;
        t.a = e.a;
        ops[pe].dead = true;
        ++fused;
    }
    // Accumulation chains: acc_t = acc_{t-1} + a_t b_t where acc_{t-1} is itself a single-use, dead
    // product or multiply-add collapse into one sum-of-products op (same evaluation order).
    {
        // recompute producers/readers on the rewritten list (indices unchanged; dead ops skipped)
        std::fill(readers.begin(), readers.end(), 0);
        std::vector<int> prod_c(n, -1);
        writer.clear();
        for (size_t i = 0; i < n; ++i) {
            const VecOp &o = ops[i];
            if (o.dead) continue;
            auto src = [&](const u64 *p) {
                auto it = writer.find(p);
                if (p && it != writer.end()) { ++readers[it->second]; return it->second; }
                return -1;
            };
            src(o.a); src(o.b);
            for (auto &t : o.ext) src(t.x);                               // operands of the ops fused above are reads too
            for (auto &t : o.terms) { src(t.first); src(t.second); }
            prod_c[i] = src(o.c);
            writer[o.dst] = (int)i;
        }
        for (size_t i = 0; i < n; ++i) {
            VecOp &o = ops[i];
            if (o.dead || o.kind != K_MULADD) continue;
            const int pc = prod_c[i];
            if (pc < 0 || ops[pc].dead || !single_use_temp(pc) || ops[pc].q != o.q || ops[pc].iq != o.iq || ops[pc].n != o.n) continue;
            VecOp &prev = ops[pc];
            std::vector<std::pair<const u64 *, const u64 *>> terms;
            if (prev.kind == K_SOP) terms = prev.terms;
            else if (prev.kind == K_EW && prev.alu == A_MULVV) terms.emplace_back(prev.a, prev.b);
            else continue;
            if (terms.size() >= 1024) continue;
            // none of the collected operands may be overwritten between the head of the chain and here
            bool clobbered = false;
            for (size_t j = pc + 1; j < i && !clobbered; ++j) {
                if (ops[j].dead) continue;
                for (auto &tm : terms)
                    if (overlap(ops[j].dst, ops[j].n, tm.first, o.n) || overlap(ops[j].dst, ops[j].n, tm.second, o.n)) { clobbered = true; break; }
            }
            for (auto &tm : terms)
                if ((tm.first != o.dst && overlap(o.dst, o.n, tm.first, o.n)) || (tm.second != o.dst && overlap(o.dst, o.n, tm.second, o.n)))
                    clobbered = true;
            if (clobbered) continue;
            terms.emplace_back(o.a, o.b);
            o.kind = K_SOP;
            o.terms = std::move(terms);
            o.a = o.b = o.c = nullptr;
            prev.dead = true;
            ++fused;
        }
    }
    if (fused) {
        std::vector<VecOp> kept;
        kept.reserve(n);
        for (auto &o : ops) if (!o.dead) kept.push_back(o);
        ops.swap(kept);
    }
    return fused;
}

// ASAP levels from true dependencies on address ranges.
struct RangeState { u64 n; int last_write, last_read; };

void assign_levels(std::vector<VecOp> &ops, bool sequential) {
    if (sequential) {
        for (size_t i = 0; i < ops.size(); ++i) ops[i].level = (int)i + 1;
        return;
    }
    std::multimap<uintptr_t, RangeState> ranges;   // keyed by start address (bytes)
    u64 maxlen = 0;
    for (auto &o : ops) maxlen = std::max<u64>(maxlen, o.n);
    auto scan = [&](const u64 *p, u64 n, bool want_readers) {
        int lvl = 0;
        if (!p) return lvl;
        const uintptr_t lo = (uintptr_t)p, hi = lo + n * 8;
        for (auto it = ranges.lower_bound(lo - std::min<uintptr_t>(lo, maxlen * 8)); it != ranges.end() && it->first < hi; ++it) {
            if (it->first + it->second.n * 8 <= lo) continue;
            lvl = std::max(lvl, it->second.last_write);
            if (want_readers) lvl = std::max(lvl, it->second.last_read);
        }
        return lvl;
    };
    auto touch = [&](const u64 *p, u64 n, int level, bool is_write) {
        if (!p) return;
        auto range = ranges.equal_range((uintptr_t)p);
        for (auto it = range.first; it != range.second; ++it) {
            if (it->second.n == n) {
                if (is_write) it->second.last_write = std::max(it->second.last_write, level);
                else it->second.last_read = std::max(it->second.last_read, level);
                return;
            }
        }
        ranges.emplace((uintptr_t)p, RangeState{n, is_write ? level : 0, is_write ? 0 : level});
    };
    for (auto &o : ops) {
        int lvl = std::max(scan(o.a, o.n, false), scan(o.b, o.n, false));
        lvl = std::max(lvl, scan(o.c, o.n, false));
        for (auto &tm : o.terms) lvl = std::max(lvl, std::max(scan(tm.first, o.n, false), scan(tm.second, o.n, false)));
        for (auto &tm : o.ext) lvl = std::max(lvl, scan(tm.x, o.n, false));
        lvl = std::max(lvl, scan(o.dst, o.n, true));
        o.level = lvl + 1;
        for (auto &tm : o.terms) { touch(tm.first, o.n, o.level, false); touch(tm.second, o.n, o.level, false); }
        for (auto &tm : o.ext) touch(tm.x, o.n, o.level, false);
        touch(o.a, o.n, o.level, false);
        touch(o.b, o.n, o.level, false);
        touch(o.c, o.n, o.level, false);
        touch(o.dst, o.n, o.level, true);
    }
}

template <class T>
size_t append(std::vector<uint8_t> &buf, const T &v) {
    const size_t off = buf.size();
    buf.resize(off + sizeof(T));
    std::memcpy(buf.data() + off, &v, sizeof(T));
    return off;
}

int compile_plan(aloha *E, Builder &B, Plan *plan) {
    std::vector<VecOp> &ops = B.ops;
    if (!(E->cfg.flags & (ALOHA_F_NO_BATCH | ALOHA_F_NO_FUSE))) plan->fused = fuse_ops(ops, B.loc, E);
    assign_levels(ops, E->cfg.flags & ALOHA_F_NO_BATCH);
    std::vector<size_t> order(ops.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) {
        const VecOp &a = ops[x], &b = ops[y];
        if (a.level != b.level) return a.level < b.level;
        if (a.kind != b.kind) return a.kind < b.kind;
        if (a.alu != b.alu) return a.alu < b.alu;
        return a.n < b.n;
    });
    std::vector<uint8_t> tables;
    std::vector<std::pair<size_t, size_t>> fixups;   // (offset of a pointer field, offset it must point to) inside `tables`
    const u64 chunk_bytes = E->cfg.l2_chunk_bytes ? E->cfg.l2_chunk_bytes : ~0ull;   // default: one launch pair
    size_t i = 0;
    while (i < order.size()) {
        const VecOp &h = ops[order[i]];
        size_t j = i;
        while (j < order.size() && ops[order[j]].level == h.level && ops[order[j]].kind == h.kind &&
               ops[order[j]].alu == h.alu && ops[order[j]].n == h.n)
            ++j;
        // transforms: keep one chunk's src+dst footprint inside L2 between the two passes
        size_t max_jobs = 65535;
        if (h.kind == K_NTT || h.kind == K_INTT)
            max_jobs = std::max<u64>(1, chunk_bytes / (2ull * h.n * 8));
        while (tables.size() % 16) tables.push_back(0);
        for (size_t c = i; c < j;) {
            const size_t cnt = std::min(max_jobs, j - c);
            Launch L{h.kind, h.alu, h.n, (u32)cnt, tables.size()};
            if ((h.kind == K_NTT || (h.kind == K_INTT && E->tma_maps_ok)) && cnt >= 16) {
                // transforms: same-modulus runs of 16 first (they take the TMA-staged row pass,
                // one tile = one row of 16 polynomials sharing its twiddles), the remainder after them
                std::map<std::pair<int, u32>, std::vector<size_t>> by_mod;
                for (size_t t = c; t < c + cnt; ++t) by_mod[{ops[order[t]].mod, ops[order[t]].pre}].push_back(order[t]);
                std::vector<size_t> grouped, rest;
                for (auto &kv : by_mod) {
                    const size_t full = kv.second.size() / 16 * 16;
                    grouped.insert(grouped.end(), kv.second.begin(), kv.second.begin() + full);
                    rest.insert(rest.end(), kv.second.begin() + full, kv.second.end());
                }
                L.ngroups = (u32)(grouped.size() / 16);
                std::copy(grouped.begin(), grouped.end(), order.begin() + c);
                std::copy(rest.begin(), rest.end(), order.begin() + c + grouped.size());
            }
            std::vector<std::pair<size_t, const VecOp *>> sop_jobs, bext_jobs;
            for (size_t t = c; t < c + cnt; ++t) {
                const VecOp &o = ops[order[t]];
                switch (o.kind) {
                case K_EW: append(tables, EwJob{o.dst, o.a, o.b, o.s, o.q, o.iq}); break;
                case K_COPY: append(tables, CopyJob{o.dst, o.a}); break;
                case K_PEASE_F:
                case K_PEASE_I: {
                    const TwTable *tw;
                    int rc = get_tables(E, o.mod, ilog2(o.n), &tw);
                    if (rc) return rc;
                    append(tables, PeaseJob{o.dst, o.a, o.kind == K_PEASE_F ? tw->fwd : tw->inv, o.q, o.iq});
                    break;
                }
                case K_MULADD: append(tables, MulAddJob{o.dst, o.c, o.a, o.b, o.q, o.iq}); break;
                case K_SOP: {
                    const size_t at = append(tables, SopJob{o.dst, nullptr, o.q, o.iq, (u32)o.terms.size(), 0});
                    sop_jobs.emplace_back(at, &o);
                    break;
                }
                case K_BEXT: {
                    // the lazy path needs q to be a 60-bit modulus with its own Barrett constant and canonical scalars
                    bool fast = o.q > (1ull << 59) && o.q < (1ull << 60) && o.iq == (u64)((((u128)1) << 121) / o.q) &&
                                (!o.post || o.post_s < o.q);
                    for (auto &tm : o.ext) fast = fast && tm.s < o.q;
                    const u32 mest = fast ? (u32)((((u128)1) << 91) / o.q) : 0u;
                    const size_t at = append(tables, BextJob{o.dst, nullptr, o.q, o.iq, o.post_s, (u32)o.ext.size(), o.post, fast ? 1u : 0u, mest});
                    bext_jobs.emplace_back(at, &o);
                    break;
                }
                case K_AUTMAC: {
                    const AutPlan &ap = aut_plan_for(E, o.n, o.k);
                    L.aux = std::max(L.aux, ap.ntiles);
                    append(tables, AutMacJob{o.dst, o.c, o.a, o.b, o.q, o.iq, o.k, o.kinv, ap});
                    break;
                }
                case K_VAUT:
                    if (!(E->cfg.flags & ALOHA_F_AUT_GATHER)) {
                        const AutPlan &ap = aut_plan_for(E, o.n, o.k);
                        L.aux = std::max(L.aux, ap.ntiles);
                        append(tables, AutJob{o.dst, o.a, o.q, o.k, ap});
                        break;
                    }
                    [[fallthrough]];
                case K_VROLI: append(tables, PermJob{o.dst, o.a, o.q, o.k, o.kinv}); break;
                case K_NTT:
                case K_INTT: {
                    const TwTable *tw;
                    int rc = get_tables(E, o.mod, ilog2(o.n), &tw);
                    if (rc) return rc;
                    NttJob nj{o.a, o.dst, o.kind == K_NTT ? tw->fwd : tw->inv, o.kind == K_NTT ? tw->fwd_rows : tw->inv_rows, tw->mc};
                    nj.mc.pre = o.pre;
                    append(tables, nj);
                    break;
                }
                }
            }
            if (L.ngroups) {                   // group records of the TMA-staged row pass, after the job table
                L.group_off = tables.size();
                for (u32 g = 0; g < L.ngroups; ++g) {
                    NttRowGroup G{};
                    const NttJob *nj = reinterpret_cast<const NttJob *>(tables.data() + L.table_off) + 16 * g;
                    for (int k = 0; k < 16; ++k) {
                        G.src[k] = (h.kind == K_INTT || h.n == 256) ? nj[k].src : nj[k].dst;
                        G.dst[k] = nj[k].dst;
                        if (h.kind == K_INTT && !tensor_coords(E, G.src[k], &G.src_map[k], &G.src_line[k]))
                            return fail(E, ALOHA_E_STATE, "transform source outside the device buffers");
                    }
                    G.rtw = nj[0].rtw;
                    G.mc = nj[0].mc;
                    append(tables, G);
                }
            }
            for (auto &sj : sop_jobs) {       // operand-pointer lists go right after the launch's job table
                while (tables.size() % 16) tables.push_back(0);
                fixups.emplace_back(sj.first + offsetof(SopJob, pairs), tables.size());
                for (auto &tm : sj.second->terms) { append(tables, tm.first); append(tables, tm.second); }
            }
            for (auto &bj : bext_jobs) {
                while (tables.size() % 16) tables.push_back(0);
                fixups.emplace_back(bj.first + offsetof(BextJob, terms), tables.size());
                for (auto &tm : bj.second->ext)
                    append(tables, BextTerm{tm.x, tm.s, tm.s < bj.second->q ? (u64)(((u128)tm.s << 64) / bj.second->q) : 0ull, tm.pre});
            }
            while (tables.size() % 16) tables.push_back(0);
            plan->launches.push_back(L);
            c += cnt;
        }
        i = j;
    }
    plan->table_bytes = tables.size();
    if (!tables.empty()) {
        CU(cudaMalloc(&plan->d_tables, tables.size()));
        for (auto &fx : fixups) {
            const uint8_t *target = (const uint8_t *)plan->d_tables + fx.second;
            std::memcpy(tables.data() + fx.first, &target, sizeof target);
        }
        CU(cudaMemcpyAsync(plan->d_tables, tables.data(), tables.size(), cudaMemcpyHostToDevice, E->stream));
        CU(cudaStreamSynchronize(E->stream));   // `tables` is a pageable temporary
    }
    plan->vl = B.vl; plan->q = B.q; plan->iq = B.iq; plan->mod_idx = B.mod_idx;
    for (int r = 0; r < 32; ++r) plan->loc[r] = B.loc[r];
    plan->written = B.written;
    plan->live_in_mask = B.live_in_mask;
    plan->killed_mask = B.killed_mask;
    for (int r = 0; r < 32; ++r) plan->live_in[r] = B.live_in[r];
    plan->alloc_set = B.alloc_set;
    plan->instructions = B.instructions;
    plan->limb_ntts = B.limb_ntts;
    plan->elided = B.elided;
    plan->emitted = B.emitted;
    return ALOHA_OK;
}

int issue(aloha *E, const Plan &plan, u64 *launched) {
    const uint8_t *base = (const uint8_t *)plan.d_tables;
    const unsigned long long before = kernel_launch_count();
    for (const Launch &L : plan.launches) {
        const void *tab = base + L.table_off;
        cudaError_t e = cudaSuccess;
        switch (L.kind) {
        case K_EW: e = launch_ew(L.alu, (const EwJob *)tab, L.njobs, L.n, E->stream); break;
        case K_COPY: e = launch_copy((const CopyJob *)tab, L.njobs, L.n, E->stream); break;
        case K_PEASE_F: e = launch_pease((const PeaseJob *)tab, L.njobs, ilog2(L.n), L.alu, false, E->stream); break;
        case K_PEASE_I: e = launch_pease((const PeaseJob *)tab, L.njobs, ilog2(L.n), L.alu, true, E->stream); break;
        case K_MULADD: e = launch_muladd((const MulAddJob *)tab, L.njobs, L.n, E->stream); break;
        case K_SOP: e = launch_sop((const SopJob *)tab, L.njobs, L.n, E->stream); break;
        case K_BEXT: e = launch_bext((const BextJob *)tab, L.njobs, L.n, E->stream); break;
        case K_AUTMAC:
            e = (E->cfg.flags & ALOHA_F_AUT_TILED) ? launch_autmac_tiled((const AutMacJob *)tab, L.njobs, L.n, L.aux, E->stream)
                                                   : launch_autmac((const AutMacJob *)tab, L.njobs, L.n, E->stream);
            break;
        case K_VAUT:
            e = (E->cfg.flags & ALOHA_F_AUT_GATHER) ? launch_vaut((const PermJob *)tab, L.njobs, L.n, E->stream)
                                                    : launch_vaut_tiled((const AutJob *)tab, L.njobs, L.n, L.aux, E->stream);
            break;
        case K_VROLI: e = launch_vroli((const PermJob *)tab, L.njobs, L.n, E->stream); break;
        case K_NTT: e = launch_ntt_forward((const NttJob *)tab, L.njobs, (const NttRowGroup *)(base + L.group_off), L.ngroups, ilog2(L.n), L.alu, E->stream); break;
        case K_INTT: e = launch_ntt_inverse((const NttJob *)tab, L.njobs, (const NttRowGroup *)(base + L.group_off), L.ngroups, &E->tma_maps, ilog2(L.n), L.alu, E->stream); break;
        }
        if (e != cudaSuccess) {
            E->last_error = std::string("kernel launch: ") + cudaGetErrorString(e);
            return ALOHA_E_CUDA;
        }
    }
    *launched = kernel_launch_count() - before;
    return ALOHA_OK;
}

int execute_plan(aloha *E, Plan &plan) {
    u64 launched = 0;
    if ((E->cfg.flags & ALOHA_F_GRAPHS) && !plan.launches.empty()) {
        if (!plan.graph) {
            cudaGraph_t g;
            CU(cudaStreamBeginCapture(E->stream, cudaStreamCaptureModeThreadLocal));
            int rc = issue(E, plan, &launched);
            g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(E->stream, &g);
            if (rc || ce != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                if (rc) return rc;
                CU(ce);
            }
            plan.kernel_launches = launched;
            ce = cudaGraphInstantiate(&plan.graph, g, 0);
            cudaGraphDestroy(g);
            CU(ce);
        }
        CU(cudaGraphLaunch(plan.graph, E->stream));
        launched = plan.kernel_launches;
    } else {
        int rc = issue(E, plan, &launched);
        if (rc) return rc;
    }
    E->stats.kernel_launches += launched;
    return ALOHA_OK;
}

void mark_written(aloha *E, u64 word_off, u64 nwords) {
    const u64 b0 = word_off / 8, b1 = (word_off + nwords + 7) / 8;
    std::memset(E->written.data() + b0, 1, b1 - b0);
}

void commit(aloha *E, Plan &plan) {
    E->vl = plan.vl; E->q = plan.q; E->iq = plan.iq; E->mod_idx = plan.mod_idx;
    // registers the plan never defined keep their location (an aliased register moved by a
    // copy-on-write is recorded as read + moved, hence its exit location is the plan's too)
    for (int r = 0; r < 32; ++r)
        if (((plan.killed_mask | plan.live_in_mask) >> r) & 1) E->loc[r] = plan.loc[r];
    // the flags are never cleared, so a replayed plan has nothing left to mark (a 2048-polynomial step would
    // otherwise memset 16 MiB of flags on the host every time it is replayed)
    if (!plan.written_marked) {
        for (auto &w : plan.written) mark_written(E, w.first, w.second);
        plan.written_marked = true;
    }
    E->stats.instructions += plan.instructions;
    E->stats.limb_ntts += plan.limb_ntts;
    E->stats.copies_elided += plan.elided;
    E->stats.copies_emitted += plan.emitted;
    E->stats.ops_fused += plan.fused;
}

// Events are recycled: creating and destroying one per DMA call costs more than the call's own enqueue.
int get_event(aloha *E, cudaEvent_t *ev) {
    if (!E->event_pool.empty()) { *ev = E->event_pool.back(); E->event_pool.pop_back(); return ALOHA_OK; }
    CU(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    return ALOHA_OK;
}
void put_event(aloha *E, cudaEvent_t ev) { E->event_pool.push_back(ev); }
// `waiter` waits for everything queued on `signaller` so far
int order_after(aloha *E, cudaStream_t waiter, cudaStream_t signaller) {
    cudaEvent_t ev;
    int rc = get_event(E, &ev);
    if (rc) return rc;
    CU(cudaEventRecord(ev, signaller));
    CU(cudaStreamWaitEvent(waiter, ev, 0));
    put_event(E, ev);           // the wait has captured the event's current record; re-recording later is safe
    return ALOHA_OK;
}

// Work about to be queued on `st` writes SPM words [off, off+n): it must not overtake a pending
// asynchronous download of overlapping rows.  Waiting on an event that has already completed costs nothing on the
// device, so finished downloads are not looked for here (a completion query per pending download per call is a
// microsecond each, quadratic over a program of dumped ops): aloha_sync retires them all, and a list that has grown
// long (64) without a sync is pruned once.
int wait_for_downloads(aloha *E, cudaStream_t st, u64 off, u64 n) {
    if (E->pending_down.size() >= 64) {
        for (size_t i = 0; i < E->pending_down.size();) {
            if (cudaEventQuery(E->pending_down[i].done) == cudaSuccess) {
                put_event(E, E->pending_down[i].done);
                E->pending_down.erase(E->pending_down.begin() + i);
            } else {
                ++i;
            }
        }
    }
    for (auto &p : E->pending_down)
        if (p.off < off + n && off < p.off + p.n) CU(cudaStreamWaitEvent(st, p.done, 0));
    return ALOHA_OK;
}

std::string plan_key(const aloha *E, const uint32_t *pcs, bool same_pc, uint32_t count, const aloha_vp_args *args) {
    std::string k;
    auto put = [&](const void *p, size_t n) { k.append((const char *)p, n); };
    put(pcs, same_pc ? 4 : 4 * (size_t)count); put(&count, 4);
    put(args, sizeof(aloha_vp_args) * count);
    put(&E->vl, 8); put(&E->q, 8); put(&E->iq, 8); put(&E->mod_idx, 4);
    put(&E->isram_version, 8); put(&E->tf_version, 8);
    return k;
}

bool same_loc(const Loc &a, const Loc &b) { return a.space == b.space && a.off == b.off && a.n == b.n; }

// May `plan` be replayed from the machine's current register state?
bool plan_fits(const aloha *E, const Plan &plan) {
    for (int r = 0; r < 32; ++r) {
        const Loc &cur = E->loc[r];
        if ((plan.live_in_mask >> r) & 1) {
            if (!same_loc(cur, plan.live_in[r])) return false;
            continue;
        }
        const bool killed = (plan.killed_mask >> r) & 1;
        if (killed) continue;            // never read, then overwritten: whatever it holds is dead
        if (cur.space == SP_POOL) {
            if (plan.alloc_set[cur.off]) return false;   // the plan would scribble on a live value
        } else if (cur.space == SP_SPM) {
            // an aliased register whose range the plan stores to would have needed a copy-on-write
            for (auto &w : plan.written)
                if (cur.off < w.first + w.second && w.first < cur.off + cur.n) return false;
        }
    }
    return true;
}

// pcs: one entry (same_pc) or `count` entries
int run_batch(aloha *E, const uint32_t *pcs, bool same_pc, uint32_t count, const aloha_vp_args *args) {
    if (!count) return ALOHA_OK;
    const std::string key = plan_key(E, pcs, same_pc, count, args);
    if (E->plans.size() >= 1024 && !E->plans.count(key)) {
        CU(cudaStreamSynchronize(E->stream));
        free_plans(E);
    }
    std::vector<Plan> &cands = E->plans[key];
    Plan *hit = nullptr;
    for (auto &pl : cands)
        if (plan_fits(E, pl)) { hit = &pl; break; }
    if (!hit) {
        Builder B(E);
        for (uint32_t c = 0; c < count; ++c) {
            bool brk = false;
            for (u64 at = same_pc ? pcs[0] : pcs[c]; !brk; ++at) {
                if (at >= E->iram_depth) return fail(E, ALOHA_E_NOBREAK, "ran off the instruction ROM without BREAK");
                const Inst in = parse_word(&E->isram[at * 12]);
                int rc = B.step(in, args[c], &brk);
                if (rc) return rc;
            }
        }
        Plan plan;
        int rc = compile_plan(E, B, &plan);
        if (rc) { cudaFree(plan.d_tables); return rc; }
        if (cands.size() >= 16) {            // pathological churn: drop this key's candidates
            CU(cudaStreamSynchronize(E->stream));
            for (auto &pl : cands) { if (pl.graph) cudaGraphExecDestroy(pl.graph); cudaFree(pl.d_tables); }
            cands.clear();
        }
        cands.push_back(std::move(plan));
        hit = &cands.back();
        ++E->stats.plans_built;
    } else {
        ++E->stats.plans_reused;
    }
    if (!E->pending_down.empty()) {
        int rc = wait_for_downloads(E, E->stream, 0, 0);        // (empty range: only prunes an overgrown list)
        if (rc) return rc;
        for (auto &p : E->pending_down)
            for (auto &w : hit->written)
                if (p.off < w.first + w.second && w.first < p.off + p.n) {
                    CU(cudaStreamWaitEvent(E->stream, p.done, 0));
                    break;
                }
    }
    int rc = execute_plan(E, *hit);
    if (rc) return rc;
    commit(E, *hit);
    return ALOHA_OK;
}

// ALOHA_F_DEFER: plan and launch everything queued so far as one batch.
int flush_queue(aloha *E) {
    if (E->queued_pcs.empty()) return ALOHA_OK;
    std::vector<uint32_t> pcs;
    std::vector<aloha_vp_args> args;
    pcs.swap(E->queued_pcs);
    args.swap(E->queued_args);
    int rc = run_batch(E, pcs.data(), false, (uint32_t)pcs.size(), args.data());
    if (rc == ALOHA_OK || rc == ALOHA_E_CUDA) return rc;
    // A queued call is malformed.  Planning fails before anything is launched, so replay the queue call
    // by call: everything before the offender takes effect, exactly as without ALOHA_F_DEFER.
    for (size_t c = 0; c < pcs.size(); ++c) {
        rc = run_batch(E, &pcs[c], true, 1, &args[c]);
        if (rc) return rc;
    }
    return ALOHA_OK;
}
#define FLUSH()                          \
    do {                                 \
        int frc_ = flush_queue(E);       \
        if (frc_) return frc_;           \
    } while (0)

int enqueue_or_run(aloha *E, const uint32_t *pcs, bool same_pc, uint32_t count, const aloha_vp_args *args) {
    if (!(E->cfg.flags & ALOHA_F_DEFER)) return run_batch(E, pcs, same_pc, count, args);
    for (uint32_t c = 0; c < count; ++c) {
        E->queued_pcs.push_back(same_pc ? pcs[0] : pcs[c]);
        E->queued_args.push_back(args[c]);
    }
    if (E->queued_pcs.size() >= 4096) return flush_queue(E);   // bound the plan size
    return ALOHA_OK;
}

// A host-side write into SPM words [off, off+n): registers aliasing the range move out first.
int cow_for_host_write(aloha *E, u64 off, u64 n) {
    bool any = false;
    for (int r = 0; r < 32; ++r)
        if (E->loc[r].space == SP_SPM && E->loc[r].off < off + n && off < E->loc[r].off + E->loc[r].n) any = true;
    if (!any) return ALOHA_OK;
    Builder B(E);
    int rc = B.cow(SP_SPM, off, n, -1);
    if (rc) return rc;
    Plan plan;
    rc = compile_plan(E, B, &plan);
    if (!rc) rc = execute_plan(E, plan);
    if (!rc) {
        const cudaError_t se = cudaStreamSynchronize(E->stream);
        if (se != cudaSuccess) rc = fail(E, ALOHA_E_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(se));
        else commit(E, plan);
    }
    if (plan.graph) cudaGraphExecDestroy(plan.graph);
    cudaFree(plan.d_tables);
    return rc;
}

}  // namespace

// ================================================================================== C-ABI
extern "C" {

const char *aloha_strerror(int code) {
    switch (code) {
    case ALOHA_OK: return "ok";
    case ALOHA_E_ARG: return "bad argument";
    case ALOHA_E_RANGE: return "address out of range";
    case ALOHA_E_OPCODE: return "unknown or malformed instruction";
    case ALOHA_E_STATE: return "machine not configured for this operation";
    case ALOHA_E_ILLEGAL: return "instruction stream has no defined behaviour";
    case ALOHA_E_NOBREAK: return "no BREAK before the end of the instruction ROM";
    case ALOHA_E_UNDEFINED: return "read of an undefined vector register";
    case ALOHA_E_CUDA: return "CUDA error";
    case ALOHA_E_NOMEM: return "out of memory";
    default: return "unknown error";
    }
}

const char *aloha_last_error(const aloha_t *E) { return E ? E->last_error.c_str() : "null handle"; }

int aloha_create(const aloha_cfg *cfg, aloha_t **out) {
    if (!cfg || !out) return ALOHA_E_ARG;
    const u64 nmax = cfg->vlmax_bits / 64;
    if (cfg->vlmax_bits % 64 || nmax < 256 || nmax > 131072 || (nmax & (nmax - 1)) || !cfg->spm_rows) return ALOHA_E_ARG;
    aloha *E = new aloha();
    E->cfg = *cfg;
    E->nmax = nmax;
    E->kbits = ilog2(nmax);
    E->device = cfg->device;
    E->pool_count = cfg->pool_buffers ? cfg->pool_buffers : 256;   // (two key-switch kernels' worth of temporaries per batched call, x 4)
    if (E->pool_count < 34) E->pool_count = 34;
    *out = E;   // so the caller can read last_error on failure, then destroy
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device) {
        E->last_error = "no CUDA device " + std::to_string(cfg->device) + " (this engine has no CPU fallback)";
        return ALOHA_E_CUDA;
    }
    DeviceGuard dg_(cfg->device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) {
        E->last_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; kernels are built for sm_100a only";
        return ALOHA_E_CUDA;
    }
    CU(cudaStreamCreateWithFlags(&E->own_stream, cudaStreamNonBlocking));
    E->stream = E->own_stream;
    E->spm_words = (u64)cfg->spm_rows * kLanes;
    E->ksk_words = (u64)cfg->ksk_rows * kLanes;
    CU(cudaMalloc(&E->d_spm, E->spm_words * 8));
    CU(cudaMemsetAsync(E->d_spm, 0, E->spm_words * 8, E->stream));
    if (E->ksk_words) {
        CU(cudaMalloc(&E->d_ksk, E->ksk_words * 8));
        CU(cudaMemsetAsync(E->d_ksk, 0, E->ksk_words * 8, E->stream));
    }
    CU(cudaMalloc(&E->d_pool, (u64)E->pool_count * nmax * 8));
    build_tensor_maps(E);
    E->iram_depth = cfg->isram_depth ? cfg->isram_depth : kIramDepthDefault;
    E->isram.assign(E->iram_depth * 12, 0);
    E->written.assign((E->spm_words + 7) / 8, 0);
    CU(cudaStreamSynchronize(E->stream));
    return ALOHA_OK;
}

void aloha_destroy(aloha_t *E) {
    if (!E) return;
    DeviceGuard dg_(E->device);
    if (E->own_stream) cudaStreamSynchronize(E->own_stream);
    free_plans(E);
    free_tables(E);
    cudaFree(E->d_spm);
    cudaFree(E->d_ksk);
    cudaFree(E->d_pool);
    if (E->own_stream) cudaStreamDestroy(E->own_stream);
    if (E->up_stream) { cudaStreamSynchronize(E->up_stream); cudaStreamDestroy(E->up_stream); }
    if (E->down_stream) { cudaStreamSynchronize(E->down_stream); cudaStreamDestroy(E->down_stream); }
    for (auto &p : E->pending_down) cudaEventDestroy(p.done);
    for (auto ev : E->event_pool) cudaEventDestroy(ev);
    delete E;
}

int aloha_load_isram(aloha_t *E, const uint8_t *words, uint32_t n, uint32_t at_pc) {
    if (!E || !words) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    if ((u64)at_pc + n > E->iram_depth) return fail(E, ALOHA_E_RANGE, "beyond the instruction ROM (cfg.isram_depth)");
    std::memcpy(&E->isram[(size_t)at_pc * 12], words, (size_t)n * 12);
    ++E->isram_version;
    return ALOHA_OK;
}

int aloha_load_tf_rom(aloha_t *E, const uint64_t *q, const uint64_t *psi, uint32_t n) {
    if (!E || !q || !psi) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    CU(cudaStreamSynchronize(E->stream));
    free_tables(E);
    free_plans(E);          // they point into the tables just freed
    E->mod_q.assign(q, q + n);
    E->mod_psi.assign(psi, psi + n);
    ++E->tf_version;
    // re-resolve the current modulus against the new ROM set
    E->mod_idx = -1;
    for (size_t i = 0; i < E->mod_q.size(); ++i) if (E->mod_q[i] == E->q) { E->mod_idx = (int)i; break; }
    return ALOHA_OK;
}

int aloha_dma_mem_h2d(aloha_t *E, uint32_t row, const uint64_t *src, uint64_t bytes) {
    if (!E || !src || bytes % 64) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    const u64 off = (u64)row * kLanes, n = bytes / 8;
    if (off + n > E->spm_words) return fail(E, ALOHA_E_RANGE, "DMA beyond SPM");
    int rc = cow_for_host_write(E, off, n);
    if (rc) return rc;
    if (!E->pending_down.empty() && (rc = wait_for_downloads(E, E->stream, off, n))) return rc;
    CU(cudaMemcpyAsync(E->d_spm + off, src, bytes, cudaMemcpyHostToDevice, E->stream));
    mark_written(E, off, n);
    return ALOHA_OK;
}

int aloha_dma_mem_h2d_async(aloha_t *E, uint32_t row, const uint64_t *src, uint64_t bytes) {
    if (!E || !src || bytes % 64) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    const u64 off = (u64)row * kLanes, n = bytes / 8;
    if (off + n > E->spm_words) return fail(E, ALOHA_E_RANGE, "DMA beyond SPM");
    int rc = cow_for_host_write(E, off, n);
    if (rc) return rc;
    if (!E->up_stream) {
        CU(cudaStreamCreateWithFlags(&E->up_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&E->down_stream, cudaStreamNonBlocking));
    }
    // the upload must not overtake kernels already queued (they may still read these rows) ...
    rc = order_after(E, E->up_stream, E->stream);
    if (rc) return rc;
    // ... nor a pending download of overlapping rows
    rc = wait_for_downloads(E, E->up_stream, off, n);
    if (rc) return rc;
    CU(cudaMemcpyAsync(E->d_spm + off, src, bytes, cudaMemcpyHostToDevice, E->up_stream));
    // every later run_vp sees the data
    rc = order_after(E, E->stream, E->up_stream);
    if (rc) return rc;
    mark_written(E, off, n);
    return ALOHA_OK;
}

int aloha_dma_mem_d2h_async(aloha_t *E, uint64_t *dst, uint32_t row, uint64_t bytes) {
    if (!E || !dst || bytes % 64) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    const u64 off = (u64)row * kLanes, n = bytes / 8;
    if (off + n > E->spm_words) return fail(E, ALOHA_E_RANGE, "DMA beyond SPM");
    if (!E->up_stream) {
        CU(cudaStreamCreateWithFlags(&E->up_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&E->down_stream, cudaStreamNonBlocking));
    }
    int rc = order_after(E, E->down_stream, E->stream);     // sees every run_vp issued so far
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst, E->d_spm + off, bytes, cudaMemcpyDeviceToHost, E->down_stream));
    aloha::PendingDma p{off, n, nullptr};
    rc = get_event(E, &p.done);
    if (rc) return rc;
    CU(cudaEventRecord(p.done, E->down_stream));
    // later work that WRITES these rows waits for this event (wait_for_downloads); everything else
    // keeps running beside the copy
    E->pending_down.push_back(p);
    return ALOHA_OK;
}

int aloha_dma_mem_d2h(aloha_t *E, uint64_t *dst, uint32_t row, uint64_t bytes) {
    if (!E || !dst || bytes % 64) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    const u64 off = (u64)row * kLanes, n = bytes / 8;
    if (off + n > E->spm_words) return fail(E, ALOHA_E_RANGE, "DMA beyond SPM");
    CU(cudaMemcpyAsync(dst, E->d_spm + off, bytes, cudaMemcpyDeviceToHost, E->stream));
    CU(cudaStreamSynchronize(E->stream));
    return ALOHA_OK;
}

int aloha_dma_ksk_h2d(aloha_t *E, uint32_t row, const uint64_t *src, uint64_t bytes) {
    if (!E || !src || bytes % 64) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    const u64 off = (u64)row * kLanes, n = bytes / 8;
    if (off + n > E->ksk_words) return fail(E, ALOHA_E_RANGE, "DMA beyond KSK memory");
    // registers aliasing KSK rows (VLE base 15) keep the OLD key: move them out first
    bool any = false;
    for (int r = 0; r < 32; ++r)
        if (E->loc[r].space == SP_KSK && E->loc[r].off < off + n && off < E->loc[r].off + E->loc[r].n) any = true;
    if (any) {
        Builder B(E);
        int rc = B.cow(SP_KSK, off, n, -1);
        Plan plan;
        if (!rc) rc = compile_plan(E, B, &plan);
        if (!rc) rc = execute_plan(E, plan);
        if (!rc) commit(E, plan);
        cudaStreamSynchronize(E->stream);
        cudaFree(plan.d_tables);
        if (rc) return rc;
    }
    CU(cudaMemcpyAsync(E->d_ksk + off, src, bytes, cudaMemcpyHostToDevice, E->stream));
    return ALOHA_OK;
}

int aloha_spm_written(aloha_t *E, uint32_t row, uint64_t nwords, uint8_t *out) {
    if (!E || !out) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    const u64 off = (u64)row * kLanes;
    if (off + nwords > E->spm_words) return fail(E, ALOHA_E_RANGE, "range beyond SPM");
    // one flag per 64-byte beat (8 words); `off` is row-aligned, hence beat-aligned
    const uint8_t *flags = E->written.data() + off / 8;
    const u64 beats = (nwords + 7) / 8;
    // the flags are 0 or 1 and come in long runs (a dump is typically all written, or written up to some polynomial):
    // one memchr + one memset per run
    for (u64 b = 0; b < beats;) {
        const uint8_t v = flags[b];
        const void *e = std::memchr(flags + b, v ? 0 : 1, beats - b);
        const u64 end = e ? (u64)((const uint8_t *)e - flags) : beats;
        const u64 first = 8 * b, last = std::min<u64>(8 * end, nwords);
        std::memset(out + first, v, last - first);
        b = end;
    }
    return ALOHA_OK;
}

int aloha_run_vp(aloha_t *E, uint32_t pc, uint32_t src0, uint32_t src1, uint32_t rslt, uint32_t ksk_ptr,
                 uint32_t step) {
    if (!E) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    const aloha_vp_args a{src0, src1, rslt, ksk_ptr, step};
    return enqueue_or_run(E, &pc, true, 1, &a);
}

int aloha_run_vp_batch(aloha_t *E, uint32_t pc, uint32_t count, const aloha_vp_args *args) {
    if (!E || (count && !args)) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    return enqueue_or_run(E, &pc, true, count, args);
}

int aloha_run_vp_multi(aloha_t *E, uint32_t count, const uint32_t *pcs, const aloha_vp_args *args) {
    if (!E || (count && (!args || !pcs))) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    return enqueue_or_run(E, pcs, false, count, args);
}

int aloha_sync(aloha_t *E) {
    if (!E) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    CU(cudaStreamSynchronize(E->stream));
    if (E->up_stream) {
        CU(cudaStreamSynchronize(E->up_stream));
        CU(cudaStreamSynchronize(E->down_stream));
    }
    for (auto &p : E->pending_down) put_event(E, p.done);
    E->pending_down.clear();
    return ALOHA_OK;
}

int aloha_pinned_alloc(uint64_t bytes, void **out) {
    if (!out || !bytes) return ALOHA_E_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? ALOHA_OK : ALOHA_E_NOMEM;
}
void aloha_pinned_free(void *p) { if (p) cudaFreeHost(p); }

int aloha_flush(aloha_t *E) {
    if (!E) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    return ALOHA_OK;
}

int aloha_spm_device_ptr(aloha_t *E, uint32_t row, void **p) {
    if (!E || !p) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    if (row >= E->cfg.spm_rows) return fail(E, ALOHA_E_RANGE, "row beyond SPM");
    *p = E->d_spm + (u64)row * kLanes;
    return ALOHA_OK;
}
int aloha_ksk_device_ptr(aloha_t *E, uint32_t row, void **p) {
    if (!E || !p) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    if (row >= E->cfg.ksk_rows) return fail(E, ALOHA_E_RANGE, "row beyond KSK memory");
    *p = E->d_ksk + (u64)row * kLanes;
    return ALOHA_OK;
}
int aloha_spm_mark_written(aloha_t *E, uint32_t row, uint32_t nrows) {
    if (!E) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    if ((u64)row + nrows > E->cfg.spm_rows) return fail(E, ALOHA_E_RANGE, "rows beyond SPM");
    int rc = cow_for_host_write(E, (u64)row * kLanes, (u64)nrows * kLanes);
    if (rc) return rc;
    mark_written(E, (u64)row * kLanes, (u64)nrows * kLanes);
    return ALOHA_OK;
}
int aloha_set_stream(aloha_t *E, void *stream) {
    if (!E) return ALOHA_E_ARG;
    DeviceGuard dg_(E->device);
    FLUSH();
    CU(cudaStreamSynchronize(E->stream));
    E->stream = stream ? (cudaStream_t)stream : E->own_stream;
    return ALOHA_OK;
}

int aloha_get_stats(const aloha_t *Ec, aloha_stats *out) {
    if (!Ec || !out) return ALOHA_E_ARG;
    aloha_t *E = const_cast<aloha_t *>(Ec);   // logically const: queued calls are part of the observable state
    DeviceGuard dg_(E->device);
    FLUSH();
    *out = E->stats;
    return ALOHA_OK;
}
int aloha_get_csr(const aloha_t *Ec, uint64_t *vl, uint64_t *q, uint64_t *iq) {
    if (!Ec) return ALOHA_E_ARG;
    aloha_t *E = const_cast<aloha_t *>(Ec);
    DeviceGuard dg_(E->device);
    FLUSH();
    if (vl) *vl = E->vl;
    if (q) *q = E->q;
    if (iq) *iq = E->iq;
    return ALOHA_OK;
}

int aloha_decode(const uint8_t word[12], uint64_t csr_step, uint64_t out[17]) {
    if (!word || !out) return ALOHA_E_ARG;
    const MicroOp m = expand(parse_word(word), csr_step);
    const uint64_t f[17] = {m.cfg, m.scalar_cfg, m.b0r, m.b0w, m.b1r, m.b1w, m.alu, m.scalar_alu, m.iconn,
                            m.scalar_iconn, m.ntt, m.muxo, m.muxi, m.vmu_cfg, m.vmu_scalar_cfg, m.ls, m.scalar_ls};
    std::memcpy(out, f, sizeof f);
    return ALOHA_OK;
}

}  // extern "C"
