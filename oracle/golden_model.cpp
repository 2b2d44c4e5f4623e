// oracle/golden_model.cpp -- CPU golden model of ALOHA's VP ISA.  TEST INFRASTRUCTURE ONLY
// (see golden_model.h for scope, citations and parity status).  Plain C++17, no dependencies.
//
// The model executes from the *decoded micro-op bundle* (the 17 fields the reference's
// expander emits), not from the mnemonic, so that bank selection by register parity and the
// operand muxes behave as the RTL's do even for operand combinations the shipped microcode
// never uses.
#include "golden_model.h"

#include <algorithm>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

typedef uint64_t u64;
typedef unsigned __int128 u128;

namespace {

constexpr u64 kLanes = 128;          // vp_defines.vh:25  SYS_NUM_LANE
constexpr u64 kIramDepthDefault = 4096;     // vp_defines.vh:31
constexpr unsigned kModWidth = 60;   // vxu_lane.sv:539  .mod_width(6'd60)

// ----------------------------------------------------------------------------- arithmetic
// modalu.sv:44-46 -- every ALU input gets exactly one conditional subtract.
inline u64 prered(u64 x, u64 q) { return x >= q ? x - q : x; }

// modmul.sv:150 (prod >> mod_width-2), :172 (mid >> mod_width+3), :195 (mod_mask = 1<<mod_width+1),
// :216-252 (masked subtract, final conditional subtract).  All intermediates truncate to 64 bits
// exactly where the RTL's `logic [data_width_p-1:0]` nets do.
inline u64 barrett(u64 a, u64 b, u64 q, u64 iq) {
    u128 prod = (u128)a * b;
    u64 prod_shift = (u64)(prod >> (kModWidth - 2));
    u128 mid = (u128)prod_shift * iq;
    u64 mid_shift = (u64)(mid >> (kModWidth + 3));
    u128 estim = (u128)mid_shift * q;
    const u64 mask = 1ull << (kModWidth + 1);
    u64 dx = (u64)prod & (mask - 1);
    u64 dy = (u64)estim & (mask - 1);
    u64 diff = ((dx | mask) - dy) & (mask - 1);
    return diff < q ? diff : diff - q;
}

// modalu.sv:228-229 -- 65-bit sum, one conditional subtract, truncated to 64 bits.
inline u64 addmod(u64 a, u64 b, u64 q) {
    u128 s = (u128)a + b;
    return (u64)(s >= q ? s - q : s);
}
// modalu.sv:249
inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : q + a - b; }
// halfred.sv:23-26
inline u64 halfmod(u64 x, u64 q) { return (x >> 1) + ((x & 1) ? ((q + 1) >> 1) : 0); }

enum AluOp : uint32_t {  // modalu.sv:22-37
    MUL_VV = 0x00, MUL_VS = 0x04, ADD_VV = 0x01, ADD_VS = 0x05, SUB_VV = 0x02, SUB_VS = 0x06,
    SUB_SV = 0x0a, MOD_V = 0x03, MADD_VS = 0x15, MSUB_VS = 0x16, MSUB_SV = 0x1a, CT_VVS = 0x10,
    GS_VVS = 0x13, VVS = 0x11
};

inline u64 alu(uint32_t op, u64 a_raw, u64 b_raw, u64 s_raw, u64 q, u64 iq, u64 *res1) {
    const u64 a = prered(a_raw, q), b = prered(b_raw, q), s = prered(s_raw, q);
    u64 r1 = 0, r0 = 0;
    switch (op) {
    case MUL_VV: r0 = barrett(a, b, q, iq); break;
    case MUL_VS: r0 = barrett(a, s, q, iq); break;
    case MOD_V: r0 = barrett(a, 1, q, iq); break;
    case ADD_VV: r0 = addmod(a, b, q); break;
    case ADD_VS: r0 = addmod(a, s, q); break;
    case SUB_VV: r0 = submod(a, b, q); break;
    case SUB_VS: r0 = submod(a, s, q); break;
    case SUB_SV: r0 = submod(s, a, q); break;
    case MADD_VS: r0 = addmod(barrett(a, b, q, iq), s, q); break;
    case MSUB_VS: r0 = submod(barrett(a, b, q, iq), s, q); break;
    case MSUB_SV: r0 = submod(s, barrett(a, b, q, iq), q); break;
    case VVS: r0 = barrett(submod(a, b, q), s, q, iq); break;
    case CT_VVS: {  // modalu.sv:160-165 (mul_opa = opb), :223-249
        u64 m = barrett(b, s, q, iq);
        r0 = addmod(a, m, q);
        r1 = submod(a, m, q);
        break;
    }
    case GS_VVS: {  // modalu.sv:152-155 (gs_subred), :296-327 (both outputs halved)
        r0 = halfmod(addmod(a, b, q), q);
        r1 = halfmod(barrett(submod(a, b, q), s, q, iq), q);
        break;
    }
    default: r0 = 0; break;  // modalu.sv:368-370
    }
    if (res1) *res1 = r1;
    return r0;
}

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

inline u64 bitrev(u64 x, unsigned bits) {
    u64 r = 0;
    for (unsigned i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

inline unsigned ilog2(u64 x) { unsigned l = 0; while ((1ull << l) < x) ++l; return l; }

// tf_rom_generator.sv:28-30,111 -- tw[j] = psi^bitrev(j, logN); :61-63,147-148 same with psi^-1.
void build_twiddles(std::vector<u64> &tw, u64 n, u64 root, u64 q) {
    const unsigned logn = ilog2(n);
    tw.assign(n, 0);
    u64 cur = 1;
    for (u64 e = 0; e < n; ++e) {
        tw[bitrev(e, logn)] = cur;
        cur = (u64)((u128)cur * root % q);
    }
}

struct Twiddles { std::vector<u64> fwd, inv; };

// Constant-geometry (Pease) schedule: ntt_fsm.sv:49-81 + ntt_swap.sv:35-51 + iconn_shuffle.sv:33.
// Net effect per stage s: out[2p] = x[p] + w x[p+N/2], out[2p+1] = x[p] - w x[p+N/2],
// w = tw[2^s + (p mod 2^s)].  Buffers ping-pong (ntt_fsm.sv:80): stage 0 reads A writes B, ...
// Returns which buffer holds the result: 0 = A, 1 = B.
int ntt_pease(u64 *A, u64 *B, u64 n, u64 q, u64 iq, const u64 *tw) {
    const unsigned logn = ilog2(n);
    const u64 h = n / 2;
    u64 *src = A, *dst = B;
    for (unsigned s = 0; s < logn; ++s) {
        const u64 m = 1ull << s;
        for (u64 p = 0; p < h; ++p) {
            u64 r1;
            u64 r0 = alu(CT_VVS, src[p], src[p + h], tw[m + (p & (m - 1))], q, iq, &r1);
            dst[2 * p] = r0;
            dst[2 * p + 1] = r1;
        }
        std::swap(src, dst);
    }
    return src == A ? 0 : 1;
}

// Mirror image: rows in reverse, GS butterfly with halving (modalu.sv:296-327), reverse shuffle.
int intt_pease(u64 *A, u64 *B, u64 n, u64 q, u64 iq, const u64 *itw) {
    const unsigned logn = ilog2(n);
    const u64 h = n / 2;
    u64 *src = A, *dst = B;
    for (unsigned s = 0; s < logn; ++s) {
        const u64 m = 1ull << (logn - 1 - s);
        for (u64 p = 0; p < h; ++p) {
            u64 r1;
            u64 r0 = alu(GS_VVS, src[2 * p], src[2 * p + 1], itw[m + (p & (m - 1))], q, iq, &r1);
            dst[p] = r0;
            dst[p + h] = r1;
        }
        std::swap(src, dst);
    }
    return src == A ? 0 : 1;
}

// ----------------------------------------------------------------------------- decoder
struct Bundle {  // order of seq_top_tb.sv:138-160
    u64 cfg = 0, scalar_cfg = 0, b0r = 0, b0w = 0, b1r = 0, b1w = 0, alu = 0, scalar_alu = 0,
        iconn = 0, scalar_iconn = 0, ntt = 0, muxo = 0, muxi = 0, vmu_cfg = 0, vmu_scalar_cfg = 0,
        ls = 0, scalar_ls = 0;
    bool brk = false;
    unsigned funct6 = 0, funct3 = 0;
};

enum Funct6 : unsigned {  // expander.v:65-81
    F_VL = 0x04, F_MODQ = 0x08, F_MODIQ = 0x0c, F_BREAK = 0x10, F_NOP = 0x00, F_FQMUL = 0x01,
    F_FQADD = 0x05, F_FQSUB = 0x09, F_FQMOD = 0x0d, F_VCP = 0x11, F_VAUT = 0x15, F_ROLI = 0x19,
    F_NTT = 0x02, F_INTT = 0x06, F_VLE = 0x03, F_VSE = 0x07
};

inline u64 en(unsigned reg) { return ((u64)reg << 1) | 1; }  // {reg, 1'b1}

Bundle decode(const uint8_t w[12], u64 csr_step) {
    uint32_t inst = ((uint32_t)w[0] << 24) | ((uint32_t)w[1] << 16) | ((uint32_t)w[2] << 8) | w[3];
    u64 imm = 0;
    for (int i = 4; i < 12; ++i) imm = (imm << 8) | w[i];
    // expander.v:123-130
    const unsigned f6 = inst >> 26, vs2 = (inst >> 20) & 31, vs1 = (inst >> 15) & 31,
                   f3 = (inst >> 12) & 7, vd = (inst >> 7) & 31;
    Bundle b;
    b.funct6 = f6;
    b.funct3 = f3;
    b.brk = f6 == F_BREAK;  // expander.v:141-151
    // expander.v:154-176  (the VMU sees the same config op and scalar)
    if (f6 == F_VL) b.cfg = 1;
    else if (f6 == F_MODQ) b.cfg = 2;
    else if (f6 == F_MODIQ) b.cfg = 3;
    if (b.cfg) { b.scalar_cfg = imm; b.vmu_cfg = b.cfg; b.vmu_scalar_cfg = imm; }

    // expander.v:178-532 -- bank reads, ALU opcode, scalar, output mux
    auto vv = [&](u64 op) {
        if ((vs1 & 1) == 0) { b.b0r = en(vs1); b.b1r = en(vs2); b.muxo = 0x4; }
        else { b.b0r = en(vs2); b.b1r = en(vs1); b.muxo = 0x8; }
        b.alu = op; b.scalar_alu = 0;
    };
    auto vs = [&](u64 op, unsigned reg, u64 scalar) {
        if ((reg & 1) == 0) { b.b0r = en(reg); b.b1r = 0; b.muxo = 0x4; }
        else { b.b0r = 0; b.b1r = en(reg); b.muxo = 0x8; }
        b.alu = op; b.scalar_alu = scalar;
    };
    switch (f6) {
    case F_FQMUL:
        if (f3 == 0) vv(MUL_VV); else if (f3 == 1) vs(MUL_VS, vs1, imm); else b.alu = MUL_VV;
        break;
    case F_FQADD:
        if (f3 == 0) vv(ADD_VV); else if (f3 == 1) vs(ADD_VS, vs1, imm); else b.alu = ADD_VV;
        break;
    case F_FQSUB:  // .sv reads its vector operand from vs2 (expander.v:342-363)
        if (f3 == 0) vv(SUB_VV); else if (f3 == 1) vs(SUB_VS, vs1, imm);
        else if (f3 == 2) vs(SUB_SV, vs2, imm); else b.alu = SUB_VV;
        break;
    case F_FQMOD: vs(MOD_V, vs1, 0); break;
    case F_VCP: vs(ADD_VS, vs1, 0); break;  // VCPY = addmod(r(x), 0)   expander.v:396-417
    case F_NTT:
        vs(CT_VVS, vs1, 0); b.muxo = (vs1 & 1) ? 0x2 : 0x0; break;
    case F_INTT:
        vs(GS_VVS, vs1, 0); b.muxo = (vs1 & 1) ? 0x8 : 0x0; break;
    case F_VAUT:
    case F_ROLI:
        vs(MUL_VV, vs1, 0); b.muxo = (vs1 & 1) ? 0x2 : 0x0; break;
    case F_VSE:
        vs(MUL_VV, vs1, 0); b.muxo = (vs1 & 1) ? 0x1 : 0x0; break;
    default: break;
    }
    // expander.v:533-558
    if (f6 == F_NTT) b.iconn = 4;
    else if (f6 == F_INTT) b.iconn = 5;
    else if (f6 == F_VAUT) { b.iconn = 1; b.scalar_iconn = csr_step + imm; }
    else if (f6 == F_ROLI) { b.iconn = 2; b.scalar_iconn = imm; }
    // expander.v:559-578
    if (f6 == F_NTT) b.ntt = 2; else if (f6 == F_INTT) b.ntt = 3;
    // expander.v:579-668 -- bank writes + input mux, keyed on funct6[1:0]
    auto wr = [&](u64 mi_even, u64 mi_odd) {
        if ((vd & 1) == 0) { b.b0w = en(vd); b.b1w = 0; b.muxi = mi_even; }
        else { b.b0w = 0; b.b1w = en(vd); b.muxi = mi_odd; }
    };
    switch (f6 & 3) {
    case 1: {
        const bool perm = f6 == F_ROLI || f6 == F_VAUT;
        wr(perm ? 0x4 : 0x0, perm ? 0x1 : 0x0);
        break;
    }
    case 2:
        if (f6 == F_NTT) wr(0x0, 0x0); else if (f6 == F_INTT) wr(0x4, 0x1);
        break;
    case 3:
        if (f6 == F_VLE) wr(0xc, 0x3);
        break;
    default: break;
    }
    // expander.v:669-698
    if (f6 == F_VLE) { b.ls = 1; b.scalar_ls = imm; }
    else if (f6 == F_VSE) { b.ls = 2; b.scalar_ls = imm; }
    return b;
}

}  // namespace

// ----------------------------------------------------------------------------- machine state
struct gm {
    u64 vlmax_bits, nmax;
    unsigned kbits;                       // $clog2(NELEMENT*NLANE)   vxu_lane.sv:594
    uint32_t spm_rows, ksk_rows;
    std::vector<u64> spm, ksk;
    std::vector<uint8_t> spm_written;
    std::vector<std::vector<u64>> vreg;   // 32 registers: bank = reg & 1, index = reg >> 1
    u64 iram_depth = kIramDepthDefault;
    std::vector<uint8_t> isram;           // iram_depth x 12
    std::vector<uint8_t> isram_valid;
    u64 vl = 0, q = 0, iq = 0;            // persist across run_vp
    int tf_item = -1;
    std::vector<u64> mod_q, mod_psi;
    std::map<std::pair<int, u64>, Twiddles> tw_cache;
    std::vector<u64> scratch;
    uint32_t last_count = 0;

    const Twiddles *twiddles(u64 n) {
        if (tf_item < 0) return nullptr;
        auto key = std::make_pair(tf_item, n);
        auto it = tw_cache.find(key);
        if (it != tw_cache.end()) return &it->second;
        const u64 tq = mod_q[tf_item];
        if ((tq - 1) % (2 * n)) return nullptr;
        // ROM holds psi for Nmax; a shorter VL uses psi^(Nmax/N) (still minimal-order 2N root).
        u64 psi = powmod(mod_psi[tf_item], nmax / n, tq);
        u64 ipsi = powmod(psi, tq - 2, tq);
        Twiddles t;
        build_twiddles(t.fwd, n, psi, tq);
        build_twiddles(t.inv, n, ipsi, tq);
        return &(tw_cache[key] = std::move(t));
    }
};

namespace {

int exec_bundle(gm *m, const Bundle &b, uint32_t src0, uint32_t src1, uint32_t rslt,
                uint32_t ksk_ptr) {
    if (b.cfg == 1) {  // seq_top.v:417-429
        u64 n = b.scalar_cfg / 64;
        if (b.scalar_cfg % 64 || n < 2 * kLanes || (n & (n - 1)) || n > m->nmax) return GM_E_STATE;
        m->vl = b.scalar_cfg;
        return GM_OK;
    }
    if (b.cfg == 2) {  // vxu_top.sv:112-118: twiddle set by value match, else the last table
        m->q = b.scalar_cfg;
        m->tf_item = -1;
        for (size_t i = 0; i < m->mod_q.size(); ++i)
            if (m->mod_q[i] == m->q) { m->tf_item = (int)i; break; }
        if (m->tf_item < 0 && !m->mod_q.empty()) m->tf_item = (int)m->mod_q.size() - 1;
        return GM_OK;
    }
    if (b.cfg == 3) { m->iq = b.scalar_cfg; return GM_OK; }
    if (b.brk || b.funct6 == F_NOP) return GM_OK;
    const bool known = b.funct6 == F_FQMUL || b.funct6 == F_FQADD || b.funct6 == F_FQSUB ||
                       b.funct6 == F_FQMOD || b.funct6 == F_VCP || b.funct6 == F_VAUT ||
                       b.funct6 == F_ROLI || b.funct6 == F_NTT || b.funct6 == F_INTT ||
                       b.funct6 == F_VLE || b.funct6 == F_VSE;
    if (!known) return GM_E_OPCODE;
    if (!m->vl) return GM_E_STATE;
    const u64 n = m->vl / 64, rows = n / kLanes;

    // physical registers behind the bank ports (vxu_lane.sv:321-328; bank = reg&1, index = reg>>1)
    auto bank_reg = [](u64 field, unsigned bank) -> int {
        return (field & 1) ? (int)((((field >> 1) & 31) & ~1u) | bank) : -1;
    };
    const int r0 = bank_reg(b.b0r, 0), r1 = bank_reg(b.b1r, 1);
    const int w0 = bank_reg(b.b0w, 0), w1 = bank_reg(b.b1w, 1);
    const int wd = w0 >= 0 ? w0 : w1;
    auto rd = [&](unsigned sel_bit) -> int { return ((b.muxo >> sel_bit) & 1) ? r1 : r0; };

    // ---- VLE / VSE: vp_top_full.sv:105-117, addr_gen.v:44
    if (b.ls) {
        const unsigned sel = (unsigned)(b.scalar_ls >> 48);
        const u64 off = (b.scalar_ls >> 10) & 0xffff;
        if (b.ls == 1) {
            if (wd < 0) return GM_E_OPCODE;
            const bool from_ksk = sel == 15;
            const u64 base = from_ksk ? ksk_ptr : sel == 0 ? src0 : sel == 1 ? src1 : sel == 2 ? rslt : 0;
            const std::vector<u64> &mem = from_ksk ? m->ksk : m->spm;
            const u64 limit = from_ksk ? m->ksk_rows : m->spm_rows;
            for (u64 c = 0; c < rows; ++c) {
                const u64 row = base + ((off + c) & 0xffff);
                if (row >= limit) return GM_E_RANGE;
                std::memcpy(&m->vreg[wd][c * kLanes], &mem[row * kLanes], kLanes * 8);
            }
        } else {
            const int rs = rd(0);
            if (rs < 0) return GM_E_OPCODE;
            const u64 base = sel == 0 ? src0 : sel == 1 ? src1 : sel == 2 ? rslt : 0;
            for (u64 c = 0; c < rows; ++c) {
                const u64 row = base + ((off + c) & 0xffff);
                if (row >= m->spm_rows) return GM_E_RANGE;
                std::memcpy(&m->spm[row * kLanes], &m->vreg[rs][c * kLanes], kLanes * 8);
                std::memset(&m->spm_written[row * kLanes], 1, kLanes);
            }
        }
        return GM_OK;
    }
    if (!m->q) return GM_E_STATE;
    if (wd < 0) return GM_E_OPCODE;

    // ---- VNTT / VINTT
    if (b.ntt == 2 || b.ntt == 3) {
        const int rs = b.ntt == 2 ? rd(1) : rd(3);
        if (rs < 0) return GM_E_OPCODE;
        if (rs == wd) return GM_E_ILLEGAL;
        const Twiddles *t = m->twiddles(n);
        if (!t) return GM_E_STATE;
        u64 *A = m->vreg[rs].data(), *B = m->vreg[wd].data();
        int where = b.ntt == 2 ? ntt_pease(A, B, n, m->q, m->iq, t->fwd.data())
                               : intt_pease(A, B, n, m->q, m->iq, t->inv.data());
        // Odd logN (the RTL's N=8192): result lands in vd, vs1 keeps stage logN-2.
        // Even logN: the RTL would leave them the other way round; this model defines vd = result.
        if (where == 0) for (u64 i = 0; i < n; ++i) std::swap(A[i], B[i]);
        return GM_OK;
    }

    // ---- VAUT / VROLI: vxu_lane.sv:594-599
    if (b.iconn == 1 || b.iconn == 2) {
        const int rs = rd(1);
        if (rs < 0) return GM_E_OPCODE;
        if (rs == wd) return GM_E_ILLEGAL;
        const u64 *x = m->vreg[rs].data();
        u64 *d = m->vreg[wd].data();
        if (b.iconn == 1) {
            const u64 k = b.scalar_iconn & ((1ull << m->kbits) - 1);
            if (!(k & 1)) return GM_E_ILLEGAL;  // even k: lane conflicts in iconn_top, no defined result
            for (u64 i = 0; i < n; ++i) {
                const u64 t = i * k;
                d[t & (n - 1)] = ((t & (2 * n - 1)) >= n) ? m->q - x[i] : x[i];
            }
        } else {
            for (u64 i = 0; i < n; ++i) d[(i + n - b.scalar_iconn) & (n - 1)] = x[i];
        }
        return GM_OK;
    }

    // ---- element-wise ALU
    const int ra = rd(3), rb = rd(2);
    u64 *d = m->vreg[wd].data();
    for (u64 i = 0; i < n; ++i) {
        const u64 a = ra >= 0 ? m->vreg[ra][i] : 0;
        const u64 bb = rb >= 0 ? m->vreg[rb][i] : 0;
        d[i] = alu((uint32_t)b.alu, a, bb, b.scalar_alu, m->q, m->iq, nullptr);
    }
    return GM_OK;
}

}  // namespace

static void run_threads(uint32_t nthreads, u64 count, const std::function<void(u64, u64)> &fn) {
    if (nthreads <= 1 || count <= 1) { fn(0, count); return; }
    nthreads = (uint32_t)std::min<u64>(nthreads, count);
    std::vector<std::thread> th;
    for (uint32_t t = 0; t < nthreads; ++t) {
        u64 lo = count * t / nthreads, hi = count * (t + 1) / nthreads;
        th.emplace_back([=, &fn] { fn(lo, hi); });
    }
    for (auto &x : th) x.join();
}

// ----------------------------------------------------------------------------- C API
extern "C" {

gm_t *gm_create(uint64_t vlmax_bits, uint32_t spm_rows, uint32_t ksk_rows) {
    u64 nmax = vlmax_bits / 64;
    if (vlmax_bits % 64 || nmax < 2 * kLanes || (nmax & (nmax - 1))) return nullptr;
    gm *m = new gm();
    m->vlmax_bits = vlmax_bits;
    m->nmax = nmax;
    m->kbits = ilog2(nmax);
    m->spm_rows = spm_rows;
    m->ksk_rows = ksk_rows;
    m->spm.assign((size_t)spm_rows * kLanes, 0);
    m->spm_written.assign((size_t)spm_rows * kLanes, 0);
    m->ksk.assign((size_t)ksk_rows * kLanes, 0);
    m->vreg.assign(32, std::vector<u64>(nmax, 0));
    m->isram.assign(m->iram_depth * 12, 0);
    m->isram_valid.assign(m->iram_depth, 0);
    return m;
}

void gm_destroy(gm_t *m) { delete m; }

int gm_set_moduli(gm_t *m, const uint64_t *q, const uint64_t *psi, uint32_t n) {
    if (!m || !q || !psi) return GM_E_ARG;
    m->mod_q.assign(q, q + n);
    m->mod_psi.assign(psi, psi + n);
    m->tw_cache.clear();
    return GM_OK;
}

int gm_load_isram(gm_t *m, const uint8_t *words, uint32_t n, uint32_t at_pc) {
    if (!m || !words) return GM_E_ARG;
    if ((u64)at_pc + n > m->iram_depth) {   // IRAM_DEPTH is an elaboration parameter: grow on demand
        m->iram_depth = (u64)at_pc + n;
        m->isram.resize(m->iram_depth * 12, 0);
        m->isram_valid.resize(m->iram_depth, 0);
    }
    std::memcpy(&m->isram[(size_t)at_pc * 12], words, (size_t)n * 12);
    std::memset(&m->isram_valid[at_pc], 1, n);
    return GM_OK;
}

// DMA_CMD_MEM: linear u64 copy, DRAM word w -> SPM row w/128, lane w%128 (spm.sv:92-126).
int gm_dma_mem_h2d(gm_t *m, uint32_t row, const uint64_t *src, uint64_t bytes) {
    if (!m || !src || bytes % 8) return GM_E_ARG;
    if ((u64)row * kLanes + bytes / 8 > m->spm.size()) return GM_E_RANGE;
    std::memcpy(&m->spm[(size_t)row * kLanes], src, bytes);
    std::memset(&m->spm_written[(size_t)row * kLanes], 1, bytes / 8);
    return GM_OK;
}
int gm_dma_mem_d2h(gm_t *m, uint64_t *dst, uint32_t row, uint64_t bytes) {
    if (!m || !dst || bytes % 8) return GM_E_ARG;
    if ((u64)row * kLanes + bytes / 8 > m->spm.size()) return GM_E_RANGE;
    std::memcpy(dst, &m->spm[(size_t)row * kLanes], bytes);
    return GM_OK;
}
int gm_dma_ksk_h2d(gm_t *m, uint32_t row, const uint64_t *src, uint64_t bytes) {
    if (!m || !src || bytes % 8) return GM_E_ARG;
    if ((u64)row * kLanes + bytes / 8 > m->ksk.size()) return GM_E_RANGE;
    std::memcpy(&m->ksk[(size_t)row * kLanes], src, bytes);
    return GM_OK;
}
int gm_spm_written(gm_t *m, uint32_t row, uint64_t nwords, uint8_t *out) {
    if (!m || !out) return GM_E_ARG;
    if ((u64)row * kLanes + nwords > m->spm.size()) return GM_E_RANGE;
    std::memcpy(out, &m->spm_written[(size_t)row * kLanes], nwords);
    return GM_OK;
}

// top_noaxilite_tb.sv:396-417 (CSRs + start + poll done); seq_top.v:170-248 (PC FSM), :532-543 (BREAK)
int gm_run_vp(gm_t *m, uint32_t pc, uint32_t src0, uint32_t src1, uint32_t rslt, uint32_t ksk_ptr,
              uint32_t step) {
    if (!m) return GM_E_ARG;
    m->last_count = 0;
    for (u64 at = pc;; ++at) {
        if (at >= m->iram_depth) return GM_E_NOBREAK;
        Bundle b = decode(&m->isram[at * 12], step);
        ++m->last_count;
        int rc = exec_bundle(m, b, src0, src1, rslt, ksk_ptr);
        if (rc) return rc;
        if (b.brk) return GM_OK;
    }
}
uint32_t gm_last_inst_count(const gm_t *m) { return m ? m->last_count : 0; }

int gm_vreg_read(gm_t *m, uint32_t reg, uint64_t *dst, uint64_t nwords) {
    if (!m || reg > 31 || nwords > m->nmax) return GM_E_ARG;
    std::memcpy(dst, m->vreg[reg].data(), nwords * 8);
    return GM_OK;
}
int gm_vreg_write(gm_t *m, uint32_t reg, const uint64_t *src, uint64_t nwords) {
    if (!m || reg > 31 || nwords > m->nmax) return GM_E_ARG;
    std::memcpy(m->vreg[reg].data(), src, nwords * 8);
    return GM_OK;
}
int gm_get_csr(const gm_t *m, uint64_t *vl, uint64_t *q, uint64_t *iq) {
    if (!m) return GM_E_ARG;
    if (vl) *vl = m->vl;
    if (q) *q = m->q;
    if (iq) *iq = m->iq;
    return GM_OK;
}

int gm_decode(const uint8_t word[12], uint64_t csr_step, uint64_t out[17]) {
    if (!word || !out) return GM_E_ARG;
    Bundle b = decode(word, csr_step);
    const u64 f[17] = {b.cfg, b.scalar_cfg, b.b0r, b.b0w, b.b1r, b.b1w, b.alu, b.scalar_alu, b.iconn,
                       b.scalar_iconn, b.ntt, b.muxo, b.muxi, b.vmu_cfg, b.vmu_scalar_cfg, b.ls,
                       b.scalar_ls};
    std::memcpy(out, f, sizeof f);
    return GM_OK;
}

uint64_t gm_barrett(uint64_t a, uint64_t b, uint64_t q, uint64_t iq) { return barrett(a, b, q, iq); }
uint64_t gm_half(uint64_t x, uint64_t q) { return halfmod(x, q); }
uint64_t gm_alu(uint32_t op, uint64_t a, uint64_t b, uint64_t s, uint64_t q, uint64_t iq,
                uint64_t *res1) {
    return alu(op, a, b, s, q, iq, res1);
}

uint64_t gm_barrett_iq(uint64_t q) { return (u64)((((u128)1) << 121) / q); }
uint64_t gm_powmod(uint64_t a, uint64_t e, uint64_t q) { return powmod(a, e, q); }

uint64_t gm_min_primitive_root(uint64_t q, uint64_t two_n) {
    if (two_n < 2 || (q - 1) % two_n) return 0;
    const u64 n = two_n / 2;
    u64 root = 0;
    for (u64 g = 2; g < 1000; ++g) {
        u64 r = powmod(g, (q - 1) / two_n, q);
        if (powmod(r, n, q) == q - 1) { root = r; break; }
    }
    if (!root) return 0;
    // every primitive 2n-th root is root^odd; take the smallest
    const u64 sq = (u64)((u128)root * root % q);
    u64 best = root, cur = root;
    for (u64 j = 1; j < n; ++j) {
        cur = (u64)((u128)cur * sq % q);
        best = std::min(best, cur);
    }
    return best;
}

int gm_ntt(uint64_t *a, uint64_t *scratch, uint64_t n, uint64_t q, uint64_t iq, uint64_t psi,
           int inverse) {
    if (!a || !scratch || n < 2 || (n & (n - 1)) || (q - 1) % (2 * n)) return GM_E_ARG;
    std::vector<u64> tw;
    build_twiddles(tw, n, inverse ? powmod(psi, q - 2, q) : psi, q);
    int where = inverse ? intt_pease(a, scratch, n, q, iq, tw.data())
                        : ntt_pease(a, scratch, n, q, iq, tw.data());
    if (where == 1) std::memcpy(a, scratch, n * 8);
    return GM_OK;
}

struct gm_ntt_tables {
    u64 n;
    std::vector<u64> q, iq;
    std::vector<Twiddles> tw;
};

gm_ntt_tables_t *gm_ntt_tables_create(uint64_t n, const uint64_t *q, const uint64_t *psi,
                                      uint32_t n_moduli) {
    if (!q || !psi || n < 2 || (n & (n - 1))) return nullptr;
    auto *t = new gm_ntt_tables();
    t->n = n;
    for (uint32_t i = 0; i < n_moduli; ++i) {
        if ((q[i] - 1) % (2 * n)) { delete t; return nullptr; }
        t->q.push_back(q[i]);
        t->iq.push_back(gm_barrett_iq(q[i]));
        Twiddles tw;
        build_twiddles(tw.fwd, n, psi[i], q[i]);
        build_twiddles(tw.inv, n, powmod(psi[i], q[i] - 2, q[i]), q[i]);
        t->tw.push_back(std::move(tw));
    }
    return t;
}
void gm_ntt_tables_destroy(gm_ntt_tables_t *t) { delete t; }

int gm_ntt_batch(const gm_ntt_tables_t *t, uint64_t *a, const uint32_t *mod_idx, uint64_t count,
                 int inverse, uint32_t nthreads) {
    if (!t || !a || !mod_idx) return GM_E_ARG;
    for (u64 j = 0; j < count; ++j) if (mod_idx[j] >= t->q.size()) return GM_E_ARG;
    const u64 n = t->n;
    run_threads(nthreads, count, [&](u64 lo, u64 hi) {
        std::vector<u64> scratch(n);
        for (u64 j = lo; j < hi; ++j) {
            const uint32_t mi = mod_idx[j];
            u64 *p = a + j * n;
            int where = inverse ? intt_pease(p, scratch.data(), n, t->q[mi], t->iq[mi], t->tw[mi].inv.data())
                                : ntt_pease(p, scratch.data(), n, t->q[mi], t->iq[mi], t->tw[mi].fwd.data());
            if (where == 1) std::memcpy(p, scratch.data(), n * 8);
        }
    });
    return GM_OK;
}

int gm_automorph(uint64_t *dst, const uint64_t *src, uint64_t n, uint64_t k, uint64_t q) {
    if (!dst || !src || dst == src || (n & (n - 1)) || !(k & 1)) return GM_E_ARG;
    for (u64 i = 0; i < n; ++i) {
        const u64 t = i * k;
        dst[t & (n - 1)] = ((t & (2 * n - 1)) >= n) ? q - src[i] : src[i];
    }
    return GM_OK;
}

int gm_aut_mac_batch(uint64_t *acc, const uint64_t *x, const uint64_t *p, uint64_t n, uint64_t k,
                     const uint64_t *q, const uint32_t *mod_idx, uint64_t count,
                     uint32_t nthreads) {
    if (!acc || !x || !p || !q || !mod_idx || (n & (n - 1)) || !(k & 1)) return GM_E_ARG;
    run_threads(nthreads, count, [&](u64 lo, u64 hi) {
        std::vector<u64> tmp(n);
        for (u64 j = lo; j < hi; ++j) {
            const u64 qq = q[mod_idx[j]], iq = gm_barrett_iq(qq);
            gm_automorph(tmp.data(), x + j * n, n, k, qq);
            u64 *ac = acc + j * n;
            const u64 *pp = p + j * n;
            for (u64 i = 0; i < n; ++i) {
                u64 prod = alu(MUL_VV, tmp[i], pp[i], 0, qq, iq, nullptr);
                ac[i] = alu(ADD_VV, ac[i], prod, 0, qq, iq, nullptr);
            }
        }
    });
    return GM_OK;
}

}  // extern "C"

