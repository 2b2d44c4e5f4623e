// isa.hpp -- the 96-bit R-type HE instruction word and its decode (host side).
//
// Word layout (reference: src/vp/sequncer/expander.v:123-130; vp_defines.vh:20):
//   [95:64] inst32 = funct6[31:26] | m[25] | vs2[24:20] | vs1[19:15] | funct3[14:12] | vd[11:7] | opcode[6:0]
//   [63:0]  imm64
// The decoder is table-driven: one row per funct6 describes which register feeds which bank port
// and how the micro-op fields (the 17 columns the reference's sequencer testbench compares,
// sim/vp/sequncer/seq_top_tb.sv:138-160) are filled.
#pragma once
#include <cstdint>

namespace alb {

enum Funct6 : uint8_t {   // expander.v:65-81
    F6_NOP = 0x00, F6_FQMUL = 0x01, F6_NTT = 0x02, F6_VLE = 0x03, F6_VSETVL = 0x04, F6_FQADD = 0x05,
    F6_INTT = 0x06, F6_VSE = 0x07, F6_VSETQ = 0x08, F6_FQSUB = 0x09, F6_VSETIQ = 0x0c,
    F6_FQMOD = 0x0d, F6_BREAK = 0x10, F6_VCPY = 0x11, F6_VAUT = 0x15, F6_VROLI = 0x19
};

enum AluCode : uint8_t {  // modalu.sv:22-37 / expander.v:96-109
    A_MULVV = 0x00, A_ADDVV = 0x01, A_SUBVV = 0x02, A_MOD = 0x03, A_MULVS = 0x04, A_ADDVS = 0x05,
    A_SUBVS = 0x06, A_SUBSV = 0x0a, A_CT = 0x10, A_GS = 0x13
};

struct Inst {
    uint8_t funct6, funct3, vd, vs1, vs2, mask;
    uint64_t imm;
};

inline Inst parse_word(const uint8_t w[12]) {
    const uint32_t hi = (uint32_t(w[0]) << 24) | (uint32_t(w[1]) << 16) | (uint32_t(w[2]) << 8) | w[3];
    uint64_t imm = 0;
    for (int i = 4; i < 12; ++i) imm = (imm << 8) | w[i];
    Inst in;
    in.funct6 = uint8_t(hi >> 26);
    in.mask = uint8_t((hi >> 25) & 1);
    in.vs2 = uint8_t((hi >> 20) & 31);
    in.vs1 = uint8_t((hi >> 15) & 31);
    in.funct3 = uint8_t((hi >> 12) & 7);
    in.vd = uint8_t((hi >> 7) & 31);
    in.imm = imm;
    return in;
}

// The micro-op bundle, field order of seq_top_tb.sv:138-160.
struct MicroOp {
    uint64_t cfg, scalar_cfg, b0r, b0w, b1r, b1w, alu, scalar_alu, iconn, scalar_iconn, ntt, muxo,
        muxi, vmu_cfg, vmu_scalar_cfg, ls, scalar_ls;
};

namespace detail {
// How an instruction class sources its vector operand(s)
enum SrcForm : uint8_t { SRC_NONE, SRC_VV, SRC_VS1, SRC_VS2 };
struct ClassRow {
    SrcForm src;
    uint8_t alu;
    bool scalar_is_imm;
    uint8_t muxo_even, muxo_odd;   // output-mux code by parity of the (first) source register
};
}  // namespace detail

inline MicroOp expand(const Inst &in, uint64_t csr_step) {
    using namespace detail;
    MicroOp m{};
    const uint8_t f6 = in.funct6, f3 = in.funct3;

    // config ops: the VXU and the VMU both latch them (expander.v:154-176)
    const uint64_t cfg = f6 == F6_VSETVL ? 1 : f6 == F6_VSETQ ? 2 : f6 == F6_VSETIQ ? 3 : 0;
    if (cfg) {
        m.cfg = m.vmu_cfg = cfg;
        m.scalar_cfg = m.vmu_scalar_cfg = in.imm;
    }

    // read side (expander.v:178-532)
    ClassRow row{SRC_NONE, A_MULVV, false, 0, 0};
    switch (f6) {
    case F6_FQMUL:
        row = f3 == 0 ? ClassRow{SRC_VV, A_MULVV, false, 4, 8}
            : f3 == 1 ? ClassRow{SRC_VS1, A_MULVS, true, 4, 8} : ClassRow{SRC_NONE, A_MULVV, false, 0, 0};
        break;
    case F6_FQADD:
        row = f3 == 0 ? ClassRow{SRC_VV, A_ADDVV, false, 4, 8}
            : f3 == 1 ? ClassRow{SRC_VS1, A_ADDVS, true, 4, 8} : ClassRow{SRC_NONE, A_ADDVV, false, 0, 0};
        break;
    case F6_FQSUB:   // .sv takes its vector operand from vs2 (expander.v:342-363)
        row = f3 == 0 ? ClassRow{SRC_VV, A_SUBVV, false, 4, 8}
            : f3 == 1 ? ClassRow{SRC_VS1, A_SUBVS, true, 4, 8}
            : f3 == 2 ? ClassRow{SRC_VS2, A_SUBSV, true, 4, 8} : ClassRow{SRC_NONE, A_SUBVV, false, 0, 0};
        break;
    case F6_FQMOD: row = ClassRow{SRC_VS1, A_MOD, false, 4, 8}; break;
    case F6_VCPY: row = ClassRow{SRC_VS1, A_ADDVS, false, 4, 8}; break;     // addmod(x, 0)
    case F6_NTT: row = ClassRow{SRC_VS1, A_CT, false, 0, 2}; break;
    case F6_INTT: row = ClassRow{SRC_VS1, A_GS, false, 0, 8}; break;
    case F6_VAUT:
    case F6_VROLI: row = ClassRow{SRC_VS1, A_MULVV, false, 0, 2}; break;
    case F6_VSE: row = ClassRow{SRC_VS1, A_MULVV, false, 0, 1}; break;
    default: break;
    }
    m.alu = row.alu;
    if (row.scalar_is_imm) m.scalar_alu = in.imm;
    auto port = [](uint8_t reg) { return (uint64_t(reg) << 1) | 1; };   // {reg, enable}
    if (row.src == SRC_VV) {
        const bool odd = in.vs1 & 1;
        m.b0r = port(odd ? in.vs2 : in.vs1);
        m.b1r = port(odd ? in.vs1 : in.vs2);
        m.muxo = odd ? row.muxo_odd : row.muxo_even;
    } else if (row.src != SRC_NONE) {
        const uint8_t reg = row.src == SRC_VS1 ? in.vs1 : in.vs2;
        (reg & 1 ? m.b1r : m.b0r) = port(reg);
        m.muxo = (reg & 1) ? row.muxo_odd : row.muxo_even;
    }

    // interconnect + NTT engine (expander.v:533-578)
    switch (f6) {
    case F6_NTT: m.iconn = 4; m.ntt = 2; break;
    case F6_INTT: m.iconn = 5; m.ntt = 3; break;
    case F6_VAUT: m.iconn = 1; m.scalar_iconn = csr_step + in.imm; break;
    case F6_VROLI: m.iconn = 2; m.scalar_iconn = in.imm; break;
    default: break;
    }

    // write side, keyed on funct6[1:0] (expander.v:579-668)
    int muxi_even = -1, muxi_odd = -1;
    switch (f6 & 3) {
    case 1: {
        const bool perm = f6 == F6_VAUT || f6 == F6_VROLI;
        muxi_even = perm ? 4 : 0;
        muxi_odd = perm ? 1 : 0;
        break;
    }
    case 2:
        if (f6 == F6_NTT) muxi_even = muxi_odd = 0;
        else if (f6 == F6_INTT) { muxi_even = 4; muxi_odd = 1; }
        break;
    case 3:
        if (f6 == F6_VLE) { muxi_even = 0xc; muxi_odd = 0x3; }
        break;
    default: break;
    }
    if (muxi_even >= 0) {
        (in.vd & 1 ? m.b1w : m.b0w) = port(in.vd);
        m.muxi = uint64_t((in.vd & 1) ? muxi_odd : muxi_even);
    }

    // load/store unit (expander.v:669-698)
    if (f6 == F6_VLE) { m.ls = 1; m.scalar_ls = in.imm; }
    else if (f6 == F6_VSE) { m.ls = 2; m.scalar_ls = in.imm; }
    return m;
}

// Physical register behind a bank port: bank = reg & 1, index = reg >> 1 (vxu_lane.sv:321-328), so
// a port field naming register r on bank b reads/writes register (r & ~1) | b.  -1 = port disabled.
inline int port_reg(uint64_t field, int bank) {
    return (field & 1) ? int((((field >> 1) & 31) & ~1u) | unsigned(bank)) : -1;
}

}  // namespace alb
