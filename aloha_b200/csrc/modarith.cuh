// modarith.cuh -- device-side RNS modular arithmetic for sm_100a (integer IMAD / IADD3 pipes).
//
// Two families:
//  * rtl_*   : word-for-word the reference ALU's arithmetic (one conditional subtract on every
//              input, the 58/63/61-bit Barrett of src/vp/vxu/modmul.sv:150-252, the 65-bit add of
//              modalu.sv:228, halfred.sv:23-26).  Used by the element-wise kernels so that ANY
//              64-bit input word -- canonical or not -- produces the word the RTL would store.
//  * lazy Shoup/Harvey arithmetic for the NTT kernels.  The reference's NTT stores canonical
//              residues (< q) for every input below 2q, so any exact algorithm is bit-identical;
//              these keep values in [0, 16q) (q < 2^60) and reduce once at the end.
#pragma once
#include <cstdint>

typedef unsigned long long u64;
typedef unsigned int u32;

namespace alb {

__device__ __forceinline__ u64 prered(u64 x, u64 q) { return x >= q ? x - q : x; }

// x in [0, 2c) -> [0, c)
__device__ __forceinline__ u64 csub(u64 x, u64 c) { return x >= c ? x - c : x; }

// modmul.sv:150 (prod >> 58), :172 (mid >> 63), :195 (mask 2^61), :216-252.
__device__ __forceinline__ u64 rtl_barrett(u64 a, u64 b, u64 q, u64 iq) {
    const u64 plo = a * b, phi = __umul64hi(a, b);
    const u64 prod_shift = (plo >> 58) | (phi << 6);
    const u64 mlo = prod_shift * iq, mhi = __umul64hi(prod_shift, iq);
    const u64 mid_shift = (mlo >> 63) | (mhi << 1);
    const u64 elo = mid_shift * q;  // only bits [60:0] of estim are used
    const u64 mask = 1ull << 61;
    const u64 diff = (((plo & (mask - 1)) | mask) - (elo & (mask - 1))) & (mask - 1);
    return diff < q ? diff : diff - q;
}

// modalu.sv:228-229: 65-bit sum compared against q, result truncated to 64 bits.
__device__ __forceinline__ u64 rtl_add(u64 a, u64 b, u64 q) {
    const u64 s = a + b;
    const bool carry = s < a;
    return (carry || s >= q) ? s - q : s;
}
// modalu.sv:249
__device__ __forceinline__ u64 rtl_sub(u64 a, u64 b, u64 q) { return a >= b ? a - b : q + a - b; }
// halfred.sv:23-26
__device__ __forceinline__ u64 rtl_half(u64 x, u64 q) {
    return (x >> 1) + ((x & 1) ? ((q + 1) >> 1) : 0ull);
}

// ALU opcodes as the decoder emits them (modalu.sv:22-37)
enum AluOp : u32 {
    ALU_MUL_VV = 0x00, ALU_MUL_VS = 0x04, ALU_ADD_VV = 0x01, ALU_ADD_VS = 0x05, ALU_SUB_VV = 0x02,
    ALU_SUB_VS = 0x06, ALU_SUB_SV = 0x0a, ALU_MOD = 0x03
};

// One element-wise modalu evaluation (res0 only; CT/GS never reach the element-wise path).
template <u32 OP>
__device__ __forceinline__ u64 rtl_alu(u64 a_raw, u64 b_raw, u64 s_red, u64 q, u64 iq) {
    const u64 a = prered(a_raw, q);
    if (OP == ALU_MUL_VV) return rtl_barrett(a, prered(b_raw, q), q, iq);
    if (OP == ALU_MUL_VS) return rtl_barrett(a, s_red, q, iq);
    if (OP == ALU_MOD) return rtl_barrett(a, 1, q, iq);
    if (OP == ALU_ADD_VV) return rtl_add(a, prered(b_raw, q), q);
    if (OP == ALU_ADD_VS) return rtl_add(a, s_red, q);
    if (OP == ALU_SUB_VV) return rtl_sub(a, prered(b_raw, q), q);
    if (OP == ALU_SUB_VS) return rtl_sub(a, s_red, q);
    if (OP == ALU_SUB_SV) return rtl_sub(s_red, a, q);
    return 0;
}

// ---- lazy arithmetic for the transforms ---------------------------------------------------
// The IMAD pipe is the scarce resource on sm_100 (64 thread-ops/clk/SM, half the ALU pipe's rate;
// tools/intpipe_bench.cu), so every 64-bit accumulation is folded into a multiply-add chain and the
// conditional subtracts use the cheapest ALU-only form.

// x in [0, 2c), c <= 2^63  ->  [0, c).  x - c wraps negative exactly when x < c: 5 ALU instructions
// (IADD3, IADD3.X, ISETP on the high word, 2 SEL) instead of the 6 of a 64-bit compare.
__device__ __forceinline__ u64 csub_s(u64 x, u64 c) {
    const u64 t = x - c;
    return ((long long)t < 0) ? x : t;
}

// acc + y*w - floor(y*wp / 2^64) * q   (mod 2^64), with nq = 2^64 - q.
// Shoup: wp = floor(w * 2^64 / q); for ANY y < 2^64 the product part is y*w mod q in [0, 2q).
// 4 IMAD.WIDE (exact mulhi) + 2 IMAD.WIDE + 4 IMAD; the two 64-bit accumulations ride the wides.
__device__ __forceinline__ u64 shoup_mac(u64 acc, u64 y, u64 w, u64 wp, u64 nq) {
    const u64 qh = __umul64hi(y, wp);
    const u32 yl = (u32)y, yh = (u32)(y >> 32), wl = (u32)w, wh = (u32)(w >> 32);
    const u32 ql = (u32)qh, qhh = (u32)(qh >> 32), nl = (u32)nq, nh = (u32)(nq >> 32);
    u64 r = acc;
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(yl), "r"(wl));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(ql), "r"(nl));
    u32 hi = (u32)(r >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(yl), "r"(wh));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(yh), "r"(wl));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(ql), "r"(nh));
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(qhh), "r"(nl));
    return ((u64)hi << 32) | (u32)r;
}
__device__ __forceinline__ u64 mul_shoup(u64 y, u64 w, u64 wp, u64 nq) { return shoup_mac(0, y, w, wp, nq); }

// x < 2^64, q in (2^59, 2^60), mest = floor(2^91 / q) (32 bits).  floor(x/q) - 1 <= est <= floor(x/q),
// so x - est*q is in [0, 2q); one conditional subtract makes it canonical.
__device__ __forceinline__ u64 reduce_lazy(u64 x, u64 nq, u32 mest) {     // -> [0, 2q)
    const u32 est = __umulhi((u32)(x >> 32), mest) >> 27;
    u64 r = x;
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"(est), "r"((u32)nq));
    u32 hi = (u32)(r >> 32);
    asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(hi) : "r"(est), "r"((u32)(nq >> 32)));
    return ((u64)hi << 32) | (u32)r;
}
__device__ __forceinline__ u64 reduce_full(u64 x, u64 q, u64 nq, u32 mest) { return csub_s(reduce_lazy(x, nq, mest), q); }

// ---- pseudo-Mersenne moduli: q = 2^60 - d, d <= 2^27 ----------------------------------------------
// (the synthetic configs' prime rule -- scan down from 2^60 with q = 1 mod 2N -- only produces these)
// With the twiddle stored as the pair {w, w * 2^32 mod q} the product w*y splits into two 60x32-bit
// halves, T = w * y_lo + (w 2^32 mod q) * y_hi < 2^93, and T mod q needs ONE fold at 2^61 = 2d (mod q):
// 5 IMAD.WIDE per modular multiplication against Shoup's 10, for ANY 64-bit y.  Result in [0, 3q):
//   T >> 61 <= 2^32 - 2, so the fold is < 2^61 + 2^33 d - 4d <= 3 * 2^60 - 4d < 3q for d <= 2^27.
__device__ __forceinline__ u64 mul_pm(u64 y, u64 w, u64 w2, u32 d2) {
    u64 r;
    asm("{\n\t.reg .u64 X, B, U, C;\n\t.reg .u32 xl, xh, u0, u1, y0, y1, rh, m;\n\t"
        "mul.wide.u32 U, %5, %2;\n\t"          // U = w_hi * y_lo            (< 2^60)
        "mad.wide.u32 U, %6, %4, U;\n\t"       //   + w2_hi * y_hi           (< 2^61)
        "mov.b64 {u0, u1}, U;\n\t"
        "mul.wide.u32 X, %1, %2;\n\t"          // X = w_lo * y_lo
        "mul.wide.u32 B, %3, %4;\n\t"
        "add.cc.u64 X, X, B;\n\t"              //   + w2_lo * y_hi           (carry = 2^64)
        "addc.u32 u1, u1, 0;\n\t"
        "mov.b64 {xl, xh}, X;\n\t"
        "add.cc.u32 y0, u0, xh;\n\t"           // T = (y1 : y0 : xl)
        "addc.u32 y1, u1, 0;\n\t"
        "shf.r.wrap.b32 rh, y0, y1, 29;\n\t"   // T >> 61
        "and.b32 m, y0, 0x1fffffff;\n\t"
        "mov.b64 C, {xl, m};\n\t"              // T mod 2^61
        "mad.wide.u32 %0, rh, %7, C;\n\t}"
        : "=l"(r)
        : "r"((u32)w), "r"((u32)y), "r"((u32)w2), "r"((u32)(y >> 32)), "r"((u32)(w >> 32)), "r"((u32)(w2 >> 32)), "r"(d2));
    return r;
}

// The same product left in two pieces, w*y mod q = P + L with P = (T >> 61) * 2d and L = T mod 2^61 (both
// below 2^61): the forward butterfly adds them as separate operands, x' = x + P + L and
// y' = (x + 3q - P) - L, which ptxas turns into 3-input IADD3 / IADD3.X pairs -- two instructions fewer
// per butterfly than assembling the product first, and none of the carry adds lands on the IMAD pipe.
__device__ __forceinline__ void mul_pm_parts(u64 y, u64 w, u64 w2, u32 d2, u64 &P, u64 &L) {
    u32 xl, y0, y1;
    asm("{\n\t.reg .u64 X, B, U;\n\t.reg .u32 xh, u0, u1;\n\t"
        "mul.wide.u32 U, %7, %4;\n\t"
        "mad.wide.u32 U, %8, %6, U;\n\t"
        "mov.b64 {u0, u1}, U;\n\t"
        "mul.wide.u32 X, %3, %4;\n\t"
        "mul.wide.u32 B, %5, %6;\n\t"
        "add.cc.u64 X, X, B;\n\t"
        "addc.u32 u1, u1, 0;\n\t"
        "mov.b64 {%0, xh}, X;\n\t"
        "add.cc.u32 %1, u0, xh;\n\t"
        "addc.u32 %2, u1, 0;\n\t}"
        : "=r"(xl), "=r"(y0), "=r"(y1)
        : "r"((u32)w), "r"((u32)y), "r"((u32)w2), "r"((u32)(y >> 32)), "r"((u32)(w >> 32)), "r"((u32)(w2 >> 32)));
    const u32 rh = __funnelshift_r(y0, y1, 29);
    P = (u64)rh * d2;
    L = ((u64)(y0 & 0x1fffffffu) << 32) | xl;
}
// any x < 2^64  ->  x mod q in [0, 2^60 + 15 d) (below 2q): fold at 2^60 = d (mod q).
__device__ __forceinline__ u64 fold_pm(u64 x, u32 d) {
    u64 r = x & ((1ull << 60) - 1);
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(r) : "r"((u32)(x >> 60)), "r"(d));
    return r;
}
// any x < 2^64  ->  canonical
__device__ __forceinline__ u64 canon_pm(u64 x, u64 q, u32 d) { return csub_s(fold_pm(x, d), q); }

}  // namespace alb
