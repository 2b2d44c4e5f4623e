/*
 * oracle/golden_model.h -- C API of the CPU golden model of ALOHA's vector-processor ISA.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under aloha_b200/ links, loads or calls this; it is
 * imported solely by tests/, __graft_entry__.smoke() (as the checker) and bench.py's
 * cpu_baseline / --impl reference legs (as the timed CPU arm).
 *
 * What it restates (paths relative to the reference checkout, read-only, not shipped):
 *   src/vp/sequncer/expander.v:65-107,123-130,154-700   96-bit word -> 17 micro-op fields
 *   src/vp/sequncer/seq_top.v:417-429                   VL / Q / IQ config registers
 *   src/vp/top/vp_top_full.sv:105-117, vmu/addr_gen.v:44  VLE/VSE row addressing
 *   src/vp/vxu/vxu_lane.sv:564-603                      operand muxes, VAUT / VROLI address+sign
 *   src/vp/vxu/modalu.sv:22-46,152-249,351-379          14 ALU opcodes, single pre-reduce
 *   src/vp/vxu/modmul.sv:150-252                        Barrett with the 58 / 63 / 61-bit shifts
 *   src/vp/vxu/halfred.sv:23-26                         x/2 mod q
 *   src/vp/ntt/ntt_fsm.sv:49-81                         constant-geometry stage schedule, ping-pong
 *   sim/vp/tf_rom_generator/tf_rom_generator.sv:28-63,75-148  twiddle order psi^bitrev(j)
 *   sim/top/top_noaxilite_tb.sv:396-532                 run_vp / DMA host contract
 *
 * Parity status: PINNED.  tests/test_oracle_tv.py checks it bit-exactly against every golden
 * vector the reference ships for this path (66 per-op RTL dumps of the three tv/ cases, 44
 * kernel-level software-model vectors, 77/79 rows of the sequencer decode goldens -- the two
 * remaining rows are the reference-internal VFQSUB.sv conflict).
 */
#ifndef ALOHA_ORACLE_GOLDEN_MODEL_H
#define ALOHA_ORACLE_GOLDEN_MODEL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gm gm_t;

enum {
    GM_OK = 0,
    GM_E_ARG = -1,        /* bad argument / null */
    GM_E_RANGE = -2,      /* SPM / KSK / ISRAM row out of range */
    GM_E_OPCODE = -3,     /* funct6 the expander does not know */
    GM_E_STATE = -4,      /* vl / q not configured, N unsupported, no twiddles provisioned */
    GM_E_ILLEGAL = -5,    /* stream the RTL has no defined behaviour for (vd==vs1 on NTT/VAUT, even k) */
    GM_E_NOBREAK = -6     /* ran off the instruction ROM without BREAK */
};

/* vlmax_bits: SYS_VLMAX (vp_defines.vh:24; 524288 => N<=8192, VAUT k truncated to 13 bits). */
gm_t *gm_create(uint64_t vlmax_bits, uint32_t spm_rows, uint32_t ksk_rows);
void gm_destroy(gm_t *);

/* Twiddle ROM provisioning: table i holds psi_i (primitive 2*Nmax-th root, Nmax =
 * vlmax_bits/64) for modulus q_i.  VSETQ picks the table whose q matches, else the LAST one
 * (vxu_top.sv:112-118: q0 -> 0, q1 -> 1, anything else -> 2). */
int gm_set_moduli(gm_t *, const uint64_t *q, const uint64_t *psi, uint32_t n);

/* words: n x 12 bytes, byte 0 = most significant (as the 24-hex-digit $readmemh text reads). */
int gm_load_isram(gm_t *, const uint8_t *words, uint32_t n, uint32_t at_pc);

int gm_dma_mem_h2d(gm_t *, uint32_t spm_row, const uint64_t *src, uint64_t bytes);
int gm_dma_mem_d2h(gm_t *, uint64_t *dst, uint32_t spm_row, uint64_t bytes);
int gm_dma_ksk_h2d(gm_t *, uint32_t ksk_row, const uint64_t *src, uint64_t bytes);
/* out[i] = 1 if SPM word (spm_row*128 + i) was ever written (the 'x' lines of the RTL dumps). */
int gm_spm_written(gm_t *, uint32_t spm_row, uint64_t nwords, uint8_t *out);

int gm_run_vp(gm_t *, uint32_t pc, uint32_t src0, uint32_t src1, uint32_t rslt, uint32_t ksk_ptr,
              uint32_t step);
/* number of instructions retired by the last gm_run_vp, incl. BREAK */
uint32_t gm_last_inst_count(const gm_t *);

int gm_vreg_read(gm_t *, uint32_t reg, uint64_t *dst, uint64_t nwords);
int gm_vreg_write(gm_t *, uint32_t reg, const uint64_t *src, uint64_t nwords);
int gm_get_csr(const gm_t *, uint64_t *vl, uint64_t *q, uint64_t *iq);

/* ---- stateless pieces, exposed for unit tests and for the CPU baseline ---- */

/* 17 decoded fields in the order of sim/vp/sequncer/seq_top_tb.sv:138-160. */
int gm_decode(const uint8_t word[12], uint64_t csr_step, uint64_t out[17]);

uint64_t gm_barrett(uint64_t a, uint64_t b, uint64_t q, uint64_t iq);
uint64_t gm_half(uint64_t x, uint64_t q);
/* one modalu evaluation; returns res0, *res1 gets the second output (CT / GS), else 0. */
uint64_t gm_alu(uint32_t opcode, uint64_t a, uint64_t b, uint64_t s, uint64_t q, uint64_t iq,
                uint64_t *res1);

/* floor(2^121 / q) -- the VSETIQ immediate (modmul_tb.sv:30-36) */
uint64_t gm_barrett_iq(uint64_t q);
uint64_t gm_powmod(uint64_t a, uint64_t e, uint64_t q);
/* minimal primitive 2n-th root of unity mod q (0 if 2n does not divide q-1) */
uint64_t gm_min_primitive_root(uint64_t q, uint64_t two_n);

/* stand-alone transforms on a[0..n): constant-geometry schedule + Barrett, exactly as run_vp does.
 * inverse=0: out[k] = a(psi^(2*bitrev(k)+1)); inverse=1: exact inverse incl. 1/N by per-stage halving.
 * scratch: n words.  Result in a. */
int gm_ntt(uint64_t *a, uint64_t *scratch, uint64_t n, uint64_t q, uint64_t iq, uint64_t psi,
           int inverse);
/* batch of `count` independent limb-polys: poly j at a + j*n uses modulus index mod_idx[j].
 * nthreads host threads (>=1).  Twiddles are built once per modulus outside the timed part if
 * `tables` != NULL from gm_ntt_tables_create. */
typedef struct gm_ntt_tables gm_ntt_tables_t;
gm_ntt_tables_t *gm_ntt_tables_create(uint64_t n, const uint64_t *q, const uint64_t *psi,
                                      uint32_t n_moduli);
void gm_ntt_tables_destroy(gm_ntt_tables_t *);
int gm_ntt_batch(const gm_ntt_tables_t *, uint64_t *a, const uint32_t *mod_idx, uint64_t count,
                 int inverse, uint32_t nthreads);

/* vd[(i*k) mod n] = ((i*k) mod 2n >= n) ? q - x[i] : x[i]   (raw 64-bit, vxu_lane.sv:594-599) */
int gm_automorph(uint64_t *dst, const uint64_t *src, uint64_t n, uint64_t k, uint64_t q);
/* acc[i] = addmod(acc[i], barrett(aut_k(x)[i], p[i]))  over `count` limb-polys, nthreads threads */
int gm_aut_mac_batch(uint64_t *acc, const uint64_t *x, const uint64_t *p, uint64_t n, uint64_t k,
                     const uint64_t *q, const uint32_t *mod_idx, uint64_t count,
                     uint32_t nthreads);

#ifdef __cplusplus
}
#endif
#endif
