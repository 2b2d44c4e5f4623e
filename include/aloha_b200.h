/*
 * aloha_b200.h -- C-ABI of the B200-native execution engine for ALOHA's leveled-FHE vector datapath.
 *
 * The reference (AntChainOpenLabs/ALOHA) is RTL: its only software-visible boundary is the
 * CSR / start / done handshake and the AXI DMA of the SoC top.  Each entry point below replaces one
 * of those hardware interfaces, or the testbench task that drives it (paths relative to the
 * reference checkout):
 *
 *   aloha_create / aloha_destroy   reset + elaboration parameters   src/vp/include/vp_defines.vh:24-31
 *   aloha_load_isram               $readmemh of the instruction ROM  src/vp/sequncer/inst_rom.v;
 *                                                                     sim/vp/isram_file_generator/isram_file_generator.sv:27-31
 *   aloha_load_tf_rom              per-lane twiddle ROM images        sim/vp/tf_rom_generator/tf_rom_generator.sv:75-148;
 *                                                                     src/vp/vxu/vxu_lane.sv:210-218
 *   aloha_dma_mem_h2d              DMA_CMD_MEM read channel           sim/top/top_noaxilite_tb.sv:450-472; src/mem_buf/axi_data_rd_top.sv:58-99
 *   aloha_dma_mem_d2h              DMA write channel                  sim/top/top_noaxilite_tb.sv:474-520
 *   aloha_dma_mem_*_async          the same two channels, free-running beside the VP as in the RTL
 *   aloha_dma_ksk_h2d              DMA_CMD_KSK                        sim/top/top_noaxilite_tb.sv:372-394
 *   aloha_run_vp                   CSR writes + vp_start + poll done  sim/top/top_noaxilite_tb.sv:396-417;
 *                                                                     src/top/h2_top_no_axilite.sv:145-182; src/mem_buf/axil_parse.sv:50-72
 *   aloha_run_vp_batch             the same handshake issued back-to-back for `count` CSR sets;
 *                                  architecturally identical to a loop over aloha_run_vp, handed to
 *                                  the batcher in one piece so independent limbs share launches
 *   aloha_run_vp_multi             ditto, one start pc per call
 *   aloha_flush                    (ALOHA_F_DEFER) launch what has been queued, without waiting
 *   aloha_group_*                  no counterpart on the single-chip reference: the DMA block's shape
 *                                  (SPM rows, asynchronous, start/done ordering) extended to scratchpad-to-
 *                                  scratchpad transfers between the engines of a limb-sharded machine; NCCL
 *                                  over NVLink underneath (SURVEY 8(b) `n_gpus`, 8(e))
 *   aloha_host_*                   the testbench's host driver        sim/top/top_noaxilite_tb.sv:249-298 (parse_op),
 *                                                                     :419-532 (run_* tasks), :536-565 (dump_poly), :596-638 (run)
 *
 * Contract: one in-order command stream per aloha_t; calls on one handle are not re-entrant (the
 * hardware has a single start/done handshake).  Host buffers belong to the caller; SPM, KSK memory,
 * the 32 vector registers and {vl, q, iq} live on the device and persist across calls, as in the
 * RTL.  All functions return 0 or a negative ALOHA_E_* code; the reference has no error reporting
 * at all (its testbench says "TODO: error handle"), so every code below is new.
 *
 * There is no CPU fallback: aloha_create fails with ALOHA_E_CUDA when no sm_100 device is usable.
 */
#ifndef ALOHA_B200_H
#define ALOHA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aloha aloha_t;
typedef struct aloha_host aloha_host_t;
typedef struct aloha_group aloha_group_t;

enum {
    ALOHA_OK = 0,
    ALOHA_E_ARG = -1,         /* null / misaligned / malformed argument */
    ALOHA_E_RANGE = -2,       /* SPM, KSK or instruction-ROM address out of range */
    ALOHA_E_OPCODE = -3,      /* funct6 / funct3 the decoder does not know */
    ALOHA_E_STATE = -4,       /* vl or q not configured, N unsupported, modulus without twiddles */
    ALOHA_E_ILLEGAL = -5,     /* stream with no defined RTL behaviour (vd == vs1 on VNTT/VINTT/VAUT/VROLI, even k) */
    ALOHA_E_NOBREAK = -6,     /* ran off the instruction ROM without BREAK */
    ALOHA_E_UNDEFINED = -7,   /* read of a vector register whose content this engine does not define
                                 (never written, or clobbered as the source of VNTT/VINTT) */
    ALOHA_E_CUDA = -8,        /* CUDA runtime error; see aloha_last_error */
    ALOHA_E_NOMEM = -9
};

enum {
    ALOHA_F_NO_BATCH = 1u << 0,  /* one launch per instruction, in program order (debug / parity bisection) */
    ALOHA_F_NO_ALIAS = 1u << 1,  /* materialise every VLE / VSE as a copy (debug) */
    ALOHA_F_GRAPHS = 1u << 2,    /* replay cached plans as CUDA graphs */
    ALOHA_F_NO_FUSE = 1u << 3,   /* keep VAUT / VFQMUL / VFQADD chains as separate kernels */
    ALOHA_F_DEFER = 1u << 5,     /* PROGRAM-level scheduling: aloha_run_vp* only queue the call; the queue is planned and
                                    launched as ONE batch at the next DMA / aloha_sync / state query, so consecutive host
                                    ops (mul_plain, hom_add, rotate ... of one program) share launches wherever their
                                    data dependencies allow.  Results are identical; an error in a queued call is
                                    reported by the call that flushes it. */
    ALOHA_F_STRICT = 1u << 4,    /* VNTT / VINTT run the RTL's constant-geometry schedule stage by stage with the
                                    RTL ALU: word-exact for ANY input (also >= 2q) and the source register keeps the
                                    ping-pong intermediate the RTL leaves there.  ~10x slower transforms. */
    ALOHA_F_AUT_GATHER = 1u << 7,  /* Automorphisms have two kernels: a shared-memory tile permutation that moves both sides
                                    in whole lines (aut_plan.hpp) and a destination-ordered 8-byte gather from L2.  By
                                    default a standalone VAUT takes the tiles (the gather is bound by L2 transactions, one
                                    per element) and the fused rotate-MAC takes the gather (one stream in four: the
                                    transactions hide behind the other three).  This flag forces the gather for both, */
    ALOHA_F_AUT_TILED = 1u << 8,   /* this one the tiles for both.  Same results; for A/B measurements and tests. */
    ALOHA_F_GENERIC_MODMUL = 1u << 6  /* transforms use the any-prime (Shoup) arithmetic even for moduli of the form
                                    2^60 - d, d <= 2^27, which otherwise take the cheaper pseudo-Mersenne product.
                                    Same results; for A/B measurements and tests. */
};

typedef struct aloha_cfg {
    uint64_t vlmax_bits;      /* SYS_VLMAX; 524288 on the reference (N <= 8192, VAUT k kept to 13 bits) */
    uint32_t spm_rows;        /* scratchpad rows of 128 x u64 (1 KiB); reference: 16384 */
    uint32_t ksk_rows;        /* key-switch-key memory rows; reference: 9216 */
    int32_t device;           /* CUDA device ordinal */
    uint32_t flags;           /* ALOHA_F_* */
    uint32_t pool_buffers;    /* renaming buffers of vlmax_bits/64 words (0 = 256) */
    uint64_t l2_chunk_bytes;  /* split transform launches so one chunk's footprint stays below this (0 = never split) */
    uint32_t isram_depth;     /* instruction ROM entries, IRAM_DEPTH (vp_defines.vh:31); 0 = 4096 as on the reference */
    uint32_t reserved;
} aloha_cfg;

typedef struct aloha_vp_args {
    uint32_t src0, src1, rslt, ksk_ptr, step;   /* the five CSRs of one run_vp */
} aloha_vp_args;

typedef struct aloha_stats {
    uint64_t kernel_launches;     /* CUDA kernels launched by this handle */
    uint64_t instructions;        /* instructions retired (incl. config and BREAK) */
    uint64_t plans_built, plans_reused;
    uint64_t copies_elided;       /* VLE / VSE turned into aliases or forwarded stores */
    uint64_t copies_emitted;      /* VLE / VSE / copy-on-write that needed a copy kernel */
    uint64_t limb_ntts;           /* VNTT + VINTT executed */
    uint64_t ops_fused;           /* vector ops folded into a fused kernel by the batcher */
} aloha_stats;

int aloha_create(const aloha_cfg *cfg, aloha_t **out);
void aloha_destroy(aloha_t *);
const char *aloha_strerror(int code);
const char *aloha_last_error(const aloha_t *);   /* detail text of the last failing call */

/* words: n x 12 bytes, byte 0 most significant, i.e. the 24 hex digits of one $readmemh line */
int aloha_load_isram(aloha_t *, const uint8_t *words, uint32_t n, uint32_t at_pc);
/* Twiddle provisioning (the ISA cannot set psi; the RTL bakes it into ROMs selected by q):
 * psi[i] = primitive 2*Nmax-th root of unity mod q[i], Nmax = vlmax_bits / 64. */
int aloha_load_tf_rom(aloha_t *, const uint64_t *q, const uint64_t *psi, uint32_t n_moduli);

/* bytes must be a multiple of 64 (one 512-bit AXI beat) */
int aloha_dma_mem_h2d(aloha_t *, uint32_t spm_row, const uint64_t *src, uint64_t bytes);
int aloha_dma_mem_d2h(aloha_t *, uint64_t *dst, uint32_t spm_row, uint64_t bytes);
int aloha_dma_ksk_h2d(aloha_t *, uint32_t ksk_row, const uint64_t *src, uint64_t bytes);
/* Asynchronous DMA: the reference's DMA engine is a block of its own next to the VP
 * (src/mem_buf/axi_data_rd_top.sv, axi_data_wr_top.sv); these run the copy on dedicated upload /
 * download streams so transfers in both directions overlap each other and the kernels.  Ordering is
 * the hardware's: an upload is visible to every run_vp issued after the call; a download sees every
 * run_vp issued before it; an upload never overtakes earlier work that touches the same rows.
 * Host buffers must be page-locked for the copies to be truly asynchronous and must stay valid until
 * aloha_sync (which also waits for both DMA streams). */
int aloha_dma_mem_h2d_async(aloha_t *, uint32_t spm_row, const uint64_t *src, uint64_t bytes);
int aloha_dma_mem_d2h_async(aloha_t *, uint64_t *dst, uint32_t spm_row, uint64_t bytes);
/* out[i] = 1 iff SPM word spm_row*128 + i has ever been written (the 'x' lines of the RTL dumps) */
int aloha_spm_written(aloha_t *, uint32_t spm_row, uint64_t nwords, uint8_t *out);

/* Executes from pc until BREAK.  Returns once the work is queued on the handle's stream ("done" is
 * implied by stream order: any later DMA or aloha_sync observes the results). */
int aloha_run_vp(aloha_t *, uint32_t pc, uint32_t src0, uint32_t src1, uint32_t rslt,
                 uint32_t ksk_ptr, uint32_t step);
int aloha_run_vp_batch(aloha_t *, uint32_t pc, uint32_t count, const aloha_vp_args *args);
/* As aloha_run_vp_batch, with a start pc per call (different microcode kernels in one batch, e.g. the
 * per-modulus sections of a key-switch stream whose VSETQ immediates differ). */
int aloha_run_vp_multi(aloha_t *, uint32_t count, const uint32_t *pcs, const aloha_vp_args *args);
int aloha_sync(aloha_t *);
/* Page-locked host memory for the asynchronous DMA channels (what a driver's DMA-buffer allocator hands out) */
int aloha_pinned_alloc(uint64_t bytes, void **out);
void aloha_pinned_free(void *);
/* ALOHA_F_DEFER: plan and launch everything queued so far; does not wait for the device. */
int aloha_flush(aloha_t *);

/* Zero-copy access for callers that already live on the device (and for the multi-GPU host layer,
 * which hands these to NCCL): device address of an SPM / KSK row.  Writes through these pointers
 * bypass the 'x' tracking and the register-alias bookkeeping -- call aloha_spm_mark_written. */
int aloha_spm_device_ptr(aloha_t *, uint32_t spm_row, void **dev_ptr);
int aloha_ksk_device_ptr(aloha_t *, uint32_t ksk_row, void **dev_ptr);
int aloha_spm_mark_written(aloha_t *, uint32_t spm_row, uint32_t nrows);
int aloha_set_stream(aloha_t *, void *cuda_stream);   /* cudaStream_t; NULL = the handle's own */

int aloha_get_stats(const aloha_t *, aloha_stats *out);
int aloha_get_csr(const aloha_t *, uint64_t *vl, uint64_t *q, uint64_t *iq);

/* Stateless decoder: the 17 micro-op fields in the order of sim/vp/sequncer/seq_top_tb.sv:138-160
 * (vxu cfg, scalar cfg, b0r, b0w, b1r, b1w, alu, scalar alu, iconn, scalar iconn, ntt, muxo, muxi,
 * vmu cfg, vmu scalar cfg, ls op, ls scalar). */
int aloha_decode(const uint8_t word[12], uint64_t csr_step, uint64_t out[17]);

/* ---- limb-sharded machine: several engines (one per GPU) exchanging SPM rows over NVLink -----------
 * RNS limbs are independent units: every engine holds its limbs' slice of each polynomial, its twiddles and
 * its KSK slices, and runs the same instruction streams on them.  The only cross-limb step of the path is
 * the base extension of the key-switch stream (keyswitch.mem insts 8, 12, 28, 32 generalised): before it the
 * coefficient-form digits are all-gathered, and before the mod-down the special-prime accumulators are
 * broadcast.  Two ways to form a group:
 *   one process per GPU   rank 0 calls aloha_group_unique_id and ships the 128 bytes to the other ranks by
 *                         whatever channel the host program has; every rank calls aloha_group_create;
 *   one process, n GPUs   aloha_group_create_local over engines created on n different devices (the
 *                         testbench-shaped caller of INTEGRATION.md); member i has rank i and every group call
 *                         acts on all members at once.
 * Transfers run on a communication stream per device, ordered after all engine work issued before the call
 * (like the DMA block next to the VP); aloha_group_wait orders later engine work after a transfer.  Rows of a
 * transfer in flight must not be written by engine work issued before the matching wait. */
enum { ALOHA_GROUP_ID_BYTES = 128 };
enum { ALOHA_GROUP_CHUNKED = 1u };                     /* all-gather as one transfer per source rank */
enum { ALOHA_GROUP_ALL = -1, ALOHA_GROUP_BCAST = -2 };  /* aloha_group_wait sources */
int aloha_group_unique_id(uint8_t id[ALOHA_GROUP_ID_BYTES]);
int aloha_group_create(aloha_t *engine, const uint8_t id[ALOHA_GROUP_ID_BYTES], int rank, int nranks,
                       aloha_group_t **out);
int aloha_group_create_local(aloha_t *const *engines, int n, aloha_group_t **out);
void aloha_group_destroy(aloha_group_t *);
int aloha_group_size(const aloha_group_t *);
int aloha_group_rank(const aloha_group_t *);
const char *aloha_group_last_error(const aloha_group_t *);
/* `count` all-gathers issued as one transfer, the c-th over the rows starting at spm_row + c*stride_rows:
 * rank r contributes rows [start + r*rows_per_rank, +rows_per_rank) of its SPM; all members end up with all
 * blocks.  (count > 1: the digit regions of a batch of key-switches.) */
int aloha_group_all_gather_rows(aloha_group_t *, uint32_t spm_row, uint32_t rows_per_rank, uint32_t count,
                                uint32_t stride_rows, uint32_t flags);
int aloha_group_broadcast_rows(aloha_group_t *, uint32_t spm_row, uint32_t nrows, int root);
/* source >= 0: that rank's block of the latest all-gather; ALOHA_GROUP_ALL; ALOHA_GROUP_BCAST */
int aloha_group_wait(aloha_group_t *, int source);

/* ---- host driver: the testbench's op-list replay ------------------------------------------- */
/* PROGRAM text: one op per line, "a0,a1,a2" hex u32 (top_noaxilite_tb.sv:249-298).  dram_bytes is
 * the size of the modelled DDR (the TB uses 64 MiB). */
int aloha_host_create(aloha_t *, const char *program_text, uint64_t dram_bytes, uint32_t n,
                      aloha_host_t **out);
void aloha_host_destroy(aloha_host_t *);
int aloha_host_num_ops(const aloha_host_t *);
int aloha_host_dram_write(aloha_host_t *, uint64_t byte_addr, const void *src, uint64_t bytes);
int aloha_host_dram_read(aloha_host_t *, uint64_t byte_addr, void *dst, uint64_t bytes);
/* The floating-point encoder (src/encoder, Xilinx IP) is out of scope; its SPM output for encode op
 * `op_index` (2n words) is injected, e.g. from rtl_result/inst_<i>_0_out.txt. */
int aloha_host_set_encoder_output(aloha_host_t *, uint32_t op_index, const uint64_t *data,
                                  uint64_t nwords);
/* Runs op `op_index` and produces its dump(s): dump / written get 4n entries (inst_<i>_out.txt);
 * for encode ops sub_dump / sub_written get the 4n entries of inst_<i>_0_out.txt and *has_sub = 1.
 * dump == NULL runs the op without the read-back (nothing is copied to the host, nothing blocks). */
int aloha_host_run_op(aloha_host_t *, uint32_t op_index, uint64_t *dump, uint8_t *written,
                      uint64_t *sub_dump, uint8_t *sub_written, int *has_sub);
/* PROGRAM-level scheduling: the same op without the blocking read-back the testbench does
 * after every op (:600-632).  Dumps and store_cipher data are copied out by the download channel into a
 * page-locked ring while later ops run; dump / sub_dump (and the modelled DDR) are valid after
 * aloha_host_sync.  The written masks are filled in immediately. */
int aloha_host_run_op_async(aloha_host_t *, uint32_t op_index, uint64_t *dump, uint8_t *written,
                            uint64_t *sub_dump, uint8_t *sub_written, int *has_sub);
/* ops [first, first+count) in one call; op first+j uses dumps + j*4n, written + j*4n, sub_dumps + j*4n,
 * sub_written + j*4n and has_sub[j] */
int aloha_host_run_range_async(aloha_host_t *, uint32_t first, uint32_t count, uint64_t *dumps, uint8_t *written,
                               uint64_t *sub_dumps, uint8_t *sub_written, int *has_sub);
int aloha_host_sync(aloha_host_t *);
/* "%0d" per line, 'x' for never-written words (dump_poly, top_noaxilite_tb.sv:536-565) */
int aloha_write_dump_text(const char *path, const uint64_t *data, const uint8_t *written,
                          uint64_t nwords);

#ifdef __cplusplus
}
#endif
#endif
