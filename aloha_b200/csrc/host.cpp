// host.cpp -- the host driver: PROGRAM parsing and op-list replay over a modelled DDR.
//
// Restates the reference testbench's software-visible flow (sim/top/top_noaxilite_tb.sv):
//   parse_op :249-298, run_vp :396-417, run_encode :419-448, run_load_cipher :450-472,
//   run_store_cipher :474-496, copy_spm_to_dram :498-520, run_mul_plain/hom_add/rotate :522-532,
//   dump_poly :536-565, run :596-638.
// Everything here sits ABOVE the C-ABI boundary and only calls the aloha_* entry points.
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/aloha_b200.h"

namespace {

enum OpType { OP_LOAD = 1, OP_STORE = 2, OP_ENCODE = 3, OP_ENCODE_POST = 4, OP_MUL_PLAIN = 5, OP_HOM_ADD = 6, OP_ROTATE = 7 };

struct HostOp {
    int type;
    uint32_t spm_addr, src1, src2, step;
    uint64_t dram_addr;
};

constexpr uint32_t kPcEncodePost = 0, kPcMulPlain = 64, kPcHomAdd = 160, kPcKeyswitch = 256;  // tb:63-66
constexpr uint64_t kDramVpBase = 10485760;                                                     // tb:45
constexpr uint32_t kLanes = 128;

uint32_t clog2(uint32_t x) { uint32_t l = 0; while ((1u << l) < x) ++l; return l; }

uint64_t pow3_mod(uint32_t e, uint64_t m) {
    uint64_t r = 1, b = 3 % m;
    while (e) { if (e & 1) r = r * b % m; b = b * b % m; e >>= 1; }
    return r;
}

}  // namespace

// A store to the modelled DDR whose read-back is still in flight (the download channel writes straight into
// the page-locked DDR image); its dump, if one was asked for, is copied out of the DDR by aloha_host_sync.
struct PendingStore {
    uint64_t dram_addr, bytes;
    uint64_t *dump;           // caller's buffer, or null
};

struct aloha_host {
    aloha_t *eng;
    uint32_t n;
    std::vector<HostOp> ops;
    uint8_t *dram_data = nullptr;     // the modelled DDR, page-locked so both DMA channels reach it asynchronously
    uint64_t dram_size = 0;
    std::map<uint32_t, std::vector<uint64_t>> encoder;
    std::vector<PendingStore> pending;
    bool in_flight = false;           // asynchronous read-backs issued since the last sync
    struct Dram {                     // vector-like view used by the op code below
        aloha_host *h;
        uint8_t *data() const { return h->dram_data; }
        uint64_t size() const { return h->dram_size; }
    } dram{this};
};

extern "C" {

int aloha_host_create(aloha_t *eng, const char *program, uint64_t dram_bytes, uint32_t n, aloha_host_t **out) {
    if (!eng || !program || !out || !n || (n & (n - 1))) return ALOHA_E_ARG;
    aloha_host *H = new aloha_host();
    H->eng = eng;
    H->n = n;
    std::istringstream in(program);
    std::string line;
    while (std::getline(in, line)) {
        if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
        unsigned a0, a1, a2;
        if (std::sscanf(line.c_str(), "%x,%x,%x", &a0, &a1, &a2) != 3) { delete H; return ALOHA_E_ARG; }
        HostOp op{};
        op.type = (a0 >> 28) & 0xf;
        op.spm_addr = a0 & 0x3fff;
        switch (op.type) {
        case OP_LOAD: case OP_STORE: case OP_ENCODE: op.dram_addr = ((uint64_t)a1 << 32) | a2; break;
        case OP_ENCODE_POST: case OP_MUL_PLAIN: case OP_HOM_ADD: op.src1 = a1 & 0x3fff; op.src2 = a2 & 0x3fff; break;
        case OP_ROTATE: op.step = a1 & 0x3fff; op.src1 = a2 & 0x3fff; break;
        default: delete H; return ALOHA_E_OPCODE;   // the TB has "TODO: error handle" here
        }
        H->ops.push_back(op);
    }
    if (aloha_pinned_alloc(dram_bytes, (void **)&H->dram_data) != ALOHA_OK) { delete H; return ALOHA_E_NOMEM; }
    std::memset(H->dram_data, 0, dram_bytes);
    H->dram_size = dram_bytes;
    *out = H;
    return ALOHA_OK;
}

void aloha_host_destroy(aloha_host_t *H) {
    if (!H) return;
    if (H->in_flight) aloha_sync(H->eng);
    aloha_pinned_free(H->dram_data);
    delete H;
}

int aloha_host_num_ops(const aloha_host_t *H) { return H ? (int)H->ops.size() : ALOHA_E_ARG; }

// Every read-back requested through aloha_host_run_op_async so far is in its destination.
int aloha_host_sync(aloha_host_t *H) {
    if (!H) return ALOHA_E_ARG;
    int rc = aloha_sync(H->eng);
    if (rc) return rc;
    for (const PendingStore &p : H->pending)
        if (p.dump) std::memcpy(p.dump, H->dram_data + p.dram_addr, p.bytes);
    H->pending.clear();
    H->in_flight = false;
    return ALOHA_OK;
}

int aloha_host_dram_write(aloha_host_t *H, uint64_t addr, const void *src, uint64_t bytes) {
    if (!H || !src) return ALOHA_E_ARG;
    if (addr + bytes > H->dram.size()) return ALOHA_E_RANGE;
    if (H->in_flight) { int rc = aloha_host_sync(H); if (rc) return rc; }
    std::memcpy(H->dram.data() + addr, src, bytes);
    return ALOHA_OK;
}
int aloha_host_dram_read(aloha_host_t *H, uint64_t addr, void *dst, uint64_t bytes) {
    if (!H || !dst) return ALOHA_E_ARG;
    if (addr + bytes > H->dram.size()) return ALOHA_E_RANGE;
    if (H->in_flight) { int rc = aloha_host_sync(H); if (rc) return rc; }
    std::memcpy(dst, H->dram.data() + addr, bytes);
    return ALOHA_OK;
}
int aloha_host_set_encoder_output(aloha_host_t *H, uint32_t op_index, const uint64_t *data, uint64_t nwords) {
    if (!H || !data || op_index >= H->ops.size() || nwords != 2ull * H->n) return ALOHA_E_ARG;
    H->encoder[op_index].assign(data, data + nwords);
    return ALOHA_OK;
}

namespace {

// SPM rows -> host buffer on the download channel; the buffer is valid after aloha_host_sync
int read_back_async(aloha_host *H, uint32_t row, uint64_t bytes, uint64_t *dst) {
    H->in_flight = true;
    return aloha_dma_mem_d2h_async(H->eng, dst, row, bytes);
}

}  // namespace

// aloha_host_run_op without the blocking read-backs: the op is issued, its dump(s) are copied out by the
// engine's download channel while later ops run, and the caller's buffers are valid after aloha_host_sync.
// dump / sub_dump should be page-locked (aloha_pinned_alloc) -- pageable buffers work but make the copy
// synchronous.  The `written` masks are host-side state and are filled in immediately.
int aloha_host_run_op_async(aloha_host_t *H, uint32_t i, uint64_t *dump, uint8_t *written, uint64_t *sub_dump,
                            uint8_t *sub_written, int *has_sub) {
    if (!H || i >= H->ops.size()) return ALOHA_E_ARG;
    const bool want_dump = dump && written;
    const HostOp &op = H->ops[i];
    const uint64_t words = 4ull * H->n, bytes = words * 8;
    aloha_t *E = H->eng;
    int rc = ALOHA_OK;
    if (has_sub) *has_sub = 0;
    switch (op.type) {
    case OP_LOAD: {
        const uint64_t a = kDramVpBase + op.dram_addr;
        if (a + bytes > H->dram.size()) return ALOHA_E_RANGE;
        for (const PendingStore &p : H->pending)         // a store to these DDR bytes may still be in flight
            if (p.dram_addr < a + bytes && a < p.dram_addr + p.bytes) {
                rc = aloha_host_sync(H);
                if (rc) return rc;
                break;
            }
        H->in_flight = true;                             // the upload channel reads the DDR image asynchronously
        rc = aloha_dma_mem_h2d_async(E, op.spm_addr, (const uint64_t *)(H->dram.data() + a), bytes);
        break;
    }
    case OP_STORE: {
        const uint64_t a = kDramVpBase + op.dram_addr;
        if (a + bytes > H->dram.size()) return ALOHA_E_RANGE;
        for (const PendingStore &p : H->pending)         // an earlier store to these DDR bytes still owes its dump,
            if (p.dram_addr < a + bytes && a < p.dram_addr + p.bytes) {   // which is copied out of the DDR at the sync
                rc = aloha_host_sync(H);
                if (rc) return rc;
                break;
            }
        rc = read_back_async(H, op.spm_addr, bytes, (uint64_t *)(H->dram.data() + a));
        if (rc) return rc;
        H->pending.push_back(PendingStore{a, bytes, want_dump ? dump : nullptr});
        return want_dump ? aloha_spm_written(E, op.spm_addr, words, written) : ALOHA_OK;
    }
    case OP_ENCODE: {
        auto it = H->encoder.find(i);
        if (it == H->encoder.end()) return ALOHA_E_STATE;
        rc = aloha_dma_mem_h2d(E, op.spm_addr, it->second.data(), it->second.size() * 8);
        if (rc) return rc;
        if (sub_dump && sub_written) {
            rc = read_back_async(H, op.spm_addr, bytes, sub_dump);
            if (!rc) rc = aloha_spm_written(E, op.spm_addr, words, sub_written);
            if (rc) return rc;
            if (has_sub) *has_sub = 1;
        }
        rc = aloha_run_vp(E, kPcEncodePost, op.spm_addr, 0, op.spm_addr, 0, 0);
        break;
    }
    case OP_MUL_PLAIN: rc = aloha_run_vp(E, kPcMulPlain, op.src1, op.src2, op.spm_addr, 0, 0); break;
    case OP_HOM_ADD: rc = aloha_run_vp(E, kPcHomAdd, op.src1, op.src2, op.spm_addr, 0, 0); break;
    case OP_ROTATE: {
        const uint32_t k = (uint32_t)pow3_mod(op.step, 2ull * H->n);
        const uint32_t ksk_ptr = (clog2(op.step) - 1) * H->n * 12 / kLanes;
        rc = aloha_run_vp(E, kPcKeyswitch, op.src1, 0, op.spm_addr, ksk_ptr, k);
        break;
    }
    default: return ALOHA_E_OPCODE;
    }
    if (rc || !want_dump) return rc;
    rc = read_back_async(H, op.spm_addr, bytes, dump);
    if (rc) return rc;
    // the mask describes SPM after this op: the run_vp above has been planned by now (aloha_spm_written
    // flushes a deferred queue), so the host-side bitmap is current
    return aloha_spm_written(E, op.spm_addr, words, written);
}

// Ops [first, first + count) in one call: op first + j uses dumps + j*4n, written + j*4n, sub_dumps + j*4n,
// sub_written + j*4n and has_sub[j].  (The per-op loop of run(), top_noaxilite_tb.sv:596-638, on this side of
// the boundary: a caller in another language pays one FFI crossing per program, not per op.)
//
// PROGRAM-level scheduling.  The testbench runs one kernel per op and dumps after each; here
// consecutive compute ops (encode_post, mul_plain, hom_add, rotate) are collected and handed to the engine as ONE
// aloha_run_vp_multi, which levels them by their true dependencies -- the eight encode_post kernels of case2 become
// one transform launch of sixteen jobs, four rotates share their eleven launches -- and their dumps are read back
// after the batch.  That is the same as dumping after each op as long as nobody in the batch writes the rows another
// member dumps, so an op joins the batch only if its dump rows are disjoint from those of every member; an encode
// op's upload (and the dump of the raw encoder output) happens at once, ahead of the members, which is the program's
// order as long as the members touch none of those rows.  Loads and stores close the batch.
int aloha_host_run_range_async(aloha_host_t *H, uint32_t first, uint32_t count, uint64_t *dumps, uint8_t *written,
                               uint64_t *sub_dumps, uint8_t *sub_written, int *has_sub) {
    if (!H || (uint64_t)first + count > H->ops.size() || !dumps || !written || !sub_dumps || !sub_written || !has_sub)
        return ALOHA_E_ARG;
    const uint64_t w = 4ull * H->n, bytes = w * 8;
    const uint32_t ct_rows = (uint32_t)(w / kLanes), pt_rows = ct_rows / 2;
    aloha_t *E = H->eng;
    struct Member { uint32_t j, pc; aloha_vp_args a; uint32_t out; uint32_t src[2], src_rows[2]; };
    std::vector<Member> batch;
    auto apart = [](uint32_t a, uint32_t an, uint32_t b, uint32_t bn) { return a + an <= b || b + bn <= a; };
    auto flush = [&]() -> int {
        if (batch.empty()) return ALOHA_OK;
        std::vector<uint32_t> pcs;
        std::vector<aloha_vp_args> args;
        for (const Member &m : batch) { pcs.push_back(m.pc); args.push_back(m.a); }
        int rc = aloha_run_vp_multi(E, (uint32_t)batch.size(), pcs.data(), args.data());
        for (size_t k = 0; k < batch.size() && !rc; ++k) {
            const Member &m = batch[k];
            rc = read_back_async(H, m.out, bytes, dumps + m.j * w);
            if (!rc) rc = aloha_spm_written(E, m.out, w, written + m.j * w);
        }
        batch.clear();
        return rc;
    };
    for (uint32_t j = 0; j < count; ++j) {
        const uint32_t i = first + j;
        const HostOp &op = H->ops[i];
        has_sub[j] = 0;
        Member m{j, 0, aloha_vp_args{0, 0, 0, 0, 0}, op.spm_addr, {0, 0}, {0, 0}};
        switch (op.type) {
        case OP_ENCODE:
            m.pc = kPcEncodePost; m.a = aloha_vp_args{op.spm_addr, 0, op.spm_addr, 0, 0};
            m.src[0] = op.spm_addr; m.src_rows[0] = pt_rows;
            break;
        case OP_MUL_PLAIN:
            m.pc = kPcMulPlain; m.a = aloha_vp_args{op.src1, op.src2, op.spm_addr, 0, 0};
            m.src[0] = op.src1; m.src_rows[0] = ct_rows; m.src[1] = op.src2; m.src_rows[1] = pt_rows;
            break;
        case OP_HOM_ADD:
            m.pc = kPcHomAdd; m.a = aloha_vp_args{op.src1, op.src2, op.spm_addr, 0, 0};
            m.src[0] = op.src1; m.src_rows[0] = ct_rows; m.src[1] = op.src2; m.src_rows[1] = ct_rows;
            break;
        case OP_ROTATE:
            m.pc = kPcKeyswitch;
            m.a = aloha_vp_args{op.src1, 0, op.spm_addr, (uint32_t)((clog2(op.step) - 1) * H->n * 12 / kLanes),
                                (uint32_t)pow3_mod(op.step, 2ull * H->n)};
            m.src[0] = op.src1; m.src_rows[0] = ct_rows;
            break;
        default: {                                        // load / store (or an op the single-op path rejects): in order
            int rc = flush();
            if (!rc) rc = aloha_host_run_op_async(H, i, dumps + j * w, written + j * w, sub_dumps + j * w, sub_written + j * w, has_sub + j);
            if (rc) return rc;
            continue;
        }
        }
        bool joins = batch.size() < 64;
        for (const Member &b : batch) {
            if (!apart(m.out, ct_rows, b.out, ct_rows)) joins = false;
            if (op.type == OP_ENCODE)                     // the upload overtakes the members: they must not read those rows
                for (int s = 0; s < 2; ++s)
                    if (b.src_rows[s] && !apart(m.out, pt_rows, b.src[s], b.src_rows[s])) joins = false;
        }
        if (!joins) {
            int rc = flush();
            if (rc) return rc;
        }
        if (op.type == OP_ENCODE) {
            auto it = H->encoder.find(i);
            if (it == H->encoder.end()) return ALOHA_E_STATE;
            int rc = aloha_dma_mem_h2d(E, op.spm_addr, it->second.data(), it->second.size() * 8);
            if (!rc) rc = read_back_async(H, op.spm_addr, bytes, sub_dumps + j * w);
            if (!rc) rc = aloha_spm_written(E, op.spm_addr, w, sub_written + j * w);
            if (rc) return rc;
            has_sub[j] = 1;
        }
        batch.push_back(m);
    }
    return flush();
}

int aloha_host_run_op(aloha_host_t *H, uint32_t i, uint64_t *dump, uint8_t *written, uint64_t *sub_dump,
                      uint8_t *sub_written, int *has_sub) {
    if (!H || i >= H->ops.size()) return ALOHA_E_ARG;
    if (H->in_flight) {                                   // mixing with the asynchronous flavour: drain first
        int rc0 = aloha_host_sync(H);
        if (rc0) return rc0;
    }
    const bool want_dump = dump && written;   // NULL dump = run the op only (no read-back, no sync)
    const HostOp &op = H->ops[i];
    const uint64_t words = 4ull * H->n, bytes = words * 8;
    aloha_t *E = H->eng;
    int rc = ALOHA_OK;
    if (has_sub) *has_sub = 0;
    switch (op.type) {
    case OP_LOAD: {
        const uint64_t a = kDramVpBase + op.dram_addr;
        if (a + bytes > H->dram.size()) return ALOHA_E_RANGE;
        rc = aloha_dma_mem_h2d(E, op.spm_addr, (const uint64_t *)(H->dram.data() + a), bytes);
        break;
    }
    case OP_STORE: {
        const uint64_t a = kDramVpBase + op.dram_addr;
        if (a + bytes > H->dram.size()) return ALOHA_E_RANGE;
        rc = aloha_dma_mem_d2h(E, (uint64_t *)(H->dram.data() + a), op.spm_addr, bytes);
        if (rc || !want_dump) return rc;
        std::memcpy(dump, H->dram.data() + a, bytes);   // the TB dumps straight from DRAM (:608-611)
        return aloha_spm_written(E, op.spm_addr, words, written);
    }
    case OP_ENCODE: {
        auto it = H->encoder.find(i);
        if (it == H->encoder.end()) return ALOHA_E_STATE;   // encoder output must be injected
        rc = aloha_dma_mem_h2d(E, op.spm_addr, it->second.data(), it->second.size() * 8);
        if (rc) return rc;
        if (sub_dump && sub_written) {
            rc = aloha_dma_mem_d2h(E, sub_dump, op.spm_addr, bytes);
            if (!rc) rc = aloha_spm_written(E, op.spm_addr, words, sub_written);
            if (rc) return rc;
            if (has_sub) *has_sub = 1;
        }
        rc = aloha_run_vp(E, kPcEncodePost, op.spm_addr, 0, op.spm_addr, 0, 0);
        break;
    }
    case OP_MUL_PLAIN: rc = aloha_run_vp(E, kPcMulPlain, op.src1, op.src2, op.spm_addr, 0, 0); break;
    case OP_HOM_ADD: rc = aloha_run_vp(E, kPcHomAdd, op.src1, op.src2, op.spm_addr, 0, 0); break;
    case OP_ROTATE: {
        // tb:530-532: galois element 3^step mod 2N in the step CSR, key at (clog2(step)-1)*12N/128 rows
        const uint32_t k = (uint32_t)pow3_mod(op.step, 2ull * H->n);
        const uint32_t ksk_ptr = (clog2(op.step) - 1) * H->n * 12 / kLanes;
        rc = aloha_run_vp(E, kPcKeyswitch, op.src1, 0, op.spm_addr, ksk_ptr, k);
        break;
    }
    default: return ALOHA_E_OPCODE;   // encode_post as a host op is parsed but never run by the TB
    }
    if (rc || !want_dump) return rc;
    rc = aloha_dma_mem_d2h(E, dump, op.spm_addr, bytes);
    if (rc) return rc;
    return aloha_spm_written(E, op.spm_addr, words, written);
}

int aloha_write_dump_text(const char *path, const uint64_t *data, const uint8_t *written, uint64_t nwords) {
    if (!path || !data) return ALOHA_E_ARG;
    FILE *f = std::fopen(path, "w");
    if (!f) return ALOHA_E_ARG;
    for (uint64_t i = 0; i < nwords; ++i) {
        if (written && !written[i]) std::fputs("x\n", f);
        else std::fprintf(f, "%llu\n", (unsigned long long)data[i]);
    }
    std::fclose(f);
    return ALOHA_OK;
}

}  // extern "C"
