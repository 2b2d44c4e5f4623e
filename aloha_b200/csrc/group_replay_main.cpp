// group_replay_main.cpp -- a host program that knows only include/aloha_b200.h: it replays a limb-sharded
// instruction-stream case (config 5: the key-switch / relinearise streams) on n GPUs of one box, one engine
// per GPU in ONE process (aloha_group_create_local), without Python.
//
// It plays the role of the reference's testbench process (sim/top/top_noaxilite_tb.sv:596-638 run():
// load the ROM images, DMA the inputs, walk an op list, dump the results) for a machine of several chips.
// The case directory is written by aloha_b200.hks.write_replay_case (the stream generator); file formats
// are described there.
//
//   aloha_group_replay <case-dir> <nranks> [first-device]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/aloha_b200.h"

namespace {

struct Op {
    std::string kind;
    std::vector<uint32_t> pcs;
    std::vector<aloha_vp_args> args;
    long v[5] = {0, 0, 0, 0, 0};
};

[[noreturn]] void die(const std::string &what) {
    std::fprintf(stderr, "aloha_group_replay: %s\n", what.c_str());
    std::exit(1);
}

void check(int rc, aloha_t *e, const char *what) {
    if (rc) die(std::string(what) + ": " + aloha_strerror(rc) + " -- " + (e ? aloha_last_error(e) : ""));
}

std::vector<uint8_t> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) die("cannot read " + path);
    return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

std::vector<Op> read_prog(const std::string &path) {
    std::ifstream f(path);
    if (!f) die("cannot read " + path);
    std::vector<Op> ops;
    std::string kind;
    while (f >> kind) {
        Op op;
        op.kind = kind;
        if (kind == "run") {
            size_t n;
            f >> n;
            op.pcs.resize(n);
            op.args.resize(n);
            for (size_t i = 0; i < n; ++i)
                f >> op.pcs[i] >> op.args[i].src0 >> op.args[i].src1 >> op.args[i].rslt >> op.args[i].ksk_ptr >> op.args[i].step;
        } else if (kind == "allgather") {
            f >> op.v[0] >> op.v[1] >> op.v[2] >> op.v[3] >> op.v[4];
        } else if (kind == "bcast") {
            f >> op.v[0] >> op.v[1] >> op.v[2];
        } else if (kind == "wait") {
            f >> op.v[0];
        } else {
            die("unknown op '" + kind + "' in " + path);
        }
        if (!f) die("malformed " + path);
        ops.push_back(op);
    }
    return ops;
}

}  // namespace

int main(int argc, char **argv) {
    if (argc < 3) die("usage: aloha_group_replay <case-dir> <nranks> [first-device]");
    const std::string dir = argv[1];
    const int n = std::atoi(argv[2]), dev0 = argc > 3 ? std::atoi(argv[3]) : 0;
    std::vector<aloha_t *> eng(n, nullptr);
    std::vector<std::vector<Op>> prog(n);
    for (int r = 0; r < n; ++r) {
        const std::string base = dir + "/rank" + std::to_string(r);
        std::ifstream cfgf(base + ".cfg");
        if (!cfgf) die("cannot read " + base + ".cfg");
        aloha_cfg cfg;
        std::memset(&cfg, 0, sizeof cfg);
        size_t nmod;
        cfgf >> cfg.vlmax_bits >> cfg.spm_rows >> cfg.ksk_rows >> cfg.pool_buffers >> cfg.isram_depth >> nmod;
        std::vector<uint64_t> q(nmod), psi(nmod);
        for (size_t i = 0; i < nmod; ++i) cfgf >> q[i] >> psi[i];
        if (!cfgf) die("malformed " + base + ".cfg");
        cfg.device = dev0 + r;
        const int created = aloha_create(&cfg, &eng[r]);      // (eng[r] is read after the call: it carries the error text)
        check(created, eng[r], "aloha_create");
        check(aloha_load_tf_rom(eng[r], q.data(), psi.data(), (uint32_t)nmod), eng[r], "aloha_load_tf_rom");
        const std::vector<uint8_t> rom = slurp(base + ".isram");
        check(aloha_load_isram(eng[r], rom.data(), (uint32_t)(rom.size() / 12), 0), eng[r], "aloha_load_isram");
        std::ifstream io(base + ".io");
        std::string kind, name;
        while (io >> kind) {
            if (kind == "dump") { long a, b; io >> a >> b >> name; continue; }
            uint32_t row;
            io >> row >> name;
            const std::vector<uint8_t> data = slurp(dir + "/" + name);
            if (kind == "spm") check(aloha_dma_mem_h2d(eng[r], row, (const uint64_t *)data.data(), data.size()), eng[r], "aloha_dma_mem_h2d");
            else if (kind == "ksk") check(aloha_dma_ksk_h2d(eng[r], row, (const uint64_t *)data.data(), data.size()), eng[r], "aloha_dma_ksk_h2d");
            else die("unknown io line '" + kind + "'");
        }
        check(aloha_sync(eng[r]), eng[r], "aloha_sync");
        prog[r] = read_prog(base + ".prog");
        if (prog[r].size() != prog[0].size()) die("the ranks' op lists differ in length");
    }
    aloha_group_t *grp = nullptr;
    int rc = aloha_group_create_local(eng.data(), n, &grp);
    if (rc) die(std::string("aloha_group_create_local: ") + aloha_group_last_error(grp));

    // the op lists in lockstep: run ops go to each engine, a collective is issued once for the group
    for (size_t s = 0; s < prog[0].size(); ++s) {
        const Op &op = prog[0][s];
        for (int r = 0; r < n; ++r)
            if (prog[r][s].kind != op.kind) die("the ranks' op lists differ at step " + std::to_string(s));
        if (op.kind == "run") {
            for (int r = 0; r < n; ++r) {
                const Op &o = prog[r][s];
                if (!o.pcs.empty())
                    check(aloha_run_vp_multi(eng[r], (uint32_t)o.pcs.size(), o.pcs.data(), o.args.data()), eng[r], "aloha_run_vp_multi");
            }
        } else if (op.kind == "allgather") {
            rc = aloha_group_all_gather_rows(grp, (uint32_t)op.v[0], (uint32_t)op.v[1], (uint32_t)op.v[2], (uint32_t)op.v[3],
                                             op.v[4] ? ALOHA_GROUP_CHUNKED : 0);
            if (rc) die(std::string("aloha_group_all_gather_rows: ") + aloha_group_last_error(grp));
        } else if (op.kind == "bcast") {
            rc = aloha_group_broadcast_rows(grp, (uint32_t)op.v[0], (uint32_t)op.v[1], (int)op.v[2]);
            if (rc) die(std::string("aloha_group_broadcast_rows: ") + aloha_group_last_error(grp));
        } else {
            rc = aloha_group_wait(grp, (int)op.v[0]);
            if (rc) die(std::string("aloha_group_wait: ") + aloha_group_last_error(grp));
        }
    }
    for (int r = 0; r < n; ++r) {
        std::ifstream io(dir + "/rank" + std::to_string(r) + ".io");
        std::string kind, name;
        while (io >> kind) {
            if (kind != "dump") { long a; io >> a >> name; continue; }
            uint32_t row;
            uint64_t nwords;
            io >> row >> nwords >> name;
            std::vector<uint64_t> out(nwords);
            check(aloha_dma_mem_d2h(eng[r], out.data(), row, nwords * 8), eng[r], "aloha_dma_mem_d2h");
            std::ofstream f(dir + "/" + name, std::ios::binary);
            f.write((const char *)out.data(), (std::streamsize)(nwords * 8));
        }
    }
    aloha_stats st;
    uint64_t launches = 0;
    for (int r = 0; r < n; ++r) {
        check(aloha_get_stats(eng[r], &st), eng[r], "aloha_get_stats");
        launches += st.kernel_launches;
    }
    std::printf("aloha_group_replay: %d rank(s), %zu ops, %llu kernel launches\n", n, prog[0].size(), (unsigned long long)launches);
    aloha_group_destroy(grp);
    for (auto *e : eng) aloha_destroy(e);
    return 0;
}
