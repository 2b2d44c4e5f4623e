// sim_kernels.cpp -- the launch interface of aloha_b200/csrc/kernels.cuh on host memory (TEST INFRASTRUCTURE; see
// cuda_runtime.h in this directory).  Each launcher evaluates the CONTRACT its job structure documents, written
// here independently of the sm_100a kernels (128-bit host arithmetic, one element at a time), and first checks what
// the real kernels silently rely on:
//   * job tables, twiddle tables and every operand lie inside device allocations;
//   * a destination overlaps an operand only where the real kernel tolerates it -- exactly equal ranges for the
//     index-preserving kernels, never for the permuting ones (a thread there reads words another thread writes);
//   * the row groups / tensor-map coordinates / tile plans / Shoup companions the engine precomputes on the host
//     describe the same work as the plain job records.
// A broken contract makes the launch (and every later runtime call) fail, so it surfaces as ALOHA_E_CUDA.
// The transforms are evaluated as the reference's constant-geometry network with the RTL ALU (PeaseJob's formulas),
// which on in-domain inputs is what the fast kernels store, and on other inputs is what ALOHA_F_STRICT stores.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../aloha_b200/csrc/kernels.cuh"
#include "sim.hpp"

namespace alb {

std::atomic<unsigned long long> g_launches{0};
unsigned long long kernel_launch_count() { return g_launches.load(); }

namespace {

typedef unsigned __int128 u128;

// ---- the reference ALU (src/vp/vxu/modalu.sv:44-46,228-229,249; modmul.sv:150-252; halfred.sv:23-26)
inline u64 once(u64 x, u64 q) { return x >= q ? x - q : x; }
inline u64 barrett(u64 a, u64 b, u64 q, u64 iq) {
    const u128 prod = (u128)a * b;
    const u64 ps = (u64)(prod >> 58);
    const u64 ms = (u64)(((u128)ps * iq) >> 63);
    const u64 M = (1ull << 61) - 1;
    const u64 diff = ((((u64)prod & M) | (M + 1)) - ((ms * q) & M)) & M;
    return diff < q ? diff : diff - q;
}
inline u64 addm(u64 a, u64 b, u64 q) {
    const u128 s = (u128)a + b;
    return s >= q ? (u64)(s - q) : (u64)s;
}
inline u64 subm(u64 a, u64 b, u64 q) { return a >= b ? a - b : q + a - b; }
inline u64 half(u64 x, u64 q) { return (x >> 1) + ((x & 1) ? ((q + 1) >> 1) : 0); }
inline u64 alu(u32 op, u64 a_raw, u64 b_raw, u64 s, u64 q, u64 iq) {
    const u64 a = once(a_raw, q);
    switch (op) {
    case 0x00: return barrett(a, once(b_raw, q), q, iq);
    case 0x04: return barrett(a, s, q, iq);
    case 0x03: return barrett(a, 1, q, iq);
    case 0x01: return addm(a, once(b_raw, q), q);
    case 0x05: return addm(a, s, q);
    case 0x02: return subm(a, once(b_raw, q), q);
    case 0x06: return subm(a, s, q);
    case 0x0a: return subm(s, a, q);
    }
    return 0;
}
inline u64 barrett_constant(u64 q) { return (u64)((((u128)1) << 121) / q); }

// ---- contract checks
struct Check {
    std::string what;
    bool ok = true;
    void fail(const std::string &m) { if (ok) { ok = false; what = m; } }
    std::vector<std::pair<const u64 *, u64>> reads;       // operand ranges, for the cross-job check of a launch
    void dev(const void *p, u64 words, const char *name) {
        if (!p || !sim::device_range(p, words * 8)) fail(std::string(name) + " is not inside a device allocation");
    }
    void in(const u64 *p, u64 words, const char *name) { dev(p, words, name); reads.emplace_back(p, words); }
    static bool overlap(const u64 *a, u64 an, const u64 *b, u64 bn) { return a < b + bn && b < a + an; }
    // index-preserving kernels: element i of dst depends on element i of the operand only
    void same_or_apart(const u64 *dst, const u64 *x, u64 n, const char *name) {
        if (x && x != dst && overlap(dst, n, x, n)) fail(std::string("dst partially overlaps ") + name);
    }
    // permuting kernels: element d of dst depends on some other element of the operand
    void apart(const u64 *dst, const u64 *x, u64 n, const char *name) {
        if (x && overlap(dst, n, x, n)) fail(std::string("dst overlaps the permuted operand ") + name);
    }
};

template <class Job, class Fn>
cudaError_t run(cudaStream_t st, const char *kernel, const Job *jobs, u32 njobs, unsigned launches, Fn fn) {
    g_launches += launches;
    static const bool trace = std::getenv("ALOHA_SIM_TRACE") != nullptr;     // one line per launch, for reading schedules
    if (trace) std::fprintf(stderr, "sim launch %s jobs %u stream %p\n", kernel, njobs, (void *)st);
    if (!njobs) return cudaErrorInvalidValue;                 // a zero-sized grid is a launch error on the device too
    static const bool null_device = std::getenv("ALOHA_SIM_NULL") != nullptr;   // launches cost nothing and do nothing:
    if (null_device) return cudaSuccess;                                          // what is left is the host's own time
    sim::enqueue(st, [=]() {
        if (!sim::device_range(jobs, (size_t)njobs * sizeof(Job))) { sim::violation((std::string(kernel) + ": job table is not device memory").c_str()); return; }
        // every job is checked before any runs, and all of a launch's jobs read their operands before any writes
        // (jobs of one launch are unordered on the device: a job must not consume another job's output)
        std::vector<std::vector<u64>> out(njobs);
        std::vector<u64 *> dst(njobs);
        std::vector<u64> len(njobs);
        std::vector<Check> checks(njobs);
        for (u32 j = 0; j < njobs; ++j) {
            Check &c = checks[j];
            fn(jobs[j], c, out[j], dst[j]);
            if (!c.ok) { sim::violation((std::string(kernel) + " job " + std::to_string(j) + ": " + c.what).c_str()); return; }
            len[j] = out[j].size();
        }
        for (u32 j = 0; j < njobs; ++j)
            for (u32 i = 0; i < njobs; ++i) {
                if (i == j) continue;
                if (i < j && Check::overlap(dst[j], len[j], dst[i], len[i])) { sim::violation((std::string(kernel) + ": two jobs of one launch write the same words").c_str()); return; }
                for (auto &r : checks[i].reads)
                    if (Check::overlap(dst[j], len[j], r.first, r.second)) { sim::violation((std::string(kernel) + ": a job reads words another job of the same launch writes").c_str()); return; }
            }
        sim::access(jobs, (size_t)njobs * sizeof(Job), false, kernel);
        for (u32 j = 0; j < njobs; ++j) {
            for (auto &r : checks[j].reads) sim::access(r.first, r.second * 8, false, kernel);
            sim::access(dst[j], len[j] * 8, true, kernel);
            std::memcpy(dst[j], out[j].data(), len[j] * 8);
        }
    });
    return sim::status();
}

// One stage of the constant-geometry network (kernels.cuh PeaseJob)
void stage_fwd(const std::vector<u64> &x, std::vector<u64> &y, const Tw *tw, u32 logn, u32 s, u64 q, u64 iq) {
    const u32 h = 1u << (logn - 1), m = 1u << s;
    for (u32 p = 0; p < h; ++p) {
        const u64 w = once(tw[m + (p & (m - 1))].w, q);
        const u64 a = once(x[p], q), b = once(x[p + h], q), t = barrett(b, w, q, iq);
        y[2 * p] = addm(a, t, q);
        y[2 * p + 1] = subm(a, t, q);
    }
}
void stage_inv(const std::vector<u64> &x, std::vector<u64> &y, const Tw *tw, u32 logn, u32 s, u64 q, u64 iq) {
    const u32 h = 1u << (logn - 1), m = 1u << (logn - 1 - s);
    for (u32 p = 0; p < h; ++p) {
        const u64 w = once(tw[m + (p & (m - 1))].w, q);
        const u64 a = once(x[2 * p], q), b = once(x[2 * p + 1], q);
        y[p] = half(addm(a, b, q), q);
        y[p + h] = half(barrett(subm(a, b, q), w, q, iq), q);
    }
}

void transform(const NttJob &job, u32 logn, u32 form, bool inverse, Check &c, std::vector<u64> &out, u64 *&dst) {
    const u32 n = 1u << logn;
    const u64 q = job.mc.q, iq = barrett_constant(q);
    c.in(job.src, n, "src"); c.dev(job.dst, n, "dst"); c.dev(job.tw, 2 * n, "twiddle table"); c.dev(job.rtw, 2 * n, "row-order twiddle table");
    c.same_or_apart(job.dst, job.src, n, "src");
    if (job.mc.form != form) c.fail("job of another modulus form in this launch");
    if (form == FORM_PM && (job.mc.d != (1ull << 60) - q || job.mc.q3 != 3 * q)) c.fail("pseudo-Mersenne constants do not belong to q");
    if (job.mc.mest != (u32)((((u128)1) << 91) / q)) c.fail("mest is not floor(2^91 / q)");
    if (job.mc.pre > PRE_VFQMOD) c.fail("unknown load op");
    if (!c.ok) return;
    std::vector<u64> x(job.src, job.src + n), y(n);
    for (u64 &v : x) {
        if (job.mc.pre == PRE_VCPY) v = alu(0x05, v, 0, 0, q, iq);
        else if (job.mc.pre == PRE_VFQMOD) v = alu(0x03, v, 0, 0, q, iq);
    }
    for (u32 s = 0; s < logn; ++s) {
        if (inverse) stage_inv(x, y, job.tw, logn, s, q, iq); else stage_fwd(x, y, job.tw, logn, s, q, iq);
        x.swap(y);
    }
    out = std::move(x);
    dst = job.dst;
}

// the 16-job records of the TMA-staged row passes must describe the jobs they shadow
void check_groups(const NttJob *jobs, const NttRowGroup *groups, u32 ngroups, u32 logn, bool inverse, const TmaMaps *maps, Check &c) {
    if (!ngroups) return;
    if (!sim::device_range(groups, (size_t)ngroups * sizeof(NttRowGroup))) { c.fail("row groups are not device memory"); return; }
    for (u32 g = 0; g < ngroups && c.ok; ++g)
        for (u32 k = 0; k < 16; ++k) {
            const NttJob &j = jobs[16 * g + k];
            const NttRowGroup &G = groups[g];
            if (G.dst[k] != j.dst) c.fail("row group dst differs from its job");
            const u64 *want_src = inverse ? j.src : (logn > 8 ? j.dst : j.src);
            if (G.src[k] != want_src) c.fail("row group src is not what the row pass reads");
            if (G.rtw != j.rtw || G.mc.q != j.mc.q || G.mc.pre != j.mc.pre || G.mc.form != j.mc.form) c.fail("row group of mixed moduli / load ops");
            if (inverse) {
                if (!maps || G.src_map[k] > 2) { c.fail("row group without a tensor map"); continue; }
                const u64 *base = (const u64 *)(uintptr_t)maps->m[G.src_map[k]].opaque[0];
                if (base + (u64)G.src_line[k] * 16 != j.src) c.fail("tensor-map coordinates do not address the job's source");
            }
        }
}

}  // namespace

cudaError_t launch_ntt_forward(const NttJob *jobs, u32 njobs, const NttRowGroup *groups, u32 ngroups, u32 logn, u32 form, cudaStream_t st) {
    if (16 * ngroups > njobs || logn < 8 || logn > 16 || form > FORM_PM) return cudaErrorInvalidValue;
    sim::enqueue(st, [=]() { Check c; if (sim::device_range(jobs, (size_t)njobs * sizeof(NttJob))) check_groups(jobs, groups, ngroups, logn, false, nullptr, c); if (!c.ok) sim::violation(("ntt_forward: " + c.what).c_str()); });
    const unsigned launches = (logn > 8) + (ngroups != 0) + (njobs > 16 * ngroups);
    return run(st, "ntt_forward", jobs, njobs, launches, [=](const NttJob &j, Check &c, std::vector<u64> &o, u64 *&d) { transform(j, logn, form, false, c, o, d); });
}
cudaError_t launch_ntt_inverse(const NttJob *jobs, u32 njobs, const NttRowGroup *groups, u32 ngroups, const TmaMaps *maps, u32 logn, u32 form, cudaStream_t st) {
    if (16 * ngroups > njobs || (ngroups && !maps) || logn < 8 || logn > 16 || form > FORM_PM) return cudaErrorInvalidValue;
    const TmaMaps held = maps ? *maps : TmaMaps{};          // passed by value to the real kernel
    sim::enqueue(st, [=]() { Check c; if (sim::device_range(jobs, (size_t)njobs * sizeof(NttJob))) check_groups(jobs, groups, ngroups, logn, true, &held, c); if (!c.ok) sim::violation(("ntt_inverse: " + c.what).c_str()); });
    const unsigned launches = (logn > 8) + (ngroups != 0) + (njobs > 16 * ngroups);
    return run(st, "ntt_inverse", jobs, njobs, launches, [=](const NttJob &j, Check &c, std::vector<u64> &o, u64 *&d) { transform(j, logn, form, true, c, o, d); });
}

cudaError_t launch_pease(const PeaseJob *jobs, u32 njobs, u32 logn, u32 stage, bool inverse, cudaStream_t st) {
    if (stage >= logn) return cudaErrorInvalidValue;
    return run(st, "pease", jobs, njobs, 1, [=](const PeaseJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        const u32 n = 1u << logn;
        c.in(j.src, n, "src"); c.dev(j.dst, n, "dst"); c.dev(j.tw, 2 * n, "twiddle table");
        c.apart(j.dst, j.src, n, "src");
        if (!c.ok) return;
        std::vector<u64> x(j.src, j.src + n);
        o.resize(n);
        if (inverse) stage_inv(x, o, j.tw, logn, stage, j.q, j.iq); else stage_fwd(x, o, j.tw, logn, stage, j.q, j.iq);
        d = j.dst;
    });
}

cudaError_t launch_ew(u32 op, const EwJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    const bool vv = op == 0x00 || op == 0x01 || op == 0x02;
    if (!(vv || op == 0x04 || op == 0x05 || op == 0x06 || op == 0x0a || op == 0x03)) return cudaErrorInvalidValue;
    return run(st, "ew", jobs, njobs, 1, [=](const EwJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.in(j.a, n, "a"); c.dev(j.dst, n, "dst");
        c.same_or_apart(j.dst, j.a, n, "a");
        if (vv) { c.in(j.b, n, "b"); c.same_or_apart(j.dst, j.b, n, "b"); }
        if (!c.ok) return;
        o.resize(n);
        for (u32 i = 0; i < n; ++i) o[i] = alu(op, j.a[i], vv ? j.b[i] : 0, j.s, j.q, j.iq);
        d = j.dst;
    });
}

cudaError_t launch_copy(const CopyJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "copy", jobs, njobs, 1, [=](const CopyJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.in(j.src, n, "src"); c.dev(j.dst, n, "dst");
        c.same_or_apart(j.dst, j.src, n, "src");
        if (!c.ok) return;
        o.assign(j.src, j.src + n);
        d = j.dst;
    });
}

namespace {
// dst[(i k) mod n] = ((i k) mod 2n >= n) ? q - x[i] : x[i]   (raw subtraction: 0 becomes q)
void automorph(const u64 *x, u32 n, u64 k, u64 q, std::vector<u64> &o) {
    o.resize(n);
    for (u32 i = 0; i < n; ++i) {
        const u64 ik = (u64)i * k;
        o[ik & (n - 1)] = (ik & (2ull * n - 1)) >= n ? q - x[i] : x[i];
    }
}
bool odd_and_inverse(u64 k, u64 kinv, u32 n) { return (k & 1) && ((k * kinv) & (n - 1)) == 1; }
}  // namespace

cudaError_t launch_vaut(const PermJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "vaut", jobs, njobs, 1, [=](const PermJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.in(j.src, n, "src"); c.dev(j.dst, n, "dst"); c.apart(j.dst, j.src, n, "src");
        if (!odd_and_inverse(j.k, j.kinv, n)) c.fail("kinv is not k^-1 mod n (or k is even)");
        if (!c.ok) return;
        automorph(j.src, n, j.k, j.q, o);
        d = j.dst;
    });
}
cudaError_t launch_vaut_tiled(const AutJob *jobs, u32 njobs, u32 n, u32 max_tiles, cudaStream_t st) {
    return run(st, "vaut_tiled", jobs, njobs, 1, [=](const AutJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.in(j.src, n, "src"); c.dev(j.dst, n, "dst"); c.apart(j.dst, j.src, n, "src");
        const AutPlan want = make_aut_plan(n, j.k);
        if (std::memcmp(&want, &j.plan, sizeof want) != 0) c.fail("tile plan is not make_aut_plan(n, k)");
        if (j.plan.ntiles > max_tiles) c.fail("the grid does not cover the job's tiles");
        if (!(j.k & 1)) c.fail("even Galois element");
        if (!c.ok) return;
        automorph(j.src, n, j.k, j.q, o);
        d = j.dst;
    });
}
cudaError_t launch_vroli(const PermJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "vroli", jobs, njobs, 1, [=](const PermJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.in(j.src, n, "src"); c.dev(j.dst, n, "dst"); c.apart(j.dst, j.src, n, "src");
        if (!c.ok) return;
        o.resize(n);
        for (u32 i = 0; i < n; ++i) o[i] = j.src[(i + (u32)j.kinv) & (n - 1)];
        d = j.dst;
    });
}

namespace {
void autmac(const AutMacJob &j, u32 n, bool tiled, u32 max_tiles, Check &c, std::vector<u64> &o, u64 *&d) {
    c.in(j.x, n, "x"); c.in(j.p, n, "p"); c.in(j.c, n, "c"); c.dev(j.dst, n, "dst");
    c.apart(j.dst, j.x, n, "x");
    c.same_or_apart(j.dst, j.p, n, "p"); c.same_or_apart(j.dst, j.c, n, "c");
    if (!odd_and_inverse(j.k, j.kinv, n)) c.fail("kinv is not k^-1 mod n (or k is even)");
    if (tiled) {
        const AutPlan want = make_aut_plan(n, j.k);
        if (std::memcmp(&want, &j.plan, sizeof want) != 0) c.fail("tile plan is not make_aut_plan(n, k)");
        if (j.plan.ntiles > max_tiles) c.fail("the grid does not cover the job's tiles");
    }
    if (!c.ok) return;
    std::vector<u64> r;
    automorph(j.x, n, j.k, j.q, r);
    o.resize(n);
    for (u32 i = 0; i < n; ++i) o[i] = alu(0x01, j.c[i], alu(0x00, r[i], j.p[i], 0, j.q, j.iq), 0, j.q, j.iq);
    d = j.dst;
}
}  // namespace
cudaError_t launch_autmac(const AutMacJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "autmac", jobs, njobs, 1, [=](const AutMacJob &j, Check &c, std::vector<u64> &o, u64 *&d) { autmac(j, n, false, 0, c, o, d); });
}
cudaError_t launch_autmac_tiled(const AutMacJob *jobs, u32 njobs, u32 n, u32 max_tiles, cudaStream_t st) {
    return run(st, "autmac_tiled", jobs, njobs, 1, [=](const AutMacJob &j, Check &c, std::vector<u64> &o, u64 *&d) { autmac(j, n, true, max_tiles, c, o, d); });
}

cudaError_t launch_mac(const MacJob *jobs, u32 njobs, u32 terms, u32 n, cudaStream_t st) {
    if (terms < 1 || terms > 4) return cudaErrorInvalidValue;
    return run(st, "mac", jobs, njobs, 1, [=](const MacJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.dev(j.dst, n, "dst");
        for (u32 t = 0; t < terms; ++t) { c.in(j.a[t], n, "a"); c.in(j.b[t], n, "b"); c.same_or_apart(j.dst, j.a[t], n, "a"); c.same_or_apart(j.dst, j.b[t], n, "b"); }
        if (!c.ok) return;
        o.resize(n);
        for (u32 i = 0; i < n; ++i) {
            u64 acc = 0;
            for (u32 t = 0; t < terms; ++t) {
                const u64 m = alu(0x00, j.a[t][i], j.b[t][i], 0, j.q, j.iq);
                acc = t ? addm(acc, m, j.q) : m;
            }
            o[i] = acc;
        }
        d = j.dst;
    });
}
cudaError_t launch_muladd(const MulAddJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "muladd", jobs, njobs, 1, [=](const MulAddJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.in(j.a, n, "a"); c.in(j.b, n, "b"); c.in(j.c, n, "c"); c.dev(j.dst, n, "dst");
        c.same_or_apart(j.dst, j.a, n, "a"); c.same_or_apart(j.dst, j.b, n, "b"); c.same_or_apart(j.dst, j.c, n, "c");
        if (!c.ok) return;
        o.resize(n);
        for (u32 i = 0; i < n; ++i) o[i] = alu(0x01, j.c[i], alu(0x00, j.a[i], j.b[i], 0, j.q, j.iq), 0, j.q, j.iq);
        d = j.dst;
    });
}
cudaError_t launch_sop(const SopJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "sop", jobs, njobs, 1, [=](const SopJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.dev(j.dst, n, "dst");
        if (!j.terms) c.fail("no terms");
        if (((uintptr_t)j.pairs & 15) || !sim::device_range(j.pairs, (size_t)j.terms * 16)) { c.fail("pointer table is not 16-byte aligned device memory"); return; }
        for (u32 t = 0; t < 2 * j.terms; ++t) { c.in(j.pairs[t], n, "operand"); c.same_or_apart(j.dst, j.pairs[t], n, "an operand"); }
        if (!c.ok) return;
        o.resize(n);
        for (u32 i = 0; i < n; ++i) {
            u64 acc = 0;
            for (u32 t = 0; t < j.terms; ++t) {
                const u64 m = alu(0x00, j.pairs[2 * t][i], j.pairs[2 * t + 1][i], 0, j.q, j.iq);
                acc = t ? alu(0x01, acc, m, 0, j.q, j.iq) : m;
            }
            o[i] = acc;
        }
        d = j.dst;
    });
}
cudaError_t launch_bext(const BextJob *jobs, u32 njobs, u32 n, cudaStream_t st) {
    return run(st, "bext", jobs, njobs, 1, [=](const BextJob &j, Check &c, std::vector<u64> &o, u64 *&d) {
        c.dev(j.dst, n, "dst");
        if (!j.nterms) c.fail("no terms");
        if (!sim::device_range(j.terms, (size_t)j.nterms * sizeof(BextTerm))) { c.fail("term table is not device memory"); return; }
        for (u32 t = 0; t < j.nterms; ++t) {
            const BextTerm &tm = j.terms[t];
            c.in(tm.x, n, "x"); c.same_or_apart(j.dst, tm.x, n, "x");
            if (tm.pre > PRE_VFQMOD) c.fail("unknown pre-op");
            if (j.fast && (tm.s >= j.q || tm.sp != (u64)((((u128)tm.s) << 64) / j.q))) c.fail("fast path: scalar not below q or wrong Shoup companion");
        }
        if (j.fast && (j.q >> 59) != 1) c.fail("fast path: q is not a 60-bit modulus");
        if (j.fast && j.iq != barrett_constant(j.q)) c.fail("fast path: iq is not q's Barrett constant");
        if (j.fast && j.mest != (u32)((((u128)1) << 91) / j.q)) c.fail("fast path: mest is not floor(2^91 / q)");
        if (j.post > 1) c.fail("unknown post-op");
        if (!c.ok) return;
        o.resize(n);
        for (u32 i = 0; i < n; ++i) {
            u64 acc = 0;
            for (u32 t = 0; t < j.nterms; ++t) {
                const BextTerm &tm = j.terms[t];
                u64 a = tm.x[i];
                if (tm.pre == PRE_VCPY) a = alu(0x05, a, 0, 0, j.q, j.iq);
                else if (tm.pre == PRE_VFQMOD) a = alu(0x03, a, 0, 0, j.q, j.iq);
                const u64 m = alu(0x04, a, 0, tm.s, j.q, j.iq);
                acc = t ? alu(0x01, acc, m, 0, j.q, j.iq) : m;
            }
            if (j.post == 1) acc = alu(0x06, acc, 0, j.post_s, j.q, j.iq);
            o[i] = acc;
        }
        d = j.dst;
    });
}

}  // namespace alb

// ---- self-tests of the checks above (tests/test_sim_engine.py): each returns the launch's status
namespace {
struct Arena {
    alb::u64 *buf = nullptr;
    alb::EwJob *ew = nullptr;
    alb::PermJob *perm = nullptr;
    Arena() {
        cudaMalloc((void **)&buf, 4 * 256 * 8);
        cudaMalloc((void **)&ew, 2 * sizeof(alb::EwJob));
        cudaMalloc((void **)&perm, sizeof(alb::PermJob));
        for (int i = 0; i < 4 * 256; ++i) buf[i] = (alb::u64)i;
    }
};
const alb::u64 kQ = 576460825317867521ull, kIq = 0x3fffff78000120f7ull;
}  // namespace
extern "C" {
int sim_test_legal() {
    Arena a;
    a.ew[0] = alb::EwJob{a.buf, a.buf, a.buf + 256, 0, kQ, kIq};             // in place on an operand: allowed
    a.ew[1] = alb::EwJob{a.buf + 512, a.buf + 768, a.buf + 256, 0, kQ, kIq};
    return alb::launch_ew(0x01, a.ew, 2, 256, nullptr);
}
int sim_test_vaut_in_place() {
    Arena a;
    *a.perm = alb::PermJob{a.buf, a.buf, kQ, 3, 171};                        // 3 * 171 = 513 = 1 mod 256
    return alb::launch_vaut(a.perm, 1, 256, nullptr);
}
int sim_test_cross_job_read() {
    Arena a;
    a.ew[0] = alb::EwJob{a.buf, a.buf + 256, a.buf + 512, 0, kQ, kIq};
    a.ew[1] = alb::EwJob{a.buf + 768, a.buf, a.buf + 512, 0, kQ, kIq};       // reads job 0's destination
    return alb::launch_ew(0x01, a.ew, 2, 256, nullptr);
}
static int two_streams(bool ordered, bool through_host) {
    Arena a;
    cudaStream_t s1, s2;
    cudaEvent_t e;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    a.ew[0] = alb::EwJob{a.buf, a.buf + 256, a.buf + 512, 0, kQ, kIq};
    cudaMemsetAsync(a.buf + 256, 0, 256 * 8, s1);                            // s1 writes an operand ...
    if (ordered) { cudaEventRecord(e, s1); cudaStreamWaitEvent(s2, e, 0); }
    if (through_host) cudaStreamSynchronize(s1);
    alb::launch_ew(0x01, a.ew, 1, 256, s2);                                  // ... a kernel on s2 reads
    return (int)cudaStreamSynchronize(s2);
}
int sim_test_race() { return two_streams(false, false); }
int sim_test_race_ordered_by_event() { return two_streams(true, false); }
int sim_test_race_ordered_by_host() { return two_streams(false, true); }
int sim_test_foreign_pointer() {
    Arena a;
    static alb::u64 host_words[256];
    a.ew[0] = alb::EwJob{a.buf, host_words, a.buf + 512, 0, kQ, kIq};
    return alb::launch_ew(0x01, a.ew, 1, 256, nullptr);
}
}
