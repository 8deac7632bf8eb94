// join_exacts.hpp — C++ host-side mirror of the reference interface for the join path.
//
// The reference is Rust (enclone_ranger@cbd15ef, un-vendored; no Rust toolchain in the build image),
// so the host side above the C ABI is written in C++ with the reference's own names and argument
// meaning:
//
//   join_exacts(is_bcr, to_bc, refdata, ctl, exact_clonotypes, info, raw_joins, sr, dref) -> EquivRel
//
// It flattens (exact_clonotypes, info, refdata, dref) into the COMPACT form of ejoin_pack_t — the V..J
// bases straight from CloneInfo.tigsp (2 bits per base), the CDR3s 2 bits per base, one 64-byte record
// per entry: exactly what the Rust `Flat::pack` of INTEGRATION.md §2 does — on all host threads, calls
// libenclone_join.so, and rebuilds the outputs the callers consume: raw_joins
// (enclone_process::process_clonotypes, enclone_main/src/stop.rs:56), the EquivRel (orbits), and the
// PotentialJoin numbers behind join_info (enclone_tail/src/group.rs:70-74).
// Errors follow the reference's convention for this call (it cannot fail short of a panic):
// a library error throws std::runtime_error carrying ejoin_last_error().
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/enclone_join.h"

namespace enclone {

// crate `equiv`: the methods the callers use (enclone_tail/src/grouper.rs:406-445)
class EquivRel {
  public:
    explicit EquivRel(int32_t n = 0) : parent_(n) { for (int32_t i = 0; i < n; i++) parent_[i] = i; }
    int32_t class_id(int32_t a) const { while (parent_[a] != a) a = parent_[a]; return a; }
    void join(int32_t a, int32_t b) {
        a = class_id(a); b = class_id(b);
        if (a == b) return;
        if (a < b) parent_[b] = a; else parent_[a] = b;
    }
    void orbit_reps(std::vector<int32_t> &reps) const {
        reps.clear();
        for (int32_t i = 0; i < (int32_t)parent_.size(); i++) if (parent_[i] == i) reps.push_back(i);
    }
    void orbit(int32_t a, std::vector<int32_t> &o) const {
        o.clear();
        const int32_t r = class_id(a);
        for (int32_t i = 0; i < (int32_t)parent_.size(); i++) if (class_id(i) == r) o.push_back(i);
    }
    int32_t size() const { return (int32_t)parent_.size(); }
    // the library's partition labels are the smallest member of each class: adopt them as the parent array
    static EquivRel from_min_labels(const uint32_t *class_of, uint32_t n) {
        EquivRel e;
        e.parent_.assign(class_of, class_of + n);
        return e;
    }
  private:
    std::vector<int32_t> parent_;
};

struct Junction { uint32_t hcomp = 0; };                    // TigData1.jun (proc_lvar_auto.rs:996-1006)

struct TigData0 { int64_t donor_index = -1; };              // per cell, per chain; -1 = None (print_stats.rs:157-173)

struct TigData1 {                                           // per exact subclonotype, per chain
    bool left = false;                                      // heavy / TRB
    int64_t c_ref_id = -1, u_ref_id = -1;                   // Option<usize>
    size_t v_ref_id = 0, j_ref_id = 0;
    size_t cdr3_start = 0;
    int64_t fr1_start = 0, cdr1_start = -1, fr2_start = -1, cdr2_start = -1, fr3_start = -1;   // Option -> -1
    Junction jun;
};

struct ExactClonotype {
    std::vector<TigData1> share;
    std::vector<std::vector<TigData0>> clones;              // cell x chain
    size_t ncells() const { return clones.size(); }
};

// debruijn::dna_string::DnaString as the path uses it: 32 bases per u64, base i in word i >> 5 at bits
// 63-2(i&31), 62-2(i&31); A,C,G,T = 0,1,2,3
struct DnaString {
    std::vector<uint64_t> storage;
    size_t len = 0;
    static DnaString from_acgt_bytes(const std::string &s) {
        DnaString d;
        d.len = s.size();
        d.storage.assign((s.size() + 31) / 32, 0);
        for (size_t i = 0; i < s.size(); i++) {
            uint64_t code = 0;
            switch (s[i]) { case 'C': code = 1; break; case 'G': code = 2; break; case 'T': code = 3; break; default: code = 0; }
            d.storage[i >> 5] |= code << (62 - 2 * (i & 31));
        }
        return d;
    }
};

struct CloneInfo {                                          // one per (heavy, light) pair of an exact subclonotype
    std::vector<size_t> lens;                               // first sort key
    std::vector<std::string> tigs;                          // ASCII V..J, '-' = deletion
    std::vector<DnaString> tigsp;                           // the same, packed (what join_one's ndiffs reads when !has_del)
    std::vector<bool> has_del;
    size_t clonotype_id = 0;                                // index into exact_clonotypes
    std::vector<size_t> exact_cols;                         // columns of `share` used by this entry
    std::vector<std::string> vs, js;                        // germline actually compared (donor allele / indel-edited)
    std::vector<size_t> vsids, jsids;
    std::vector<int64_t> dref;                              // Option<usize> -> -1
    std::vector<std::string> cdr3s;
};

struct RefData { std::vector<std::string> refs, name; };   // refdata.refs / refdata.name
struct DonorReferenceItem { std::string nt_sequence; };

struct JoinAlgOpt {
    double max_score = 100000.0, mult_pow = 80.0, join_cdr3_ident = 85.0, fwr1_cdr12_delta = 20.0;
    int auto_share = 15, comp_filt = 8, comp_filt_bound = 80, max_cdr3_diffs = 1000, cdr3_mult = 5;
    bool easy = false, old_light = false, old_mult = false;
};
struct Heur { int max_diffs = 1000000, max_degradation = 2, ref_v_trim = 15, ref_j_trim = 15; };
struct EncloneControl {
    JoinAlgOpt join_alg_opt; Heur heur;
    int cdr3_normal_len = 42;                               // gen_opt.cdr3_normal_len
    bool mix_donors = false;                                // clono_filt_opt_def.donor
    int device = 0;
    int threads = 0;                                        // host threads for the flatten (0 = all)
    bool pinned = true;                                     // flat arenas in pinned memory (read in place by the GPU)
};

struct PotentialJoin {                                      // what join_info's log is formatted from
    int32_t k1, k2; int nrefs; int cd; int diffs; std::vector<int> shares, indeps;
    std::vector<std::vector<int>> shares_details; double p1, mult, score; bool err;
};

using Double = double;                                      // the `sr` argument is accepted and ignored (the library has its own tables)

struct JoinTimes { double ms_pack = 0, ms_create = 0, ms_run = 0, ms_unpack = 0; uint64_t h2d_bytes = 0, d2h_bytes = 0; int threads = 0; };

namespace detail {
inline double now_ms() { using namespace std::chrono; return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count(); }

template <typename F> void parallel_for(uint32_t n, int threads, F f) {      // f(thread, begin, end)
    threads = std::max(1, std::min(threads, (int)((n + 2047) / 2048)));
    if (threads == 1) { f(0, 0u, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < threads; t++) {
        const uint32_t a = (uint32_t)((uint64_t)n * t / threads), b = (uint32_t)((uint64_t)n * (t + 1) / threads);
        th.emplace_back([=] { f(t, a, b); });
    }
    for (auto &x : th) x.join();
}

// The flat compact pack and the memory it points into.
struct Flat {
    ejoin_pack_t pk;
    bool pinned = false;
    std::vector<uint32_t> exact_id, nchains, ncells, dlo, dhi, c3off, ref_len;
    std::vector<uint16_t> len[2], cdr3_len[2];
    std::vector<uint64_t> ref_off;
    std::string ref_arena;
    // Pinned buffers are expensive to create (cudaMallocHost: ~0.6 ms per MB) and cheap to keep: they go back to a
    // process-wide pool instead of the driver, so only the first join of a process pays for them.
    struct Pool { std::vector<std::pair<void *, size_t>> free_list; ~Pool() { for (auto &b : free_list) ejoin_host_free(b.first); } };
    static Pool &pool() { static Pool p; return p; }
    std::vector<std::pair<void *, size_t>> owned_sz;
    ~Flat() {
        for (auto &b : owned_sz) { if (pinned) pool().free_list.push_back(b); else free(b.first); }
    }
    void *alloc(size_t bytes) {
        void *p = nullptr;
        size_t got = bytes;
        if (pinned) {
            auto &fl = pool().free_list;
            size_t best = fl.size();
            for (size_t i = 0; i < fl.size(); i++) if (fl[i].second >= bytes && (best == fl.size() || fl[i].second < fl[best].second)) best = i;
            if (best < fl.size()) { p = fl[best].first; got = fl[best].second; fl.erase(fl.begin() + best); }
            else p = ejoin_host_alloc(bytes);
        }
        if (!p) {
            if (pinned && !owned_sz.empty()) throw std::runtime_error("pinned allocation failed");
            pinned = false;
            if (posix_memalign(&p, 4096, (bytes + 4095) & ~(size_t)4095) != 0) throw std::bad_alloc();
        }
        owned_sz.emplace_back(p, got);
        return p;
    }
};

#if defined(__BMI2__)
inline uint64_t planes_of_word(uint64_t w) { return (__builtin_ia32_pext_di(w, 0xAAAAAAAAAAAAAAAAull) << 32) | __builtin_ia32_pext_di(w, 0x5555555555555555ull); }
#else
inline uint64_t even_bits32(uint64_t v) {
    v &= 0x5555555555555555ull;
    v = (v | (v >> 1)) & 0x3333333333333333ull;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFull;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFull;
    return (v | (v >> 16)) & 0xFFFFFFFFull;
}
inline uint64_t planes_of_word(uint64_t w) { return (even_bits32(w >> 1) << 32) | even_bits32(w); }
#endif
inline uint8_t code_of(char b) { switch (b) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 0x80; } }

// Flat::pack of INTEGRATION.md §2: (exact_clonotypes, info, refdata) -> compact ejoin_pack_t, on `threads` threads.
inline void pack(Flat &F, bool is_bcr, const RefData &refdata, const std::vector<ExactClonotype> &exact_clonotypes,
                 const std::vector<CloneInfo> &info, int threads, bool want_pinned) {
    const uint32_t n = (uint32_t)info.size(), nx = (uint32_t)exact_clonotypes.size();
    F.pinned = want_pinned;
    memset(&F.pk, 0, sizeof(F.pk));
    // ---- pass 1 (parallel): sizes per entry -> per-thread totals -> offsets
    const int T = std::max(1, threads);
    F.exact_id.assign(n, 0); F.c3off.assign(n ? n : 1, 0);
    for (int c = 0; c < 2; c++) { F.len[c].assign(n, 0); F.cdr3_len[c].assign(n, 0); }
    std::vector<uint32_t> woff[2], aoff[2];
    std::vector<uint8_t> asc[2];
    for (int c = 0; c < 2; c++) { woff[c].assign(n, 0); aoff[c].assign(n, 0); asc[c].assign(n, 0); }
    struct Tot { uint64_t words = 0, abytes = 0, c3words = 0; };
    std::vector<Tot> tot(T + 1);
    auto chain_ascii = [](const CloneInfo &ci, size_t c) {
        if (ci.has_del[c]) return true;
        if (c < ci.tigsp.size() && ci.tigsp[c].len == ci.tigs[c].size()) return false;      // packed form at hand: ACGT only by construction
        for (char b : ci.tigs[c]) if (code_of(b) & 0x80) return true;
        return false;
    };
    auto sizes = [&](int t, uint32_t a, uint32_t b, bool write) {
        Tot run = write ? tot[t] : Tot();
        for (uint32_t k = a; k < b; k++) {
            const CloneInfo &ci = info[k];
            uint32_t c3n = 0;
            run.words = (run.words + 7) & ~7ull;                    // every entry starts on a 64-byte boundary (in-place reads move 64-byte pieces)
            for (size_t c = 0; c < ci.lens.size() && c < 2; c++) {
                const uint32_t L = (uint32_t)ci.tigs[c].size();
                const bool x = write ? asc[c][k] != 0 : chain_ascii(ci, c);
                if (!write) asc[c][k] = x;
                if (x) { if (write) aoff[c][k] = (uint32_t)run.abytes; run.abytes += L; }
                else { if (write) woff[c][k] = (uint32_t)run.words; run.words += (L + 31) / 32; }
                c3n += (uint32_t)ci.cdr3s[c].size();
            }
            if (write) F.c3off[k] = (uint32_t)run.c3words;
            run.c3words += (c3n + 31) / 32;
        }
        run.words = (run.words + 7) & ~7ull;                        // so that the next thread's range starts aligned too
        if (!write) tot[t] = run;
    };
    parallel_for(n, T, [&](int t, uint32_t a, uint32_t b) { sizes(t, a, b, false); });
    {   // exclusive prefix over the threads' totals
        Tot run;
        for (int t = 0; t <= T; t++) { Tot x = tot[t]; tot[t] = run; run.words += x.words; run.abytes += x.abytes; run.c3words += x.c3words; }
    }
    const Tot total = tot[T];
    if (total.words >= 0xFFFFFFFFull || total.abytes >= 0xFFFFFFFFull || total.c3words >= 0xFFFFFFFFull) throw std::runtime_error("join_exacts: input too large for 32-bit arena offsets");
    parallel_for(n, T, [&](int t, uint32_t a, uint32_t b) { sizes(t, a, b, true); });
    uint64_t *tig2 = (uint64_t *)F.alloc(8 * (total.words + 16));
    ejoin_entry_rec_t *rec = (ejoin_entry_rec_t *)F.alloc(sizeof(ejoin_entry_rec_t) * ((size_t)n + 1));
    uint8_t *ascii = (uint8_t *)F.alloc(total.abytes + 128);
    uint64_t *c3w = (uint64_t *)F.alloc(8 * (total.c3words + 16));
    memset(tig2 + total.words, 0, 128);       // (alignment gaps between entries are never read: every chain has its own offset) memset(ascii + total.abytes, 0, 128); memset(c3w + total.c3words, 0, 128);
    // ---- pass 2 (parallel): arenas and records; germline strings and V names interned per thread
    // Germline strings are interned by content (every CloneInfo owns its own copy of vs / js), but found through a cheap
    // key — length and 24 sampled bytes — and confirmed with one memcmp, instead of hashing ~1.4 KB of text per entry;
    // refdata.refs / refdata.name are interned by index.
    struct Local { std::vector<std::string_view> seqs; std::unordered_map<uint64_t, std::vector<uint32_t>> seq_index;
                   std::vector<std::string> names; std::unordered_map<std::string, uint32_t> name_index;
                   std::vector<int32_t> ref_cache, name_cache; bool bad_cdr3 = false, big_id = false; };
    std::vector<Local> loc(T);
    parallel_for(n, T, [&](int t, uint32_t a, uint32_t b) {
        Local &L = loc[t];
        L.ref_cache.assign(refdata.refs.size(), -1); L.name_cache.assign(refdata.name.size(), -1);
        auto intern = [&](const std::string &s) -> uint32_t {
            const std::string_view v(s);
            uint64_t key = v.size() * 0x9E3779B97F4A7C15ull;
            if (v.size() >= 8) {
                uint64_t x0, x1, x2;
                memcpy(&x0, v.data(), 8); memcpy(&x1, v.data() + v.size() / 2 - 4, 8); memcpy(&x2, v.data() + v.size() - 8, 8);
                key ^= x0 * 0xBF58476D1CE4E5B9ull ^ (x1 * 0x94D049BB133111EBull + (x2 << 7));
            }
            auto &cand = L.seq_index[key];
            for (uint32_t id : cand) if (L.seqs[id] == v) return id;
            const uint32_t id = (uint32_t)L.seqs.size();
            L.seqs.push_back(v); cand.push_back(id);
            return id;
        };
        auto intern_ref = [&](size_t ref_id) -> uint32_t {
            int32_t &slot = L.ref_cache.at(ref_id);
            if (slot < 0) slot = (int32_t)intern(refdata.refs[ref_id]);
            return (uint32_t)slot;
        };
        auto vname = [&](size_t v_ref_id) -> uint32_t {
            int32_t &slot = L.name_cache.at(v_ref_id);
            if (slot >= 0) return (uint32_t)slot;
            std::string s = refdata.name[v_ref_id];
            const auto p = s.find('*');
            if (p != std::string::npos) s.resize(p);
            auto it = L.name_index.find(s);
            uint32_t id;
            if (it != L.name_index.end()) id = it->second;
            else { id = (uint32_t)L.names.size(); L.names.push_back(s); L.name_index.emplace(s, id); }
            slot = (int32_t)id;
            return id;
        };
        auto id16 = [&](int64_t v) -> uint16_t { if (v < 0) return 0xFFFF; if (v >= 0xFFFF) L.big_id = true; return (uint16_t)v; };
        auto opt16 = [](int64_t v) { return v < 0 ? (uint16_t)EJOIN_NONE16 : (uint16_t)v; };
        for (uint32_t k = a; k < b; k++) {
            const CloneInfo &ci = info[k];
            const ExactClonotype &ex = exact_clonotypes.at(ci.clonotype_id);
            F.exact_id[k] = (uint32_t)ci.clonotype_id;
            ejoin_entry_rec_t r;
            memset(&r, 0, sizeof(r));
            r.c_ref_id_light = 0xFFFF;
            r.fr1_start = r.cdr1_start = r.fr2_start = r.cdr2_start = r.fr3_start = EJOIN_NONE16;
            for (int c = 0; c < 2; c++) { r.v_ref_id[c] = r.j_ref_id[c] = r.dref_id[c] = r.vname_id[c] = r.utr_seq[c] = 0xFFFF; }
            uint64_t *cw = c3w + F.c3off[k];
            uint32_t cpos = 0;
            uint64_t cx = 0;
            for (size_t c = 0; c < ci.lens.size() && c < 2; c++) {
                const TigData1 &td = ex.share.at(ci.exact_cols.at(c));
                const std::string &tig = ci.tigs[c];
                const uint32_t Ln = (uint32_t)tig.size();
                F.len[c][k] = (uint16_t)Ln; F.cdr3_len[c][k] = (uint16_t)ci.cdr3s[c].size();
                if (asc[c][k]) { memcpy(ascii + aoff[c][k], tig.data(), Ln); r.tig_off[c] = aoff[c][k]; r.has_del[c] = 1; }
                else {
                    uint64_t *w = tig2 + woff[c][k];
                    // the reference's own packed form, each word split into its two bit planes (EJOIN_F_TIGS_PLANES: two `pext`s per
                    // word where BMI2 is there; the GPU would need a 45-instruction transpose for the same)
                    const size_t nwd = (Ln + 31) / 32;
                    if (c < ci.tigsp.size() && ci.tigsp[c].len == Ln) for (size_t q = 0; q < nwd; q++) w[q] = planes_of_word(ci.tigsp[c].storage[q]);
                    else {
                        const DnaString d = DnaString::from_acgt_bytes(tig);
                        for (size_t q = 0; q < nwd; q++) w[q] = planes_of_word(d.storage[q]);
                    }
                    r.tig_off[c] = woff[c][k]; r.has_del[c] = 0;
                }
                for (char ch : ci.cdr3s[c]) {                       // both CDR3s concatenated, DnaString bit order
                    const uint8_t code = code_of(ch);
                    if (code & 0x80) L.bad_cdr3 = true;
                    cx |= (uint64_t)(code & 3) << (62 - 2 * (cpos & 31));
                    if ((++cpos & 31) == 0) { cw[(cpos >> 5) - 1] = cx; cx = 0; }
                }
                r.cdr3_start[c] = (uint16_t)td.cdr3_start;
                r.v_ref_id[c] = id16((int64_t)ci.vsids[c]); r.j_ref_id[c] = id16((int64_t)ci.jsids[c]); r.dref_id[c] = id16(ci.dref[c]);
                r.vs_seq[c] = (uint16_t)intern(ci.vs[c]); r.js_seq[c] = (uint16_t)intern(ci.js[c]);       // thread-local ids, remapped below
                r.vname_id[c] = (uint16_t)vname(ci.vsids[c]); r.vuni_seq[c] = (uint16_t)intern_ref(ci.vsids[c]);
                r.utr_seq[c] = td.u_ref_id < 0 ? (uint16_t)0xFFFF : (uint16_t)intern_ref((size_t)td.u_ref_id);
                if (c == 0) {
                    r.hcomp = (uint16_t)td.jun.hcomp; r.fr1_start = opt16(td.fr1_start); r.cdr1_start = opt16(td.cdr1_start);
                    r.fr2_start = opt16(td.fr2_start); r.cdr2_start = opt16(td.cdr2_start); r.fr3_start = opt16(td.fr3_start);
                } else r.c_ref_id_light = id16(td.c_ref_id);
            }
            if (cpos & 31) cw[cpos >> 5] = cx;
            r.pad[0] = (uint8_t)t;                                  // which thread's id space the germline / name ids are in
            rec[k] = r;
        }
    });
    for (auto &L : loc) {
        if (L.bad_cdr3) throw std::runtime_error("join_exacts: a CDR3 contains a character outside ACGT");
        if (L.big_id || L.seqs.size() >= 0xFFFF) throw std::runtime_error("join_exacts: an id does not fit the 16-bit entry record");
    }
    // merge the threads' germline / name tables, remap the records (parallel again)
    std::vector<std::string_view> refs;
    std::unordered_map<std::string_view, uint32_t> ref_index;
    std::unordered_map<std::string, uint32_t> name_index;
    std::vector<std::vector<uint16_t>> seq_map(T), name_map(T);
    for (int t = 0; t < T; t++) {
        for (auto v : loc[t].seqs) {
            auto it = ref_index.find(v);
            uint32_t id;
            if (it != ref_index.end()) id = it->second; else { id = (uint32_t)refs.size(); refs.push_back(v); ref_index.emplace(v, id); }
            seq_map[t].push_back((uint16_t)id);
        }
        for (auto &s : loc[t].names) {
            auto it = name_index.find(s);
            uint32_t id;
            if (it != name_index.end()) id = it->second; else { id = (uint32_t)name_index.size(); name_index.emplace(s, id); }
            name_map[t].push_back((uint16_t)id);
        }
    }
    if (refs.size() >= 0xFFFF) throw std::runtime_error("join_exacts: more than 65534 distinct germline sequences");
    parallel_for(n, T, [&](int, uint32_t a, uint32_t b) {
        for (uint32_t k = a; k < b; k++) {
            ejoin_entry_rec_t &r = rec[k];
            const auto &sm = seq_map[r.pad[0]];
            const auto &nm = name_map[r.pad[0]];
            for (int c = 0; c < 2; c++) {
                if (F.len[c][k] == 0 && c == 1) continue;
                r.vs_seq[c] = sm[r.vs_seq[c]]; r.js_seq[c] = sm[r.js_seq[c]]; r.vuni_seq[c] = sm[r.vuni_seq[c]];
                if (r.utr_seq[c] != 0xFFFF) r.utr_seq[c] = sm[r.utr_seq[c]];
                if (r.vname_id[c] != 0xFFFF) r.vname_id[c] = nm[r.vname_id[c]];
            }
            r.pad[0] = 0;
        }
    });
    F.nchains.assign(nx, 0); F.ncells.assign(nx, 0); F.dlo.assign(nx, EJOIN_NONE); F.dhi.assign(nx, EJOIN_NONE);
    parallel_for(nx, T, [&](int, uint32_t a, uint32_t b) {
        for (uint32_t x = a; x < b; x++) {
            const ExactClonotype &ex = exact_clonotypes[x];
            F.nchains[x] = (uint32_t)ex.share.size(); F.ncells[x] = (uint32_t)ex.ncells();
            for (const auto &cell : ex.clones)
                if (!cell.empty() && cell[0].donor_index >= 0) {
                    const uint32_t d = (uint32_t)cell[0].donor_index;
                    F.dlo[x] = F.dlo[x] == EJOIN_NONE ? d : std::min(F.dlo[x], d); F.dhi[x] = F.dhi[x] == EJOIN_NONE ? d : std::max(F.dhi[x], d);
                }
        }
    });
    F.ref_off.resize(refs.size()); F.ref_len.resize(refs.size());
    for (size_t r = 0; r < refs.size(); r++) { F.ref_off[r] = F.ref_arena.size(); F.ref_len[r] = (uint32_t)refs[r].size(); F.ref_arena.append(refs[r]); }
    F.ref_arena.append(64, '\0');
    ejoin_pack_t &pk = F.pk;
    pk.abi_version = EJOIN_ABI_VERSION; pk.n_info = n; pk.n_exact = nx; pk.n_refseq = (uint32_t)refs.size(); pk.is_bcr = is_bcr;
    pk.flags = EJOIN_F_TIGS_2BIT | EJOIN_F_TIGS_PLANES | EJOIN_F_ENTRY_RECS | EJOIN_F_CDR3_2BIT | (F.pinned ? EJOIN_F_TIGS_IN_PLACE : 0u);
    pk.exact_id = F.exact_id.data();
    for (int c = 0; c < 2; c++) { pk.len[c] = F.len[c].data(); pk.cdr3_len[c] = F.cdr3_len[c].data(); }
    pk.tig_arena = ascii; pk.tig_arena_bytes = total.abytes;
    pk.tig2_arena = tig2; pk.tig2_arena_words = total.words; pk.entry_rec = rec;
    pk.cdr3_2bit = c3w; pk.cdr3_2bit_words = total.c3words; pk.cdr3_2bit_off = F.c3off.data();
    pk.nchains = F.nchains.data(); pk.ncells = F.ncells.data(); pk.donor_lo = F.dlo.data(); pk.donor_hi = F.dhi.data();
    pk.refseq_off = F.ref_off.data(); pk.refseq_len = F.ref_len.data(); pk.refseq_arena = (const uint8_t *)F.ref_arena.data();
    pk.refseq_arena_bytes = F.ref_arena.size() - 64;
}
}  // namespace detail

inline ejoin_params_t params_of(const EncloneControl &ctl) {
    ejoin_params_t p; ejoin_params_default(&p);
    p.max_score = ctl.join_alg_opt.max_score; p.mult_pow = ctl.join_alg_opt.mult_pow; p.join_cdr3_ident = ctl.join_alg_opt.join_cdr3_ident;
    p.fwr1_cdr12_delta = ctl.join_alg_opt.fwr1_cdr12_delta; p.cdr3_normal_len = ctl.cdr3_normal_len; p.auto_share = ctl.join_alg_opt.auto_share;
    p.comp_filt = ctl.join_alg_opt.comp_filt; p.comp_filt_bound = ctl.join_alg_opt.comp_filt_bound; p.max_cdr3_diffs = ctl.join_alg_opt.max_cdr3_diffs;
    p.cdr3_mult = ctl.join_alg_opt.cdr3_mult; p.max_degradation = ctl.heur.max_degradation; p.max_diffs = ctl.heur.max_diffs;
    p.ref_v_trim = ctl.heur.ref_v_trim; p.ref_j_trim = ctl.heur.ref_j_trim; p.easy = ctl.join_alg_opt.easy; p.old_light = ctl.join_alg_opt.old_light;
    p.old_mult = ctl.join_alg_opt.old_mult; p.mix_donors = ctl.mix_donors;
    return p;
}

// `handle` (optional): a handle made earlier with ejoin_create(params_of(ctl)) — e.g. at program start, so that the CUDA
// context and the library's tables are not on the join's critical path; nullptr creates and destroys one here.
inline EquivRel join_exacts(bool is_bcr, const std::map<std::pair<size_t, size_t>, std::vector<std::string>> & /*to_bc*/,
                            const RefData &refdata, const EncloneControl &ctl, const std::vector<ExactClonotype> &exact_clonotypes,
                            const std::vector<CloneInfo> &info, std::vector<std::pair<int32_t, int32_t>> &raw_joins,
                            const std::vector<std::vector<Double>> & /*sr*/, const std::vector<DonorReferenceItem> & /*dref*/,
                            std::vector<PotentialJoin> *join_info = nullptr, JoinTimes *times = nullptr, ejoin_handle_t *handle = nullptr) {
    const uint32_t n = (uint32_t)info.size();
    const int threads = ctl.threads > 0 ? ctl.threads : std::max(1u, std::thread::hardware_concurrency());
    double t0 = detail::now_ms();
    ejoin_handle_t *h = handle;
    if (!h) {
        const ejoin_params_t p = params_of(ctl);
        int rc = ejoin_create(&h, &p, ctl.device);
        if (rc != 0) { std::string m = h ? ejoin_last_error(h) : "no usable CUDA device"; if (h) ejoin_destroy(h); throw std::runtime_error("ejoin_create: " + m); }
    }
    double t1 = detail::now_ms();
    detail::Flat F;
    try { detail::pack(F, is_bcr, refdata, exact_clonotypes, info, threads, ctl.pinned); }
    catch (...) { if (!handle) ejoin_destroy(h); throw; }
    double t2 = detail::now_ms();
    ejoin_result_t res;
    int rc = ejoin_run(h, &F.pk, &res);
    if (rc != 0) { std::string m = ejoin_last_error(h); if (!handle) ejoin_destroy(h); throw std::runtime_error("ejoin_run: " + m); }
    double t3 = detail::now_ms();
    raw_joins.resize(res.n_edges);
    if (join_info) join_info->clear();
    for (uint64_t i = 0; i < res.n_edges; i++) {
        const ejoin_edge_t &e = res.edges[i];
        raw_joins[i] = { (int32_t)e.k1, (int32_t)e.k2 };
        if (join_info) {
            PotentialJoin pj{ (int32_t)e.k1, (int32_t)e.k2, e.nrefs, e.cd, (int)e.diffs, {}, {}, {}, e.p1, e.mult, e.score, e.err != 0 };
            for (int u = 0; u < e.nrefs; u++) {
                pj.shares.push_back(e.shares[u]); pj.indeps.push_back(e.indeps[u]);
                pj.shares_details.push_back({ e.shares_details[u][0], e.shares_details[u][1], e.shares_details[u][2], e.shares_details[u][3] });
            }
            join_info->push_back(std::move(pj));
        }
    }
    EquivRel eq = n ? EquivRel::from_min_labels(res.class_of, n) : EquivRel(0);
    if (times) {
        times->ms_create = t1 - t0; times->ms_pack = t2 - t1; times->ms_run = t3 - t2; times->ms_unpack = detail::now_ms() - t3;
        times->h2d_bytes = res.stats.h2d_bytes; times->d2h_bytes = res.stats.d2h_bytes; times->threads = threads;
    }
    ejoin_result_free(&res);
    if (!handle) ejoin_destroy(h);
    return eq;
}

}  // namespace enclone
