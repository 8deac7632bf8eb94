// join_exacts.hpp — C++ host-side mirror of the reference interface for the join path.
//
// The reference is Rust (enclone_ranger@cbd15ef, un-vendored; no Rust toolchain in the build image),
// so the host side above the C ABI is written in C++ with the reference's own names and argument
// meaning:
//
//   join_exacts(is_bcr, to_bc, refdata, ctl, exact_clonotypes, info, raw_joins, sr, dref) -> EquivRel
//
// It flattens (exact_clonotypes, info, refdata, dref) into an ejoin_pack_t — exactly what the Rust
// `Flat::pack` of INTEGRATION.md does — calls libenclone_join.so, and rebuilds the outputs the
// callers consume: raw_joins (enclone_process::process_clonotypes, enclone_main/src/stop.rs:56), the
// EquivRel (orbits), and the PotentialJoin numbers behind join_info (enclone_tail/src/group.rs:70-74).
// Errors follow the reference's convention for this call (it cannot fail short of a panic):
// a library error throws std::runtime_error carrying ejoin_last_error().
#pragma once
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
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

struct CloneInfo {                                          // one per (heavy, light) pair of an exact subclonotype
    std::vector<size_t> lens;                               // first sort key
    std::vector<std::string> tigs;                          // ASCII V..J, '-' = deletion
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
};

struct PotentialJoin {                                      // what join_info's log is formatted from
    int32_t k1, k2; int nrefs; int cd; int diffs; std::vector<int> shares, indeps;
    std::vector<std::vector<int>> shares_details; double p1, mult, score; bool err;
};

using Double = double;                                      // the `sr` argument is accepted and ignored (the library has its own tables)

inline EquivRel join_exacts(bool is_bcr, const std::map<std::pair<size_t, size_t>, std::vector<std::string>> & /*to_bc*/,
                            const RefData &refdata, const EncloneControl &ctl, const std::vector<ExactClonotype> &exact_clonotypes,
                            const std::vector<CloneInfo> &info, std::vector<std::pair<int32_t, int32_t>> &raw_joins,
                            const std::vector<std::vector<Double>> & /*sr*/, const std::vector<DonorReferenceItem> & /*dref*/,
                            std::vector<PotentialJoin> *join_info = nullptr) {
    const uint32_t n = (uint32_t)info.size(), nx = (uint32_t)exact_clonotypes.size();
    // ---- germline table, deduplicated by content
    std::vector<std::string> refs;
    std::unordered_map<std::string, uint32_t> ref_index;
    auto intern = [&](const std::string &s) { auto it = ref_index.find(s); if (it != ref_index.end()) return it->second;
                                              uint32_t id = (uint32_t)refs.size(); refs.push_back(s); ref_index.emplace(s, id); return id; };
    std::unordered_map<std::string, uint32_t> name_index;
    auto vname = [&](size_t v_ref_id) { std::string s = refdata.name.at(v_ref_id); auto p = s.find('*'); if (p != std::string::npos) s.resize(p);
                                        auto it = name_index.find(s); if (it != name_index.end()) return it->second;
                                        uint32_t id = (uint32_t)name_index.size(); name_index.emplace(s, id); return id; };
    std::vector<uint32_t> exact_id(n), c_light(n, EJOIN_NONE);
    std::vector<uint16_t> hcomp(n, 0), fr1(n, EJOIN_NONE16), cdr1(n, EJOIN_NONE16), fr2(n, EJOIN_NONE16), cdr2(n, EJOIN_NONE16), fr3(n, EJOIN_NONE16);
    std::vector<uint16_t> len[2], cdr3_len[2], cdr3_start[2];
    std::vector<uint64_t> tig_off[2], cdr3_off[2];
    std::vector<uint8_t> has_del[2];
    std::vector<uint32_t> v_ref[2], j_ref[2], dref_id[2], vs_seq[2], js_seq[2], vname_id[2], vuni[2], utr[2];
    for (int c = 0; c < 2; c++) {
        len[c].assign(n, 0); cdr3_len[c].assign(n, 0); cdr3_start[c].assign(n, 0); tig_off[c].assign(n, 0); cdr3_off[c].assign(n, 0);
        has_del[c].assign(n, 0); v_ref[c].assign(n, EJOIN_NONE); j_ref[c].assign(n, EJOIN_NONE); dref_id[c].assign(n, EJOIN_NONE);
        vs_seq[c].assign(n, 0); js_seq[c].assign(n, 0); vname_id[c].assign(n, EJOIN_NONE); vuni[c].assign(n, 0); utr[c].assign(n, EJOIN_NONE);
    }
    std::string tig_arena, cdr3_arena;
    auto opt16 = [](int64_t v) { return v < 0 ? (uint16_t)EJOIN_NONE16 : (uint16_t)v; };
    for (uint32_t k = 0; k < n; k++) {
        const CloneInfo &ci = info[k];
        const ExactClonotype &ex = exact_clonotypes.at(ci.clonotype_id);
        exact_id[k] = (uint32_t)ci.clonotype_id;
        for (size_t c = 0; c < ci.lens.size() && c < 2; c++) {
            const TigData1 &td = ex.share.at(ci.exact_cols.at(c));
            len[c][k] = (uint16_t)ci.tigs[c].size(); tig_off[c][k] = tig_arena.size(); tig_arena += ci.tigs[c];
            has_del[c][k] = ci.has_del[c]; cdr3_len[c][k] = (uint16_t)ci.cdr3s[c].size(); cdr3_off[c][k] = cdr3_arena.size();
            cdr3_arena += ci.cdr3s[c]; cdr3_start[c][k] = (uint16_t)td.cdr3_start;
            v_ref[c][k] = (uint32_t)ci.vsids[c]; j_ref[c][k] = (uint32_t)ci.jsids[c];
            dref_id[c][k] = ci.dref[c] < 0 ? EJOIN_NONE : (uint32_t)ci.dref[c];
            vs_seq[c][k] = intern(ci.vs[c]); js_seq[c][k] = intern(ci.js[c]);
            vname_id[c][k] = vname(ci.vsids[c]); vuni[c][k] = intern(refdata.refs.at(ci.vsids[c]));
            utr[c][k] = td.u_ref_id < 0 ? EJOIN_NONE : intern(refdata.refs.at((size_t)td.u_ref_id));
            if (c == 0) {
                hcomp[k] = (uint16_t)td.jun.hcomp; fr1[k] = opt16(td.fr1_start); cdr1[k] = opt16(td.cdr1_start);
                fr2[k] = opt16(td.fr2_start); cdr2[k] = opt16(td.cdr2_start); fr3[k] = opt16(td.fr3_start);
            } else c_light[k] = td.c_ref_id < 0 ? EJOIN_NONE : (uint32_t)td.c_ref_id;
        }
    }
    std::vector<uint32_t> nchains(nx), ncells(nx), dlo(nx, EJOIN_NONE), dhi(nx, EJOIN_NONE);
    for (uint32_t x = 0; x < nx; x++) {
        nchains[x] = (uint32_t)exact_clonotypes[x].share.size(); ncells[x] = (uint32_t)exact_clonotypes[x].ncells();
        for (const auto &cell : exact_clonotypes[x].clones)
            if (!cell.empty() && cell[0].donor_index >= 0) {
                uint32_t d = (uint32_t)cell[0].donor_index;
                dlo[x] = dlo[x] == EJOIN_NONE ? d : std::min(dlo[x], d); dhi[x] = dhi[x] == EJOIN_NONE ? d : std::max(dhi[x], d);
            }
    }
    std::vector<uint64_t> ref_off(refs.size()); std::vector<uint32_t> ref_len(refs.size()); std::string ref_arena;
    for (size_t r = 0; r < refs.size(); r++) { ref_off[r] = ref_arena.size(); ref_len[r] = (uint32_t)refs[r].size(); ref_arena += refs[r]; }
    tig_arena.append(64, '\0'); cdr3_arena.append(64, '\0'); ref_arena.append(64, '\0');
    ejoin_pack_t pk; memset(&pk, 0, sizeof(pk));
    pk.abi_version = EJOIN_ABI_VERSION; pk.n_info = n; pk.n_exact = nx; pk.n_refseq = (uint32_t)refs.size(); pk.is_bcr = is_bcr;
    pk.exact_id = exact_id.data();
    for (int c = 0; c < 2; c++) {
        pk.len[c] = len[c].data(); pk.tig_off[c] = tig_off[c].data(); pk.has_del[c] = has_del[c].data(); pk.cdr3_len[c] = cdr3_len[c].data();
        pk.cdr3_off[c] = cdr3_off[c].data(); pk.cdr3_start[c] = cdr3_start[c].data(); pk.v_ref_id[c] = v_ref[c].data(); pk.j_ref_id[c] = j_ref[c].data();
        pk.dref_id[c] = dref_id[c].data(); pk.vs_seq[c] = vs_seq[c].data(); pk.js_seq[c] = js_seq[c].data(); pk.vname_id[c] = vname_id[c].data();
        pk.vuni_seq[c] = vuni[c].data(); pk.utr_seq[c] = utr[c].data();
    }
    pk.tig_arena = (const uint8_t *)tig_arena.data(); pk.tig_arena_bytes = tig_arena.size() - 64;
    pk.cdr3_arena = (const uint8_t *)cdr3_arena.data(); pk.cdr3_arena_bytes = cdr3_arena.size() - 64;
    pk.c_ref_id_light = c_light.data(); pk.hcomp = hcomp.data();
    pk.fr1_start = fr1.data(); pk.cdr1_start = cdr1.data(); pk.fr2_start = fr2.data(); pk.cdr2_start = cdr2.data(); pk.fr3_start = fr3.data();
    pk.nchains = nchains.data(); pk.ncells = ncells.data(); pk.donor_lo = dlo.data(); pk.donor_hi = dhi.data();
    pk.refseq_off = ref_off.data(); pk.refseq_len = ref_len.data(); pk.refseq_arena = (const uint8_t *)ref_arena.data();
    pk.refseq_arena_bytes = ref_arena.size() - 64;
    ejoin_params_t p; ejoin_params_default(&p);
    p.max_score = ctl.join_alg_opt.max_score; p.mult_pow = ctl.join_alg_opt.mult_pow; p.join_cdr3_ident = ctl.join_alg_opt.join_cdr3_ident;
    p.fwr1_cdr12_delta = ctl.join_alg_opt.fwr1_cdr12_delta; p.cdr3_normal_len = ctl.cdr3_normal_len; p.auto_share = ctl.join_alg_opt.auto_share;
    p.comp_filt = ctl.join_alg_opt.comp_filt; p.comp_filt_bound = ctl.join_alg_opt.comp_filt_bound; p.max_cdr3_diffs = ctl.join_alg_opt.max_cdr3_diffs;
    p.cdr3_mult = ctl.join_alg_opt.cdr3_mult; p.max_degradation = ctl.heur.max_degradation; p.max_diffs = ctl.heur.max_diffs;
    p.ref_v_trim = ctl.heur.ref_v_trim; p.ref_j_trim = ctl.heur.ref_j_trim; p.easy = ctl.join_alg_opt.easy; p.old_light = ctl.join_alg_opt.old_light;
    p.old_mult = ctl.join_alg_opt.old_mult; p.mix_donors = ctl.mix_donors;
    ejoin_handle_t *h = nullptr;
    int rc = ejoin_create(&h, &p, ctl.device);
    if (rc != 0) { std::string m = h ? ejoin_last_error(h) : "no usable CUDA device"; if (h) ejoin_destroy(h); throw std::runtime_error("ejoin_create: " + m); }
    ejoin_result_t res;
    rc = ejoin_run(h, &pk, &res);
    if (rc != 0) { std::string m = ejoin_last_error(h); ejoin_destroy(h); throw std::runtime_error("ejoin_run: " + m); }
    raw_joins.clear();
    if (join_info) join_info->clear();
    for (uint64_t i = 0; i < res.n_edges; i++) {
        const ejoin_edge_t &e = res.edges[i];
        raw_joins.emplace_back((int32_t)e.k1, (int32_t)e.k2);
        if (join_info) {
            PotentialJoin pj{ (int32_t)e.k1, (int32_t)e.k2, e.nrefs, e.cd, (int)e.diffs, {}, {}, {}, e.p1, e.mult, e.score, e.err != 0 };
            for (int u = 0; u < e.nrefs; u++) {
                pj.shares.push_back(e.shares[u]); pj.indeps.push_back(e.indeps[u]);
                pj.shares_details.push_back({ e.shares_details[u][0], e.shares_details[u][1], e.shares_details[u][2], e.shares_details[u][3] });
            }
            join_info->push_back(std::move(pj));
        }
    }
    EquivRel eq((int32_t)n);
    for (uint32_t k = 0; k < n; k++) if (res.class_of[k] != k) eq.join((int32_t)res.class_of[k], (int32_t)k);
    ejoin_result_free(&res);
    ejoin_destroy(h);
    return eq;
}

}  // namespace enclone
