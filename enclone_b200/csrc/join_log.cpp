// join_log.cpp — the text block enclone prints for one join (JoinInfo.log, enclone_tail/src/group.rs:70-74,434-441;
// golden: enclone_exec/testx/inputs/outputs/enclone_test256_output:29-70), rebuilt from one ejoin_edge_t plus the
// caller's inputs.  Host code only.  The upstream formatter lives in the un-vendored enclone crate; the layout, the
// labels and the number formats below are read off the golden (Rust `{}` for f64 = shortest round-trip digits in
// positional notation; `{:.1}` for the one-decimal fields).  What the golden does not show (how light-chain or
// J share positions, or a second reference choice, are laid out) is marked [GUESS].
#include <algorithm>
#include <charconv>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/enclone_join.h"

namespace {

std::string f64_display(double v) {              // Rust `{}`
    char b[512];
    auto r = std::to_chars(b, b + sizeof(b), v, std::chars_format::fixed);
    return std::string(b, r.ptr);
}
std::string f64_fixed1(double v) { char b[64]; snprintf(b, sizeof(b), "%.1f", v); return b; }

struct Entry {
    std::string tig[2];
    const uint8_t *v[2], *j[2];
    int vl[2], jl[2], cs[2], cl[2];
    int fr1, cdr1, fr2, cdr2, fr3;
};

// one chain's bases as ASCII, whatever form the pack holds them in
std::string chain_bytes(const ejoin_pack_t *pk, int c, uint32_t k) {
    const bool recs = pk->flags & EJOIN_F_ENTRY_RECS, two_bit = pk->flags & EJOIN_F_TIGS_2BIT;
    const uint32_t n = pk->len[c][k];
    const uint64_t off = recs ? (uint64_t)pk->entry_rec[k].tig_off[c] : pk->tig_off[c][k];
    const bool asc = !two_bit || (recs ? pk->entry_rec[k].has_del[c] != 0 : pk->has_del[c][k] != 0);
    std::string s(n, 'A');
    if (asc) memcpy(&s[0], pk->tig_arena + off, n);
    else if (pk->flags & EJOIN_F_TIGS_PLANES)
        for (uint32_t i = 0; i < n; i++) { const uint64_t w = pk->tig2_arena[off + (i >> 5)]; s[i] = "ACGT"[(((w >> (63 - (i & 31))) & 1) << 1) | ((w >> (31 - (i & 31))) & 1)]; }
    else for (uint32_t i = 0; i < n; i++) s[i] = "ACGT"[(pk->tig2_arena[off + (i >> 5)] >> (62 - 2 * (i & 31))) & 3];
    return s;
}
Entry load(const ejoin_pack_t *pk, uint32_t k) {
    Entry e;
    const bool recs = pk->flags & EJOIN_F_ENTRY_RECS;
    for (int c = 0; c < 2; c++) {
        e.tig[c] = chain_bytes(pk, c, k);
        const uint32_t vi = recs ? pk->entry_rec[k].vs_seq[c] : pk->vs_seq[c][k], ji = recs ? pk->entry_rec[k].js_seq[c] : pk->js_seq[c][k];
        e.v[c] = pk->refseq_arena + pk->refseq_off[vi]; e.vl[c] = (int)pk->refseq_len[vi];
        e.j[c] = pk->refseq_arena + pk->refseq_off[ji]; e.jl[c] = (int)pk->refseq_len[ji];
        e.cs[c] = recs ? pk->entry_rec[k].cdr3_start[c] : pk->cdr3_start[c][k];
        e.cl[c] = pk->cdr3_len[c][k];
    }
    auto o16 = [](uint16_t v) { return v == EJOIN_NONE16 ? -1 : (int)v; };
    if (recs) { const ejoin_entry_rec_t &r = pk->entry_rec[k]; e.fr1 = o16(r.fr1_start); e.cdr1 = o16(r.cdr1_start); e.fr2 = o16(r.fr2_start); e.cdr2 = o16(r.cdr2_start); e.fr3 = o16(r.fr3_start); }
    else { e.fr1 = o16(pk->fr1_start[k]); e.cdr1 = o16(pk->cdr1_start[k]); e.fr2 = o16(pk->fr2_start[k]); e.cdr2 = o16(pk->cdr2_start[k]); e.fr3 = o16(pk->fr3_start[k]); }
    return e;
}

}  // namespace

extern "C" int ejoin_format_join_log(const ejoin_pack_t *pk, const ejoin_params_t *pr, const ejoin_edge_t *e, const ejoin_log_names_t *nm, char *buf,
                                     size_t cap, size_t *len_out) {
    if (!pk || !pr || !e || !nm || !len_out || e->k1 >= pk->n_info || e->k2 >= pk->n_info) return EJOIN_E_ARG;
    const Entry a = load(pk, e->k1), b = load(pk, e->k2);
    std::string s;
    auto line = [&](const std::string &x) { s += x; s += '\n'; };
    auto str = [](const char *p) { return std::string(p ? p : ""); };
    if (e->err) { line("JOIN ERROR"); line(""); }
    line("[" + std::to_string(nm->index) + "] = " + str(nm->id1) + " * " + str(nm->id2));
    line("ref v segs1 = " + str(nm->vsegs1));
    line("ref v segs2 = " + str(nm->vsegs2));
    line(str(nm->dataset1));
    line(str(nm->dataset2));
    line(str(nm->cdr3_aa1) + ", mult = " + std::to_string(pk->ncells[pk->exact_id[e->k1]]));
    line("bcs = " + str(nm->bcs1));
    line(str(nm->cdr3_aa2) + ", mult = " + std::to_string(pk->ncells[pk->exact_id[e->k2]]));
    line("bcs = " + str(nm->bcs2));
    line("score = " + f64_fixed1(e->score));
    // share positions per reference choice (share_pos_v / share_pos_j of PotentialJoin): recomputed here, on the host, for
    // the one edge being printed; the windows are those of join_one (print_utils5.rs:78-81)
    std::vector<std::vector<int>> is_share(2);           // per chain: 1 at a shared mutation (for the difference pattern; reference choice 0)
    for (int u = 0; u < e->nrefs; u++) {
        const Entry &r = u == 0 ? a : b;
        std::vector<int> pos_v, pos_j;
        for (int c = 0; c < 2; c++) {
            const int n = (int)a.tig[c].size();
            if (u == 0) is_share[c].assign(n, 0);
            const int vend = std::min(n, r.vl[c] - pr->ref_v_trim);
            for (int p = 0; p < vend; p++)
                if (a.tig[c][p] == b.tig[c][p] && (uint8_t)a.tig[c][p] != r.v[c][p]) { pos_v.push_back(p); if (u == 0) is_share[c][p] = 1; }
            const int jw = r.jl[c] - pr->ref_j_trim;
            if (jw > 0)
                for (int p = std::max(0, n - jw); p < n; p++)
                    if (a.tig[c][p] == b.tig[c][p] && (uint8_t)a.tig[c][p] != r.j[c][r.jl[c] - (n - p)]) { pos_j.push_back(p); if (u == 0) is_share[c][p] = 1; }
        }
        const uint16_t *d = e->shares_details[u];
        line("versus ref " + std::to_string(u + 1) + ", shares = " + std::to_string(e->shares[u]) + ", indeps = " + std::to_string(e->indeps[u]) +
             " (shares breakdown = V" + std::to_string(d[0]) + ", J" + std::to_string(d[1]) + "; V" + std::to_string(d[2]) + ", J" + std::to_string(d[3]) + ")");
        auto list = [](const std::vector<int> &v) { std::string o; for (size_t i = 0; i < v.size(); i++) { if (i) o += ","; o += std::to_string(v[i]); } return o; };
        if (!pos_v.empty()) line("              share pos v = " + list(pos_v));      // [GUESS] both chains in one list, heavy first
        if (!pos_j.empty()) line("              share pos j = " + list(pos_j));      // [GUESS] the golden has no J share
    }
    line("CDR3 nucleotide diffs = " + std::to_string(e->cd));
    line("in amino acid space:");
    line(str(nm->cdr3_aa1));
    line(str(nm->cdr3_aa2));
    // heavy-chain region differences and identities (k1's bounds, as gate m of join_one)
    if (a.fr1 >= 0 && a.cdr1 >= 0 && a.fr2 >= 0 && a.cdr2 >= 0 && a.fr3 >= 0) {
        auto diffs = [&](int lo, int hi) { int d = 0; for (int p = std::max(0, lo); p < hi && p < (int)a.tig[0].size(); p++) d += a.tig[0][p] != b.tig[0][p]; return d; };
        const int dfw = diffs(a.fr1, a.cdr1), dc1 = diffs(a.cdr1, a.fr2), dc2 = diffs(a.cdr2, a.fr3), dc3 = diffs(a.cs[0], a.cs[0] + a.cl[0]);
        const int lfw = a.cdr1 - a.fr1, l1 = a.fr2 - a.cdr1, l2 = a.fr3 - a.cdr2, l3 = a.cl[0];
        auto ident = [](int d, int l) { return 100.0 * (1.0 - (double)d / (double)l); };
        line("heavy chain FWR1 diffs = " + std::to_string(dfw));
        line("heavy chain CDR1 diffs = " + std::to_string(dc1));
        line("heavy chain CDR2 diffs = " + std::to_string(dc2));
        line("heavy chain CDR3 diffs = " + std::to_string(dc3));
        line("nucleotide identity on heavy chain FWR1 = " + f64_fixed1(ident(dfw, lfw)) + "%");
        line("nucleotide identity on heavy chain CDR1-2 = " + f64_fixed1(ident(dc1 + dc2, l1 + l2)) + "%");
        line("nucleotide identity on heavy chain CDR2-3 = " + f64_fixed1(ident(dc2 + dc3, l2 + l3)) + "%");
        line("nucleotide identity on heavy chain CDR3 = " + f64_fixed1(ident(dc3, l3)) + "%");
    }
    line("p1 = prob of getting so many shares by accident = " + f64_display(e->p1));
    line("computed using k = " + std::to_string(e->k) + ", d = " + std::to_string(e->d) + ", n = " + std::to_string(e->n));
    line("mult = CDR3: partial_bernoulli_sum(3 * cn, cd as usize) = " + f64_display(e->mult));      // the label is stale upstream text
    line("score = p1 * mult = " + f64_display(e->score));
    for (int c = 0; c < 2; c++) {
        line("difference pattern for chain " + std::to_string(c + 1));
        const int n = (int)a.tig[c].size();
        std::string row;
        for (int p = 0; p < n; p++) {
            if (is_share[c][p]) row += "\xE2\x96\x93";                 // U+2593
            else row += a.tig[c][p] == b.tig[c][p] ? '-' : 'x';
            if ((p + 1) % 80 == 0 || p + 1 == n) { line(row); row.clear(); }
        }
    }
    *len_out = s.size();
    if (buf && cap) { const size_t m = std::min(cap - 1, s.size()); memcpy(buf, s.data(), m); buf[m] = 0; }
    return EJOIN_OK;
}
