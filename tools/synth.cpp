// synth.cpp — synthetic V(D)J repertoire generator that emits a JoinPack (include/enclone_join.h).
//
// Test/bench infrastructure (neither product nor oracle).  It fabricates the inputs that
// enclone's join_exacts sees — Vec<CloneInfo> sorted with `lens` first, the exact
// subclonotype table and the germline table — with the shapes SURVEY.md §8(d) names:
// Zipf-skewed V/J usage, Gaussian CDR3H length, power-law clonal families with trunk +
// private somatic hypermutation, naive singletons, donor alleles, 3-/4-/1-chain exact
// subclonotypes, single-codon indels (has_del / edited V reference), public (cross-donor)
// rearrangements, a TCR mode without SHM, and an adversarial single-giant-bucket mode.
// Germline sequences are random but fixed by the seed (the real VDJ reference is not in
// /root/reference: SURVEY.md §4.3).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../include/enclone_join.h"

extern "C" {
typedef struct esynth_cfg {
    uint64_t seed;
    uint32_t n_cells;        // target number of cells
    uint32_t n_donors;       // 1..64
    uint32_t is_bcr;         // 1 = BCR (SHM), 0 = TCR (no SHM, rare per-base errors)
    uint32_t mode;           // 0 = repertoire, 1 = single giant bucket (n_cells = entries)
    double   frac_naive;     // fraction of cells that are naive singletons (BCR)   [0.55]
    double   family_alpha;   // power-law exponent of family sizes                  [2.3]
    double   frac_public;    // fraction of families duplicated into another donor  [0.02]
    double   frac_indel;     // fraction of members carrying a 3-nt indel           [0.01]
    double   frac_3chain;    // [0.03]
    double   frac_4chain;    // [0.005]
    double   frac_1chain;    // [0.04]
    uint32_t max_family;     // cap on exact subclonotypes per family               [4000]
    uint32_t reserved;
    void *(*alloc)(size_t);  // allocator for the big arenas (NULL = malloc)
    void (*dealloc)(void *);
} esynth_cfg_t;
void esynth_cfg_default(esynth_cfg_t *c);
int esynth_generate(const esynth_cfg_t *c, ejoin_pack_t **pack, uint32_t **truth_family);
void esynth_free(ejoin_pack_t *pack);
}

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t sm(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto &v : s) v = sm(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
    uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
    double gauss() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    char base() { return "ACGT"[next() >> 62]; }
    char other(char b) { for (;;) { char c = base(); if (c != b) return c; } }
};

struct VGene {
    std::string seq, utr, name;
    uint32_t name_id, leader;   // leader length = fr1_start
    uint32_t seq_id, utr_id;    // refseq ids
};
struct JGene { std::string seq; uint32_t seq_id, cdr3_part; /* J bases that belong to the CDR3 */ };

struct Locus {                 // one chain type (IGH, IGK, IGL, TRB, TRA)
    std::vector<VGene> v;
    std::vector<JGene> j;
    std::vector<double> vcum, jcum;
    uint32_t first_v_ref_id, first_j_ref_id;
    std::vector<uint32_t> c_ids;
};

struct Pack {
    ejoin_pack_t pk;
    esynth_cfg_t cfg;
    std::vector<uint32_t> exact_id, v_ref_id[2], j_ref_id[2], dref_id[2], vs_seq[2], js_seq[2], vname_id[2],
        vuni_seq[2], utr_seq[2], c_light, nchains, ncells, donor_lo, donor_hi, refseq_len;
    std::vector<uint16_t> len[2], cdr3_len[2], cdr3_start[2], hcomp, fr1, cdr1, fr2, cdr2, fr3;
    std::vector<uint64_t> tig_off[2], cdr3_off[2], refseq_off;
    std::vector<uint8_t> has_del[2];
    uint8_t *tig_arena = nullptr, *cdr3_arena = nullptr;
    std::vector<uint8_t> refseq_arena;
};

struct RefTable {
    std::vector<std::string> seqs;
    uint32_t add(const std::string &s) { seqs.push_back(s); return (uint32_t)seqs.size() - 1; }
};

void make_cum(std::vector<double> &cum, size_t n, double zipf_s, Rng &rng) {
    std::vector<uint32_t> perm(n);
    for (size_t i = 0; i < n; i++) perm[i] = (uint32_t)i;
    for (size_t i = n; i > 1; i--) std::swap(perm[i - 1], perm[rng.below((uint32_t)i)]);
    std::vector<double> w(n);
    double tot = 0;
    for (size_t r = 0; r < n; r++) { w[perm[r]] = 1.0 / std::pow((double)(r + 1), zipf_s); tot += w[perm[r]]; }
    cum.resize(n);
    double a = 0;
    for (size_t i = 0; i < n; i++) { a += w[i] / tot; cum[i] = a; }
    cum[n - 1] = 1.0;
}
uint32_t pick(const std::vector<double> &cum, Rng &rng) {
    double u = rng.uni();
    return (uint32_t)(std::lower_bound(cum.begin(), cum.end(), u) - cum.begin());
}

// V gene: leader + V region ending in the conserved Cys codon + 8 more bases, so that the
// CDR3 starts 11 nt before the V end.
void make_locus(Locus &L, const char *prefix, int nv, int vlen_lo, int vlen_hi, int nj, int jlen_lo, int jlen_hi,
                const char *jmotif, RefTable &refs, uint32_t &next_ref_id, uint32_t &next_name_id, Rng &rng,
                int nfam) {
    std::vector<std::string> founders;
    for (int f = 0; f < nfam; f++) {
        std::string s(vlen_hi, 'A');
        for (auto &ch : s) ch = rng.base();
        founders.push_back(s);
    }
    L.first_v_ref_id = next_ref_id;
    for (int g = 0; g < nv; g++) {
        VGene v;
        int vlen = vlen_lo + (int)rng.below((uint32_t)(vlen_hi - vlen_lo + 1));
        bool dup = (g > 0 && g % 9 == 8);     // near-duplicate of the previous gene ("IGKV3D-15" vs "IGKV3-15")
        if (dup) {
            v.seq = L.v[g - 1].seq;
            if (g % 2 == 0) { size_t p = 60 + rng.below(200); v.seq[p] = rng.other(v.seq[p]); }
            else if (rng.below(2)) v.seq.resize(v.seq.size() - 3 > 300 ? v.seq.size() - 3 : v.seq.size()); // identical up to 3' extension
            v.leader = L.v[g - 1].leader;
            v.utr = L.v[g - 1].utr;
            if (g % 3 == 0) v.utr = std::string("G") + v.utr; // longer 5'UTR, identical after left-truncation
        } else {
            const std::string &fd = founders[rng.below((uint32_t)nfam)];
            v.seq = fd.substr(0, (size_t)vlen);
            for (auto &ch : v.seq) if (rng.uni() < 0.10) ch = rng.other(ch);
            const char *tail = "TGTGCGAGAGA";
            for (int i = 0; i < 11; i++) v.seq[(size_t)(vlen - 11 + i)] = tail[i];
            v.seq[(size_t)(vlen - 5)] = rng.base();
            v.seq[(size_t)(vlen - 4)] = rng.base();
            v.leader = 54 + rng.below(7);
            int ul = 20 + (int)rng.below(60);
            v.utr.resize((size_t)ul);
            for (auto &ch : v.utr) ch = rng.base();
        }
        v.name = std::string(prefix) + "V" + std::to_string(g);
        v.name_id = next_name_id++;
        v.seq_id = refs.add(v.seq);
        v.utr_id = refs.add(v.utr);
        next_ref_id++;
        L.v.push_back(v);
    }
    L.first_j_ref_id = next_ref_id;
    for (int g = 0; g < nj; g++) {
        JGene j;
        int jlen = jlen_lo + (int)rng.below((uint32_t)(jlen_hi - jlen_lo + 1));
        int pl = jlen - 34;                    // bases before the W/F codon; CDR3 ends after it
        j.seq.resize((size_t)jlen);
        for (auto &ch : j.seq) ch = rng.base();
        for (int i = 0; i < 3; i++) j.seq[(size_t)(pl + i)] = jmotif[i];
        j.seq[(size_t)(pl + 3)] = 'G'; j.seq[(size_t)(pl + 4)] = 'G';
        j.cdr3_part = (uint32_t)(pl + 3);
        j.seq_id = refs.add(j.seq);
        next_ref_id++;
        L.j.push_back(j);
    }
    make_cum(L.vcum, L.v.size(), 1.0, rng);
    make_cum(L.jcum, L.j.size(), 0.8, rng);
}

struct Rearr {                 // one rearranged chain, germline state
    const Locus *loc;
    uint32_t vg, jg, c_id;
    std::string seq;            // V..J
    uint32_t vkeep, nlen, jtrim, cdr3_start, cdr3_len, hcomp;
};

Rearr rearrange(const Locus &L, int cdr3_nt_target, Rng &rng, int fix_v = -1, int fix_j = -1) {
    Rearr r;
    r.loc = &L;
    r.vg = fix_v >= 0 ? (uint32_t)fix_v : pick(L.vcum, rng);
    r.jg = fix_j >= 0 ? (uint32_t)fix_j : pick(L.jcum, rng);
    const VGene &v = L.v[r.vg];
    const JGene &j = L.j[r.jg];
    uint32_t vtrim = 0, jtrim = 0;
    while (vtrim < 6 && rng.uni() < 0.45) vtrim++;
    while (jtrim < j.cdr3_part - 3 && jtrim < 8 && rng.uni() < 0.5) jtrim++;
    int nlen = cdr3_nt_target - (11 - (int)vtrim) - ((int)j.cdr3_part - (int)jtrim);
    while (nlen < 0) {          // CDR3 too short for these trims: trim more
        if (vtrim < 8) { vtrim++; nlen++; }
        else if (jtrim + 3 < j.cdr3_part) { jtrim++; nlen++; }
        else { nlen = 0; break; }
    }
    r.vkeep = (uint32_t)v.seq.size() - vtrim;
    r.nlen = (uint32_t)nlen;
    r.jtrim = jtrim;
    r.seq = v.seq.substr(0, r.vkeep);
    for (int i = 0; i < nlen; i++) r.seq.push_back(rng.base());
    r.seq += j.seq.substr(jtrim);
    r.cdr3_start = (uint32_t)v.seq.size() - 11;
    r.cdr3_len = (11 - vtrim) + (uint32_t)nlen + (j.cdr3_part - jtrim);
    r.hcomp = (uint32_t)nlen + (vtrim > 0) + (jtrim > 0);
    r.c_id = L.c_ids.empty() ? EJOIN_NONE : L.c_ids[rng.below((uint32_t)L.c_ids.size())];
    return r;
}

void mutate(std::string &s, uint32_t lo, double rate, Rng &rng) {
    if (rate <= 0) return;
    uint32_t n = (uint32_t)s.size();
    if (lo >= n) return;
    double expect = rate * (n - lo);
    uint32_t m = (uint32_t)expect;
    if (rng.uni() < expect - m) m++;
    for (uint32_t i = 0; i < m; i++) {
        uint32_t p = lo + rng.below(n - lo);
        if (s[p] != '-') s[p] = rng.other(s[p]);
    }
}

struct Member {                // one exact subclonotype
    std::string h, l, l2, h2;   // V..J sequences (l2/h2: extra chains)
    Rearr *rh, *rl, *rl2, *rh2;
    uint32_t ncells, donor, family;
    uint8_t del_h = 0;          // has_del on heavy
    int ins_h = -1;             // insertion position on heavy (edited reference), -1 none
    uint32_t nch;
};

struct Entry {                 // one CloneInfo
    uint32_t exact;
    const std::string *t[2];
    const Rearr *r[2];
    uint8_t hd[2];
    int ins[2];
    uint32_t donor, family;
};

}  // namespace

extern "C" void esynth_cfg_default(esynth_cfg_t *c) {
    memset(c, 0, sizeof(*c));
    c->seed = 20261019;
    c->n_cells = 10000;
    c->n_donors = 1;
    c->is_bcr = 1;
    c->frac_naive = 0.55;
    c->family_alpha = 2.3;
    c->frac_public = 0.02;
    c->frac_indel = 0.01;
    c->frac_3chain = 0.03;
    c->frac_4chain = 0.005;
    c->frac_1chain = 0.04;
    c->max_family = 4000;
}

extern "C" int esynth_generate(const esynth_cfg_t *cfgp, ejoin_pack_t **out, uint32_t **truth_out) {
    esynth_cfg_t cfg = *cfgp;
    if (cfg.n_donors < 1) cfg.n_donors = 1;
    if (cfg.n_donors > 64) cfg.n_donors = 64;
    Rng rng(cfg.seed);
    RefTable refs;
    uint32_t next_ref_id = 0, next_name_id = 0;
    Locus H, K, Lm;
    if (cfg.is_bcr) {
        make_locus(H, "IGH", 55, 350, 365, 6, 46, 63, "TGG", refs, next_ref_id, next_name_id, rng, 7);
        make_locus(K, "IGK", 40, 340, 365, 5, 37, 39, "TTC", refs, next_ref_id, next_name_id, rng, 6);
        make_locus(Lm, "IGL", 30, 340, 365, 5, 37, 39, "TTC", refs, next_ref_id, next_name_id, rng, 5);
        K.c_ids = { next_ref_id++ };
        Lm.c_ids = { next_ref_id++, next_ref_id++, next_ref_id++, next_ref_id++ };
    } else {
        make_locus(H, "TRB", 48, 340, 350, 13, 47, 53, "TTC", refs, next_ref_id, next_name_id, rng, 8);
        make_locus(K, "TRA", 45, 330, 345, 50, 55, 66, "TTC", refs, next_ref_id, next_name_id, rng, 8);
        K.c_ids = { next_ref_id++ };
    }
    // donor alleles: (donor, locus, gene) -> (dref id, refseq id)
    struct Allele { uint32_t dref, seq_id; };
    std::map<uint64_t, Allele> alleles;
    uint32_t next_dref = 0;
    Locus *loci[3] = { &H, &K, &Lm };
    int nloci = cfg.is_bcr ? 3 : 2;
    for (uint32_t d = 0; d < cfg.n_donors; d++)
        for (int li = 0; li < nloci; li++)
            for (size_t g = 0; g < loci[li]->v.size(); g++)
                if (rng.uni() < 0.15) {
                    std::string s = loci[li]->v[g].seq;
                    int ns = 1 + (int)rng.below(3);
                    for (int i = 0; i < ns; i++) { size_t p = 60 + rng.below((uint32_t)s.size() - 80); s[p] = rng.other(s[p]); }
                    alleles[((uint64_t)d << 32) | ((uint64_t)li << 16) | g] = Allele{ next_dref++, refs.add(s) };
                }
    // donor sizes: skewed
    std::vector<double> dcum(cfg.n_donors);
    { double tot = 0; std::vector<double> w(cfg.n_donors);
      for (uint32_t d = 0; d < cfg.n_donors; d++) { w[d] = 1.0 / (1.0 + 0.6 * d); tot += w[d]; }
      double a = 0; for (uint32_t d = 0; d < cfg.n_donors; d++) { a += w[d] / tot; dcum[d] = a; } dcum.back() = 1.0; }

    std::vector<Rearr *> rearrs;
    std::vector<Member> members;
    auto new_rearr = [&](const Rearr &r) { rearrs.push_back(new Rearr(r)); return rearrs.back(); };
    auto cdr3h_target = [&]() {
        int aa = (int)std::lround(16.0 + 4.0 * rng.gauss());
        if (aa < 5) aa = 5; if (aa > 35) aa = 35;
        return 3 * aa;
    };
    auto cdr3l_target = [&]() { double u = rng.uni(); return 3 * (u < 0.6 ? 9 : (u < 0.85 ? 10 : 11)); };
    auto light_locus = [&]() -> const Locus & { return (!cfg.is_bcr || rng.uni() < 0.6) ? K : Lm; };

    uint32_t cells = 0, family_id = 0;
    const bool giant = cfg.mode == 1;
    // giant-bucket mode: every entry shares V, J, V..J lengths and CDR3 lengths
    Rearr giant_h, giant_l;
    if (giant) {
        giant_h = rearrange(H, 48, rng, 0, 0);
        giant_l = rearrange(K, 27, rng, 0, 0);
    }
    auto make_family_rearr = [&](Rearr &rh, Rearr &rl) {
        if (giant) {
            rh = giant_h; rl = giant_l;
            // fresh junction of the same length: unrelated CDR3s
            for (uint32_t i = 0; i < rh.nlen; i++) rh.seq[rh.vkeep + i] = rng.base();
            for (uint32_t i = 0; i < rl.nlen; i++) rl.seq[rl.vkeep + i] = rng.base();
            // a few more junction bases vary (P/N additions replacing templated ones)
            for (uint32_t p = rh.cdr3_start + 6; p < rh.cdr3_start + rh.cdr3_len - 6; p++) if (rng.uni() < 0.5) rh.seq[p] = rng.base();
            for (uint32_t p = rl.cdr3_start + 9; p < rl.cdr3_start + rl.cdr3_len - 6; p++) if (rng.uni() < 0.5) rl.seq[p] = rng.base();
        } else {
            rh = rearrange(H, cdr3h_target(), rng);
            rl = rearrange(light_locus(), cfg.is_bcr ? cdr3l_target() : 3 * (11 + (int)rng.below(5)), rng);
        }
    };
    const uint32_t target_cells = cfg.n_cells;
    while (cells < target_cells) {
        Rearr rh0, rl0;
        make_family_rearr(rh0, rl0);
        uint32_t donor = pick(dcum, rng);
        int copies = (!giant && cfg.n_donors > 1 && rng.uni() < cfg.frac_public) ? 2 : 1;
        for (int cp = 0; cp < copies && cells < target_cells; cp++) {
            if (cp == 1) donor = (donor + 1 + rng.below(cfg.n_donors - 1)) % cfg.n_donors;
            Rearr *rh = new_rearr(rh0), *rl = new_rearr(rl0);
            bool naive;
            uint32_t fam_size;
            if (giant) {
                naive = rng.uni() < 0.8;
                fam_size = naive ? 1 : (uint32_t)(1 + rng.below(39)); // mean 20
            } else if (cfg.is_bcr) {
                naive = rng.uni() < cfg.frac_naive / (cfg.frac_naive + (1 - cfg.frac_naive) / 4.0); // per family, memory families avg ~4 cells
                double u = rng.uni(); if (u < 1e-12) u = 1e-12;
                fam_size = naive ? 1 : (uint32_t)std::min<double>(cfg.max_family, std::floor(std::pow(u, -1.0 / (cfg.family_alpha - 1.0))));
            } else {
                naive = true;
                double u = rng.uni(); if (u < 1e-12) u = 1e-12;
                fam_size = (uint32_t)std::min<double>(cfg.max_family, std::floor(std::pow(u, -1.0 / (cfg.family_alpha - 1.0))));
            }
            std::string th = rh->seq, tl = rl->seq;
            double trunk = 0;
            if (cfg.is_bcr && !naive) {
                trunk = 0.01 + 0.07 * rng.uni();
                mutate(th, H.v[rh->vg].leader, trunk, rng);
                mutate(tl, rl->loc->v[rl->vg].leader, trunk * 0.6, rng);
            }
            size_t fam_begin = members.size();
            for (uint32_t m = 0; m < fam_size && cells < target_cells; m++) {
                Member mb;
                mb.rh = rh; mb.rl = rl; mb.rl2 = nullptr; mb.rh2 = nullptr;
                mb.donor = donor; mb.family = family_id;
                if (cfg.is_bcr && !naive) {
                    // derive from the trunk or from an earlier member of the family
                    if (m > 0 && rng.uni() < 0.5) {
                        const Member &par = members[fam_begin + rng.below(m)];
                        mb.h = par.h; mb.l = par.l;
                        if (par.del_h || par.ins_h >= 0) { mb.h = th; mb.l = tl; }
                    } else { mb.h = th; mb.l = tl; }
                    double priv = 0.04 * rng.uni();
                    if (rng.uni() < 0.15) priv = 0;
                    mutate(mb.h, H.v[rh->vg].leader, priv, rng);
                    mutate(mb.l, rl->loc->v[rl->vg].leader, priv * 0.6, rng);
                } else if (!cfg.is_bcr) {
                    mb.h = th; mb.l = tl;
                    if (m > 0) { mutate(mb.h, 0, 0.003, rng); mutate(mb.l, 0, 0.003, rng); }
                } else { mb.h = th; mb.l = tl; }
                mb.ncells = 1;
                while (rng.uni() < (naive && cfg.is_bcr ? 0.05 : 0.4) && mb.ncells < 50) mb.ncells++;
                // indels (heavy chain, inside FWR3, single codon)
                if (cfg.is_bcr && rng.uni() < cfg.frac_indel) {
                    uint32_t p = 200 + 3 * rng.below(20);
                    if (giant || rng.uni() < 0.6) { mb.h[p] = mb.h[p + 1] = mb.h[p + 2] = '-'; mb.del_h = 1; }
                    else { std::string ins3; for (int i = 0; i < 3; i++) ins3.push_back(rng.base()); mb.h.insert(p, ins3); mb.ins_h = (int)p; }
                }
                mb.nch = 2;
                double u = rng.uni();
                if (!giant) {
                    if (u < cfg.frac_3chain) {
                        mb.nch = 3;
                        mb.rl2 = new_rearr(rearrange(light_locus(), cfg.is_bcr ? cdr3l_target() : 36, rng));
                        mb.l2 = mb.rl2->seq;
                    } else if (u < cfg.frac_3chain + cfg.frac_4chain) {
                        mb.nch = 4;
                        mb.rl2 = new_rearr(rearrange(light_locus(), cfg.is_bcr ? cdr3l_target() : 36, rng));
                        mb.l2 = mb.rl2->seq;
                        mb.rh2 = new_rearr(rearrange(H, cdr3h_target(), rng));
                        mb.h2 = mb.rh2->seq;
                    } else if (u < cfg.frac_3chain + cfg.frac_4chain + cfg.frac_1chain) {
                        mb.nch = 1;
                    }
                }
                cells += giant ? 1 : mb.ncells;         // the giant-bucket config counts exact subclonotypes, not cells
                members.push_back(std::move(mb));
            }
            family_id++;
        }
    }
    // info entries: one per (heavy, light) chain pair of each exact subclonotype
    std::vector<Entry> ent;
    ent.reserve(members.size() * 11 / 10);
    for (uint32_t x = 0; x < members.size(); x++) {
        const Member &m = members[x];
        auto push = [&](const std::string *h, const Rearr *rh, uint8_t hdh, int insh, const std::string *l, const Rearr *rl) {
            Entry e; e.exact = x; e.t[0] = h; e.t[1] = l; e.r[0] = rh; e.r[1] = rl; e.hd[0] = hdh; e.hd[1] = 0;
            e.ins[0] = insh; e.ins[1] = -1; e.donor = m.donor; e.family = m.family; ent.push_back(e);
        };
        if (m.nch == 1) push(&m.h, m.rh, m.del_h, m.ins_h, nullptr, nullptr);
        else {
            push(&m.h, m.rh, m.del_h, m.ins_h, &m.l, m.rl);
            if (m.nch >= 3) push(&m.h, m.rh, m.del_h, m.ins_h, &m.l2, m.rl2);
            if (m.nch == 4) { push(&m.h2, m.rh2, 0, -1, &m.l, m.rl); push(&m.h2, m.rh2, 0, -1, &m.l2, m.rl2); }
        }
    }
    // sort like CloneInfo's derived Ord: lens first, then the tigs bytes
    const uint32_t n = (uint32_t)ent.size();
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) order[i] = i;
    auto L0 = [&](uint32_t i) { return (uint32_t)ent[i].t[0]->size(); };
    auto L1 = [&](uint32_t i) { return ent[i].t[1] ? (uint32_t)ent[i].t[1]->size() : 0u; };
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        uint64_t ka = ((uint64_t)L0(a) << 32) | L1(a), kb = ((uint64_t)L0(b) << 32) | L1(b);
        if (ka != kb) return ka < kb;
        return a < b;
    });
    {   // within buckets, in parallel
        std::vector<std::pair<uint32_t, uint32_t>> runs;
        for (uint32_t i = 0; i < n;) {
            uint32_t j = i + 1;
            while (j < n && L0(order[j]) == L0(order[i]) && L1(order[j]) == L1(order[i])) j++;
            runs.push_back({ i, j }); i = j;
        }
        unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th;
        std::atomic<size_t> nextrun{ 0 };
        auto cmp = [&](uint32_t a, uint32_t b) {
            int c = ent[a].t[0]->compare(*ent[b].t[0]);
            if (c) return c < 0;
            if (ent[a].t[1] && ent[b].t[1]) { c = ent[a].t[1]->compare(*ent[b].t[1]); if (c) return c < 0; }
            return a < b;
        };
        for (unsigned t = 0; t < nt; t++)
            th.emplace_back([&]() {
                for (;;) { size_t r = nextrun.fetch_add(1); if (r >= runs.size()) break;
                    std::sort(order.begin() + runs[r].first, order.begin() + runs[r].second, cmp); }
            });
        for (auto &t : th) t.join();
    }
    // fill the pack
    Pack *P = new Pack();
    P->cfg = cfg;
    auto A = [&](size_t bytes) -> uint8_t * { return (uint8_t *)(cfg.alloc ? cfg.alloc(bytes ? bytes : 16) : malloc(bytes ? bytes : 16)); };
    uint64_t tig_bytes = 0, cdr3_bytes = 0;
    for (uint32_t i = 0; i < n; i++) {
        const Entry &e = ent[i];
        tig_bytes += e.t[0]->size() + (e.t[1] ? e.t[1]->size() : 0);
        cdr3_bytes += e.r[0]->cdr3_len + (e.r[1] ? e.r[1]->cdr3_len : 0);
    }
    P->tig_arena = A(tig_bytes + 64); P->cdr3_arena = A(cdr3_bytes + 64);
    for (int c = 0; c < 2; c++) {
        P->len[c].resize(n); P->tig_off[c].resize(n); P->has_del[c].resize(n); P->cdr3_len[c].resize(n);
        P->cdr3_off[c].resize(n); P->cdr3_start[c].resize(n); P->v_ref_id[c].resize(n); P->j_ref_id[c].resize(n);
        P->dref_id[c].resize(n); P->vs_seq[c].resize(n); P->js_seq[c].resize(n); P->vname_id[c].resize(n);
        P->vuni_seq[c].resize(n); P->utr_seq[c].resize(n);
    }
    P->exact_id.resize(n); P->c_light.resize(n); P->hcomp.resize(n);
    P->fr1.resize(n); P->cdr1.resize(n); P->fr2.resize(n); P->cdr2.resize(n); P->fr3.resize(n);
    uint32_t *truth = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint64_t to = 0, co = 0;
    auto locus_index = [&](const Locus *l) { return l == &H ? 0 : (l == &K ? 1 : 2); };
    for (uint32_t kk = 0; kk < n; kk++) {
        const Entry &e = ent[order[kk]];
        P->exact_id[kk] = e.exact;
        truth[kk] = e.family;
        for (int c = 0; c < 2; c++) {
            if (!e.t[c]) {
                P->len[c][kk] = 0; P->tig_off[c][kk] = to; P->has_del[c][kk] = 0; P->cdr3_len[c][kk] = 0;
                P->cdr3_off[c][kk] = co; P->cdr3_start[c][kk] = 0;
                P->v_ref_id[c][kk] = P->j_ref_id[c][kk] = P->dref_id[c][kk] = EJOIN_NONE;
                P->vs_seq[c][kk] = P->js_seq[c][kk] = P->vuni_seq[c][kk] = 0; P->vname_id[c][kk] = EJOIN_NONE;
                P->utr_seq[c][kk] = EJOIN_NONE;
                continue;
            }
            const std::string &t = *e.t[c];
            const Rearr &r = *e.r[c];
            const VGene &v = r.loc->v[r.vg];
            const JGene &j = r.loc->j[r.jg];
            P->len[c][kk] = (uint16_t)t.size();
            P->tig_off[c][kk] = to;
            memcpy(P->tig_arena + to, t.data(), t.size());
            to += t.size();
            P->has_del[c][kk] = e.hd[c];
            uint32_t cs = r.cdr3_start + (e.ins[c] >= 0 ? 3 : 0);
            P->cdr3_start[c][kk] = (uint16_t)cs;
            P->cdr3_len[c][kk] = (uint16_t)r.cdr3_len;
            P->cdr3_off[c][kk] = co;
            memcpy(P->cdr3_arena + co, t.data() + cs, r.cdr3_len);   // cdr3_dna is the contig's CDR3
            co += r.cdr3_len;
            P->v_ref_id[c][kk] = r.loc->first_v_ref_id + r.vg;
            P->j_ref_id[c][kk] = r.loc->first_j_ref_id + r.jg;
            auto it = alleles.find(((uint64_t)e.donor << 32) | ((uint64_t)locus_index(r.loc) << 16) | r.vg);
            uint32_t vs = v.seq_id;
            P->dref_id[c][kk] = EJOIN_NONE;
            if (it != alleles.end()) { P->dref_id[c][kk] = it->second.dref; vs = it->second.seq_id; }
            if (e.ins[c] >= 0) {   // insertion: the V reference is edited with the contig's bases (help1.rs:236-240)
                std::string ed = refs.seqs[vs];
                ed.insert((size_t)e.ins[c], t.substr((size_t)e.ins[c], 3));
                vs = refs.add(ed);
            }
            P->vs_seq[c][kk] = vs;
            P->js_seq[c][kk] = j.seq_id;
            P->vname_id[c][kk] = v.name_id;
            P->vuni_seq[c][kk] = v.seq_id;
            P->utr_seq[c][kk] = (r.vg % 11 == 10) ? EJOIN_NONE : v.utr_id;
        }
        P->c_light[kk] = e.r[1] ? e.r[1]->c_id : EJOIN_NONE;
        P->hcomp[kk] = (uint16_t)e.r[0]->hcomp;
        const VGene &hv = e.r[0]->loc->v[e.r[0]->vg];
        uint32_t sh = e.ins[0] >= 0 ? 3 : 0;
        if (e.r[0]->vg % 13 == 12) {   // some genes lack region annotation (Option::None)
            P->fr1[kk] = (uint16_t)hv.leader; P->cdr1[kk] = P->fr2[kk] = P->cdr2[kk] = P->fr3[kk] = EJOIN_NONE16;
        } else {
            P->fr1[kk] = (uint16_t)hv.leader; P->cdr1[kk] = (uint16_t)(hv.leader + 75);
            P->fr2[kk] = (uint16_t)(hv.leader + 99); P->cdr2[kk] = (uint16_t)(hv.leader + 150 + sh);
            P->fr3[kk] = (uint16_t)(hv.leader + 174 + sh);
        }
    }
    const uint32_t nx = (uint32_t)members.size();
    P->nchains.resize(nx); P->ncells.resize(nx); P->donor_lo.resize(nx); P->donor_hi.resize(nx);
    for (uint32_t x = 0; x < nx; x++) {
        P->nchains[x] = members[x].nch; P->ncells[x] = members[x].ncells;
        P->donor_lo[x] = P->donor_hi[x] = members[x].donor;
        if (cfg.n_donors > 1 && x % 97 == 96) P->donor_lo[x] = P->donor_hi[x] = EJOIN_NONE; // unknown donor
    }
    const uint32_t nr = (uint32_t)refs.seqs.size();
    P->refseq_off.resize(nr); P->refseq_len.resize(nr);
    for (uint32_t r = 0; r < nr; r++) {
        P->refseq_off[r] = P->refseq_arena.size(); P->refseq_len[r] = (uint32_t)refs.seqs[r].size();
        P->refseq_arena.insert(P->refseq_arena.end(), refs.seqs[r].begin(), refs.seqs[r].end());
        while (P->refseq_arena.size() % 16) P->refseq_arena.push_back(0);
    }
    ejoin_pack_t &pk = P->pk;
    memset(&pk, 0, sizeof(pk));
    pk.abi_version = EJOIN_ABI_VERSION; pk.n_info = n; pk.n_exact = nx; pk.n_refseq = nr; pk.is_bcr = cfg.is_bcr;
    // arenas from a caller-supplied (pinned) allocator carry 64 bytes of padding: they may be read in place
    pk.flags = cfg.alloc ? EJOIN_F_TIGS_IN_PLACE : 0;
    pk.exact_id = P->exact_id.data();
    for (int c = 0; c < 2; c++) {
        pk.len[c] = P->len[c].data(); pk.tig_off[c] = P->tig_off[c].data(); pk.has_del[c] = P->has_del[c].data();
        pk.cdr3_len[c] = P->cdr3_len[c].data(); pk.cdr3_off[c] = P->cdr3_off[c].data(); pk.cdr3_start[c] = P->cdr3_start[c].data();
        pk.v_ref_id[c] = P->v_ref_id[c].data(); pk.j_ref_id[c] = P->j_ref_id[c].data(); pk.dref_id[c] = P->dref_id[c].data();
        pk.vs_seq[c] = P->vs_seq[c].data(); pk.js_seq[c] = P->js_seq[c].data(); pk.vname_id[c] = P->vname_id[c].data();
        pk.vuni_seq[c] = P->vuni_seq[c].data(); pk.utr_seq[c] = P->utr_seq[c].data();
    }
    pk.tig_arena = P->tig_arena; pk.tig_arena_bytes = to; pk.cdr3_arena = P->cdr3_arena; pk.cdr3_arena_bytes = co;
    pk.c_ref_id_light = P->c_light.data(); pk.hcomp = P->hcomp.data();
    pk.fr1_start = P->fr1.data(); pk.cdr1_start = P->cdr1.data(); pk.fr2_start = P->fr2.data();
    pk.cdr2_start = P->cdr2.data(); pk.fr3_start = P->fr3.data();
    pk.nchains = P->nchains.data(); pk.ncells = P->ncells.data(); pk.donor_lo = P->donor_lo.data(); pk.donor_hi = P->donor_hi.data();
    pk.refseq_off = P->refseq_off.data(); pk.refseq_len = P->refseq_len.data(); pk.refseq_arena = P->refseq_arena.data();
    pk.refseq_arena_bytes = P->refseq_arena.size();
    for (auto *r : rearrs) delete r;
    *out = &P->pk;
    if (truth_out) *truth_out = truth; else free(truth);
    return 0;
}

extern "C" void esynth_free(ejoin_pack_t *pack) {
    if (!pack) return;
    Pack *P = reinterpret_cast<Pack *>(pack);   // pk is the first member
    if (P->cfg.dealloc) { P->cfg.dealloc(P->tig_arena); P->cfg.dealloc(P->cdr3_arena); }
    else { free(P->tig_arena); free(P->cdr3_arena); }
    delete P;
}
extern "C" void esynth_free_truth(uint32_t *t) { free(t); }
