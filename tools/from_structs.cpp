// from_structs.cpp — the join as the reference's caller sees it: array-of-struct inputs (Vec<CloneInfo>,
// Vec<ExactClonotype>, RefData) -> join_exacts() of enclone_b200/host/join_exacts.hpp -> raw_joins + EquivRel.
// Measures (a) the COLD call of a fresh process (CUDA context + handle, flatten on all host threads, first
// ejoin_run with its table building and allocations) — join_exacts runs once per enclone process, so this is
// the latency a user sees — and (b) warm repeats with a handle made beforehand.
// The structs are built from a synthetic repertoire (tools/synth.cpp) before the clock starts.
// usage: from_structs <workload C1..C5> <scale> <warm steps> <pinned 0|1>      prints one JSON line
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../enclone_b200/host/join_exacts.hpp"

extern "C" {
typedef struct esynth_cfg {
    uint64_t seed; uint32_t n_cells, n_donors, is_bcr, mode;
    double frac_naive, family_alpha, frac_public, frac_indel, frac_3chain, frac_4chain, frac_1chain;
    uint32_t max_family, reserved;
    void *(*alloc)(size_t); void (*dealloc)(void *);
} esynth_cfg_t;
void esynth_cfg_default(esynth_cfg_t *c);
int esynth_generate(const esynth_cfg_t *c, ejoin_pack_t **pack, uint32_t **truth_family);
void esynth_free(ejoin_pack_t *pack);
void esynth_free_truth(uint32_t *t);
}

using namespace enclone;

int main(int argc, char **argv) {
    const std::string name = argc > 1 ? argv[1] : "C3";
    const double scale = argc > 2 ? atof(argv[2]) : 1.0;
    const int warm = argc > 3 ? atoi(argv[3]) : 3;
    const bool pinned = argc > 4 ? atoi(argv[4]) != 0 : true;
    esynth_cfg_t cfg;
    esynth_cfg_default(&cfg);
    struct { const char *n; uint32_t cells, donors, bcr, mode; uint64_t seed; } W[] = {
        { "C1", 1654, 1, 1, 0, 20261020 }, { "C2", 100000, 1, 1, 0, 20261021 }, { "C3", 1600000, 8, 1, 0, 20261022 },
        { "C4", 200000, 1, 1, 1, 20261023 }, { "C5", 1000000, 4, 0, 0, 20261024 } };
    bool found = false;
    for (auto &w : W) if (name == w.n) { cfg.n_cells = std::max(16u, (uint32_t)(w.cells * scale)); cfg.n_donors = w.donors; cfg.is_bcr = w.bcr; cfg.mode = w.mode; cfg.seed = w.seed; found = true; }
    if (!found) return 2;
    ejoin_pack_t *pk = nullptr; uint32_t *truth = nullptr;
    if (esynth_generate(&cfg, &pk, &truth) != 0) return 3;
    esynth_free_truth(truth);
    // ---- array-of-struct inputs, as main_enclone_start holds them (untimed setup)
    const uint32_t n = pk->n_info, nx = pk->n_exact;
    RefData refdata;
    for (uint32_t r = 0; r < pk->n_refseq; r++) { refdata.refs.emplace_back((const char *)pk->refseq_arena + pk->refseq_off[r], pk->refseq_len[r]); refdata.name.push_back("ref" + std::to_string(r)); }
    std::vector<ExactClonotype> exacts(nx);
    for (uint32_t x = 0; x < nx; x++) {
        exacts[x].share.resize(pk->nchains[x]);
        exacts[x].clones.assign(pk->ncells[x], std::vector<TigData0>(pk->nchains[x]));
        for (uint32_t i = 0; i < pk->ncells[x]; i++)
            exacts[x].clones[i][0].donor_index = pk->donor_lo[x] == EJOIN_NONE ? -1 : (int64_t)(i == 0 ? pk->donor_lo[x] : pk->donor_hi[x]);
    }
    std::vector<uint32_t> seen(nx, 0);
    std::vector<CloneInfo> info(n);
    for (uint32_t k = 0; k < n; k++) {
        CloneInfo &ci = info[k];
        const uint32_t x = pk->exact_id[k];
        ci.clonotype_id = x;
        ExactClonotype &ex = exacts[x];
        const int nch = pk->len[1][k] ? 2 : 1;
        for (int c = 0; c < nch; c++) {
            // column of `share`: heavy = 0; the light chains of a 3-chain exact subclonotype take columns 1, 2 in order
            size_t col = c == 0 ? 0 : std::min<size_t>(1 + seen[x], ex.share.size() - 1);
            TigData1 &td = ex.share[col];
            td.left = c == 0; td.v_ref_id = pk->vuni_seq[c][k]; td.j_ref_id = pk->js_seq[c][k];
            td.u_ref_id = pk->utr_seq[c][k] == EJOIN_NONE ? -1 : (int64_t)pk->utr_seq[c][k];
            td.cdr3_start = pk->cdr3_start[c][k];
            if (c == 0) {
                auto o = [](uint16_t v) { return v == EJOIN_NONE16 ? -1 : (int64_t)v; };
                td.fr1_start = o(pk->fr1_start[k]); td.cdr1_start = o(pk->cdr1_start[k]); td.fr2_start = o(pk->fr2_start[k]);
                td.cdr2_start = o(pk->cdr2_start[k]); td.fr3_start = o(pk->fr3_start[k]); td.jun.hcomp = pk->hcomp[k];
            } else td.c_ref_id = pk->c_ref_id_light[k] == EJOIN_NONE ? -1 : (int64_t)pk->c_ref_id_light[k];
            ci.lens.push_back(pk->len[c][k]);
            ci.tigs.emplace_back((const char *)pk->tig_arena + pk->tig_off[c][k], pk->len[c][k]);
            ci.has_del.push_back(pk->has_del[c][k] != 0);
            ci.tigsp.push_back(ci.has_del.back() ? DnaString() : DnaString::from_acgt_bytes(ci.tigs.back()));
            ci.exact_cols.push_back(col);
            ci.vs.push_back(refdata.refs[pk->vs_seq[c][k]]); ci.js.push_back(refdata.refs[pk->js_seq[c][k]]);
            // ids: the generator's refseq indices double as refdata ids here (names "ref<i>"; the V-name gate then compares
            // universal sequences, which is what differing names lead to)
            ci.vsids.push_back(pk->vuni_seq[c][k]); ci.jsids.push_back(pk->js_seq[c][k]);
            ci.dref.push_back(pk->dref_id[c][k] == EJOIN_NONE ? -1 : (int64_t)pk->dref_id[c][k]);
            ci.cdr3s.emplace_back((const char *)pk->cdr3_arena + pk->cdr3_off[c][k], pk->cdr3_len[c][k]);
        }
        if (nch == 2) seen[x]++;
    }
    const bool is_bcr = pk->is_bcr != 0;
    esynth_free(pk);
    EncloneControl ctl;
    ctl.pinned = pinned;
    std::vector<std::pair<int32_t, int32_t>> raw_joins;
    JoinTimes cold;
    const double c0 = detail::now_ms();
    EquivRel eq = join_exacts(is_bcr, {}, refdata, ctl, exacts, info, raw_joins, {}, {}, nullptr, &cold);
    const double cold_total = detail::now_ms() - c0;
    const size_t joins_cold = raw_joins.size();
    std::vector<int32_t> reps;
    eq.orbit_reps(reps);
    // warm: the handle exists before the call (ejoin_create at program start)
    ejoin_handle_t *h = nullptr;
    const ejoin_params_t p = params_of(ctl);
    if (ejoin_create(&h, &p, 0) != 0) return 4;
    double best = 1e30, sum = 0;
    JoinTimes tbest;
    for (int i = 0; i < warm + 1; i++) {
        JoinTimes t;
        const double w0 = detail::now_ms();
        EquivRel e2 = join_exacts(is_bcr, {}, refdata, ctl, exacts, info, raw_joins, {}, {}, nullptr, &t, h);
        const double w = detail::now_ms() - w0;
        if (i == 0) continue;                        // the first call on this handle builds the tables
        sum += w;
        if (w < best) { best = w; tbest = t; }
        if (raw_joins.size() != joins_cold) return 5;
    }
    ejoin_destroy(h);
    printf("{\"workload\": \"%s\", \"scale\": %g, \"n_info\": %u, \"joins\": %zu, \"orbits\": %zu, \"threads\": %d, \"pinned\": %d, "
           "\"cold\": {\"ms_total\": %.3f, \"ms_create\": %.3f, \"ms_pack\": %.3f, \"ms_run\": %.3f, \"ms_unpack\": %.3f}, "
           "\"warm\": {\"ms_mean\": %.3f, \"ms_best\": %.3f, \"ms_pack\": %.3f, \"ms_run\": %.3f, \"ms_unpack\": %.3f, \"steps\": %d}, "
           "\"h2d_bytes\": %llu, \"d2h_bytes\": %llu}\n",
           name.c_str(), scale, n, joins_cold, reps.size(), cold.threads, (int)pinned, cold_total, cold.ms_create, cold.ms_pack, cold.ms_run, cold.ms_unpack,
           warm ? sum / warm : 0.0, best, tbest.ms_pack, tbest.ms_run, tbest.ms_unpack, warm, (unsigned long long)tbest.h2d_bytes, (unsigned long long)tbest.d2h_bytes);
    return 0;
}
