// Exercises the C++ host mirror (enclone_b200/host/join_exacts.hpp) the way the reference's caller
// (main_enclone_start) uses join_exacts: build ExactClonotype / CloneInfo structs, call, read back
// raw_joins + EquivRel.  Input: the reference's known-answer pair (test 256) passed as a text file of
// fields written by tests/test_host_mirror.py, replicated into a few entries.  Prints the results as
// JSON for the Python test to check against the golden values and the oracle.
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../enclone_b200/host/join_exacts.hpp"

using namespace enclone;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1]);
    size_t n_entries, n_refs;
    int mix;
    f >> n_entries >> n_refs >> mix;
    RefData refdata;
    for (size_t r = 0; r < n_refs; r++) { std::string name, seq; f >> name >> seq; refdata.name.push_back(name); refdata.refs.push_back(seq); }
    std::vector<ExactClonotype> exacts;
    std::vector<CloneInfo> info;
    for (size_t k = 0; k < n_entries; k++) {
        ExactClonotype ex;
        CloneInfo ci;
        int64_t donor;
        f >> donor;
        for (int c = 0; c < 2; c++) {
            TigData1 td;
            std::string tig, cdr3, vs, js;
            size_t vid, jid;
            int64_t cref;
            f >> tig >> cdr3 >> td.cdr3_start >> vid >> jid >> vs >> js >> cref >> td.fr1_start >> td.cdr1_start >> td.fr2_start >> td.cdr2_start >> td.fr3_start >> td.jun.hcomp;
            td.left = c == 0; td.v_ref_id = vid; td.j_ref_id = jid; td.c_ref_id = cref;
            ex.share.push_back(td);
            ci.lens.push_back(tig.size()); ci.tigs.push_back(tig); ci.has_del.push_back(false); ci.exact_cols.push_back((size_t)c);
            ci.vs.push_back(vs); ci.js.push_back(js); ci.vsids.push_back(vid); ci.jsids.push_back(jid); ci.dref.push_back(-1); ci.cdr3s.push_back(cdr3);
        }
        ex.clones.push_back({ TigData0{ donor }, TigData0{ donor } });
        ci.clonotype_id = exacts.size();
        exacts.push_back(ex);
        info.push_back(ci);
    }
    EncloneControl ctl;
    ctl.mix_donors = mix != 0;
    std::vector<std::pair<int32_t, int32_t>> raw_joins;
    std::vector<PotentialJoin> ji;
    EquivRel eq = join_exacts(true, {}, refdata, ctl, exacts, info, raw_joins, {}, {}, &ji);
    std::vector<int32_t> reps;
    eq.orbit_reps(reps);
    printf("{\"n\": %d, \"norbits\": %zu, \"raw_joins\": [", eq.size(), reps.size());
    for (size_t i = 0; i < raw_joins.size(); i++) printf("%s[%d,%d]", i ? "," : "", raw_joins[i].first, raw_joins[i].second);
    printf("], \"joins\": [");
    for (size_t i = 0; i < ji.size(); i++)
        printf("%s{\"shares\": %d, \"indeps\": %d, \"cd\": %d, \"p1\": %.17g, \"mult\": %.17g, \"score\": %.17g, \"err\": %d}", i ? "," : "",
               ji[i].shares[0], ji[i].indeps[0], ji[i].cd, ji[i].p1, ji[i].mult, ji[i].score, (int)ji[i].err);
    printf("], \"class_of\": [");
    for (int32_t k = 0; k < eq.size(); k++) printf("%s%d", k ? "," : "", eq.class_id(k));
    printf("]}\n");
    return 0;
}
