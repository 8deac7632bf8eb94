"""Pins the CPU oracle to the reference's own golden vector for join_one
(enclone_exec/testx/inputs/outputs/enclone_test256_output:29-70)."""
import math

import numpy as np

import oracle_lib as O
from enclone_b200._abi import default_params
from kat import kat_pack


def test_p1_mult_score_bit_exact():
    _, kat = kat_pack()
    e = kat["expected"]
    p1 = O.p1(e["k"], e["d"], e["n"])
    assert p1 == float(e["p1"]) == 1.6600942157683246e-6
    mult = math.pow(80.0, 42.0 * 0 / 18 + 42.0 * 2 / 36)
    assert mult == float(e["mult"])
    assert p1 * mult == float(e["score"])


def test_join_one_reproduces_golden_block():
    pack, kat = kat_pack()
    e = kat["expected"]
    # MIN_DONORS=2 in the test command turns MIX_DONORS on (main_testlist.rs:625-629)
    ok, edge = O.join_one(pack.struct, default_params(mix_donors=1), 0, 1)
    assert ok
    assert edge.nrefs == 1
    assert edge.shares[0] == e["shares"] and edge.indeps[0] == e["indeps"]
    assert list(edge.shares_details[0]) == e["shares_details"]
    assert edge.cd == e["cd"] and list(edge.cd_chain) == [e["cdr3h_diffs"], e["cd"] - e["cdr3h_diffs"]]
    assert (edge.k, edge.d, edge.n) == (e["k"], e["d"], e["n"])
    assert edge.p1 == float(e["p1"])
    assert edge.mult == float(e["mult"])
    assert edge.score == float(e["score"])
    assert edge.err == 1 and e["join_error"]          # "JOIN ERROR": donors differ
    assert edge.diffs == 16 + 8                        # 16 heavy + 8 light differing positions


def test_donor_gate_without_mix_donors():
    pack, _ = kat_pack()
    ok, _ = O.join_one(pack.struct, default_params(mix_donors=0), 0, 1)
    assert not ok


def test_region_identities():
    pack, kat = kat_pack()
    e = kat["expected"]
    ok, (dfw, lfw, d12, l12) = O.region_counts(pack.struct, 0, 1)
    assert ok
    assert dfw == e["fwr1_diffs"] and d12 == e["cdr1_diffs"] + e["cdr2_diffs"]
    assert "%.1f" % (100.0 * (1 - dfw / lfw)) == e["fwr1_ident"]
    assert "%.1f" % (100.0 * (1 - d12 / l12)) == e["cdr12_ident"]


def test_join_exacts_on_the_pair():
    pack, _ = kat_pack()
    edges, cls, st = O.join_exacts(pack.struct, default_params(mix_donors=1))
    assert len(edges) == 1 and (edges[0]["k1"], edges[0]["k2"]) == (0, 1)
    assert list(cls) == [0, 0]
    # two-cell rule (help1.rs:349-353): cd=2 <= d=5 so the join stands; it falls when cd > d
    edges, cls, _ = O.join_exacts(pack.struct, default_params(mix_donors=1, auto_share=1000000, max_score=1e300))
    assert len(edges) == 1
    edges, cls, _ = O.join_exacts(pack.struct, default_params(mix_donors=0))
    assert len(edges) == 0 and list(cls) == [0, 1]


def test_p1_against_exact_rationals():
    from fractions import Fraction

    def s2row(x):
        row = [0] * (x + 1); row[0] = 1
        for i in range(1, x + 1):
            new = [0] * (x + 1)
            for u in range(1, i + 1):
                new[u] = u * row[u] + row[u - 1]
            row = new
        return row

    rng = np.random.default_rng(7)
    for _ in range(60):
        k = int(rng.integers(1, 300)); d = int(rng.integers(0, min(k, 14) + 1)); n = 3 * int(rng.integers(500, 1000))
        S = s2row(k)
        tot = Fraction(0)
        for u in range(k - d + 1, k + 1):
            f = 1
            for v in range(u):
                f *= (n - v)
            tot += Fraction(S[u] * f, n ** k)
        exact = float(1 - tot)
        got = O.p1(k, d, n)
        # double-double "1 - sum" carries ~1e-31 absolute cancellation noise
        assert abs(got - exact) <= 1e-29 + 2e-16 * exact, (k, d, n, got, exact)


def test_tile_parallel_bucket_mode_equals_the_sequential_scan():
    """The parallel mode used for the 200,000-entry bucket (row-parallel scoring + ordered replay,
    SURVEY App. A.8) must give join_core's own result: same ordered edges, same partition."""
    from enclone_b200 import synth
    cases = [synth.generate_config("C4", 0.03)[0],                                     # one bucket, 6,000 entries
             synth.generate(n_cells=5000, n_donors=3, seed=17)[0],                      # many buckets
             synth.generate(n_cells=4000, n_donors=2, is_bcr=False, seed=18)[0]]        # TCR rules (cd = 0)
    for pack in cases:
        for over in (dict(), dict(mix_donors=1, join_cdr3_ident=70.0, max_cdr3_diffs=7)):
            params = default_params(**over)
            O.set_parallel_bucket(0, 1)
            e0, c0, s0 = O.join_exacts(pack.struct, params, threads=4)
            try:
                O.set_parallel_bucket(2, 4)
                e1, c1, s1 = O.join_exacts(pack.struct, params, threads=1)
            finally:
                O.set_parallel_bucket(0, 1)
            assert np.array_equal(e0, e1) and np.array_equal(c0, c1)
            assert s0["pairs_candidate"] == s1["pairs_candidate"]


def test_max_degradation_readings():
    """help1.rs:345-348 (DOC, mode 0: the assigned references may differ in at most MAX_DEGRADATION
    positions) against the recollected rule (mode 1: |shares0 - shares1|).  On the golden pair both
    entries carry the same references, so nrefs = 1 and neither reading applies."""
    from enclone_b200 import synth
    pack, _ = kat_pack()
    for mode in (0, 1):
        ok, edge = O.join_one(pack.struct, default_params(mix_donors=1, degradation_mode=mode), 0, 1)
        assert ok and edge.nrefs == 1
    assert O.ref_positions_differing(pack.struct, 0, 1) == 0
    # a repertoire with donor alleles: the two readings are both exercised and the DOC one is the default
    sp, _ = synth.generate(n_cells=8000, n_donors=3, seed=23)
    a = sp.arrays
    n = sp.n_info
    seen = 0
    for k in range(n - 1):
        if a["len"][0][k] != a["len"][0][k + 1] or a["len"][1][k] != a["len"][1][k + 1]:
            continue
        dif = any(a["vs_seq"][c][k] != a["vs_seq"][c][k + 1] or a["js_seq"][c][k] != a["js_seq"][c][k + 1] for c in range(2))
        if not dif:
            continue
        want = 0
        for c in range(2):
            for fld, left in (("vs_seq", True), ("js_seq", False)):
                i1, i2 = int(a[fld][c][k]), int(a[fld][c][k + 1])
                if i1 == i2:
                    continue
                r1 = a["refseq_arena"][int(a["refseq_off"][i1]):int(a["refseq_off"][i1]) + int(a["refseq_len"][i1])]
                r2 = a["refseq_arena"][int(a["refseq_off"][i2]):int(a["refseq_off"][i2]) + int(a["refseq_len"][i2])]
                m = min(len(r1), len(r2))
                want += int((r1[:m] != r2[:m]).sum()) if left else int((r1[len(r1) - m:] != r2[len(r2) - m:]).sum())
                want += abs(len(r1) - len(r2))
        assert O.ref_positions_differing(sp.struct, k, k + 1) == want
        seen += 1
        if seen > 200:
            break
    assert seen > 0
    e_doc, _, _ = O.join_exacts(sp.struct, default_params(mix_donors=1), threads=4)
    e_doc0, _, _ = O.join_exacts(sp.struct, default_params(mix_donors=1, degradation_mode=0), threads=4)
    assert np.array_equal(e_doc, e_doc0)
