"""JoinPacks built from the reference's own small datasets (enclone_b200/ingest.py; fixtures committed as
tests/golden/flaky_packs.npz by tests/golden/make_flaky_packs.py): real contig sequences, CDR3s, indels and chain
multiplicities through the oracle and through the C ABI."""
import os

import numpy as np
import pytest

import oracle_lib as O
from enclone_b200._abi import JoinPack, default_params
from kat import load as load_kat

HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["flaky", "flaky2", "flaky3", "flaky4", "flaky5", "flaky10", "flaky8ab"]
CODON = {a + b + c: aa for (a, b, c), aa in zip(((a, b, c) for a in "TCAG" for b in "TCAG" for c in "TCAG"),
                                               "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG")}


def fixture_pack(name):
    z = np.load(os.path.join(HERE, "golden", "flaky_packs.npz"))
    arrays = {}
    for key in z.files:
        ds, f = key.split("/", 1)
        if ds != name or f == "is_bcr":
            continue
        if f.endswith("__0") or f.endswith("__1"):
            arrays.setdefault(f[:-3], [None, None])[int(f[-1])] = z[key]
        else:
            arrays[f] = z[key]
    return JoinPack(arrays, is_bcr=bool(int(z[f"{name}/is_bcr"][0])))


def test_loader_reproduces_the_committed_fixtures():
    ref = "/root/reference/enclone_exec/testx/inputs"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is only present in the build container")
    from enclone_b200.ingest import load_datasets
    pack, info = load_datasets([f"{ref}/flaky2/outs/all_contig_annotations.json"])
    want = fixture_pack("flaky2")
    for k, v in pack.arrays.items():
        w = want.arrays[k]
        assert all(np.array_equal(x, y) for x, y in zip(v, w)) if isinstance(v, list) else np.array_equal(v, w), k


def test_exact_subclonotypes_match_the_reference_golden_counts():
    """Regression test 245 (enclone_exec/testx/inputs/outputs/enclone_test245_output): the clonotype with heavy CDR3
    CARPRGYCSGGSCFPFASW consists of an exact subclonotype of 21 cells with two chains and one of 15 cells with the heavy chain only."""
    p = fixture_pack("flaky2")
    a = p.arrays
    seen = {}
    for k in range(p.n_info):
        nt = a["cdr3_arena"][int(a["cdr3_off"][0][k]):int(a["cdr3_off"][0][k]) + int(a["cdr3_len"][0][k])].tobytes().decode()
        aa = "".join(CODON[nt[i:i + 3]] for i in range(0, len(nt), 3))
        if aa == "CARPRGYCSGGSCFPFASW":
            e = int(a["exact_id"][k])
            seen[e] = (int(a["ncells"][e]), int(a["nchains"][e]))
    # (the 3-chain exact subclonotypes that also carry this CDR3 are cell doublets, which enclone removes in a later, out-of-scope filter)
    assert (15, 1) in seen.values() and (21, 2) in seen.values()
    assert sorted(v for v in seen.values() if v[1] < 3) == [(15, 1), (21, 2)]


def test_the_known_answer_pair_comes_out_of_the_loader():
    p = fixture_pack("flaky8ab")
    kat = load_kat()
    assert p.n_info == 2
    got = sorted((p.tig(0, k).decode(), p.tig(1, k).decode()) for k in range(2))
    want = sorted((kat[c]["heavy"]["seq"], kat[c]["light"]["seq"]) for c in ("k1", "k2"))
    assert got == want
    ok, e = O.join_one(p.struct, default_params(mix_donors=1), 0, 1)
    # real contigs, consensus germline: the CDR3 and total differences are those of the golden (cd = 2; 16 + 8 differing positions)
    assert (e.cd, e.diffs) == (2, 24) or not ok
    edges, cls, _ = O.join_exacts(p.struct, default_params(mix_donors=1))
    assert all(x["cd"] == 2 and x["diffs"] == 24 for x in edges)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_runs_on_real_datasets(name):
    p = fixture_pack(name)
    edges, cls, st = O.join_exacts(p.struct, default_params(mix_donors=1), threads=2)
    assert len(cls) == p.n_info
    key = edges["k1"].astype(np.int64) << 32 | edges["k2"]
    assert (np.diff(key) > 0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_equals_oracle_on_real_datasets(name):
    from enclone_b200.join import CompactPack, Joiner
    from parity import assert_same_join
    p = fixture_pack(name)
    for over in (dict(mix_donors=1), dict(mix_donors=1, max_score=1e9, join_cdr3_ident=50.0), dict()):
        params = default_params(**over)
        oe, oc, _ = O.join_exacts(p.struct, params, threads=2)
        j = Joiner(params)
        assert_same_join(j.run(p.struct), oe, oc, max_score=over.get("max_score", 100000.0), what=f"{name} {over}")
        cp = CompactPack(p)
        assert_same_join(j.run(cp.struct), oe, oc, max_score=over.get("max_score", 100000.0), what=f"{name} compact {over}")
        cp.close()
        j.close()
