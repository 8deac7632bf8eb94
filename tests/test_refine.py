"""Post-join refinement: the oracle against answers derived by hand from the documentation
(/root/reference/pages/heuristics.html.src:19-41, pages/default_filters.html.src:292-353) and against a reconstruction of
the clonotype the reference's doublet-filter goldens print (enclone_test170_output / enclone_test171_output); the GPU path
(erefine_run through the C ABI) against the oracle."""
import zlib

import numpy as np
import pytest

import oracle_lib as O
from enclone_b200 import refine as R

NONE = R.NONE


def _seq(tag, n=60):
    rng = np.random.default_rng(zlib.crc32(tag.encode()))
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def _chain(kind, name, cdr3_aa=None, seq=None):
    s = seq if seq is not None else _seq(name)
    return dict(seq=s, cdr3_dna="cdr3-" + (cdr3_aa or name), cdr3_aa=cdr3_aa or name, chain_type=kind)


def _two(h, l, n):
    return dict(ncells=n, chains=[h, l], entries=[(0, 1)])


def golden_170():
    """The clonotype of enclone_test170_output (BCR=123085 CDR3=CAREGGVGVVTATDWYFDLW NDOUBLET): exact subclonotypes 1..4 with
    n = 1, 5, 1, 1 cells; 1 has the heavy chain and both kappa chains, 2 and 3 heavy + the first kappa (3 with a heavy CDR3 one
    amino acid off), 4 the first kappa alone.  enclone_test171_output (the doublet filter on) prints 2, 3, 4 only, in two
    columns.  The goldens do not print the clonotype the second kappa chain comes from; `partner` cells of it are added here."""
    H = _chain("IGH", "H", "CAREGGVGVVTATDWYFDLW")
    H3 = _chain("IGH", "H3", "CAREGGVGVVTTTDWYFDLW")
    K1 = _chain("IGK", "K1", "CQQRSNWPPLFTF")
    K2 = _chain("IGK", "K2", "CQQRSNWPPSITF")

    def build(partner):
        exacts = [dict(ncells=1, chains=[dict(H), dict(K1), dict(K2)], entries=[(0, 1), (0, 2)]),
                  _two(dict(H), dict(K1), 5), _two(H3, dict(K1), 1),
                  dict(ncells=1, chains=[dict(K1)], entries=[(0, None)]),
                  _two(_chain("IGH", "HB", "CARDLLGYW"), dict(K2), partner)]
        joins = [((0, 0), (1, 0)), ((1, 0), (2, 0))]
        links = [((1, 0), (3, 0))]
        return R.pack_from_exacts(exacts, joins, links)
    return build


def test_oracle_reproduces_the_doublet_goldens():
    build = golden_170()
    pk = build(partner=5)
    # test 170: NDOUBLET.  Eight cells, three columns (heavy, kappa 1, kappa 2 in that order), nothing deleted: the third column
    # has one cell and 8 * 1 < 8 is false, so the weak-chain filter keeps it, as the golden shows
    res, st = O.refine(pk, O.refine_opts(doublet=0))
    assert list(res.fate) == [0, 0, 0, 0, 0]
    assert list(res.clonotype_of) == [0, 0, 0, 0, 4]
    assert list(res.n_columns) == [3, 3, 3, 3, 2]
    assert list(res.column_of) == [0, 1, 2, 0, 1, 0, 1, 1, 0, 1]
    # test 171: the doublet filter on.  5 * 1 <= min(6, 5): exact subclonotype 1 goes; seven cells in two columns remain
    res, st = O.refine(pk, O.refine_opts())
    assert list(res.fate) == [R.DOUBLET, 0, 0, 0, 0] and st["n_doublet"] == 1
    assert list(res.clonotype_of) == [NONE, 1, 1, 1, 4]
    assert list(res.n_columns) == [0, 2, 2, 2, 2]
    assert list(res.column_of) == [NONE, NONE, NONE, 0, 1, 0, 1, 1, 0, 1]
    # with four cells in the partner clonotype 5 * 1 <= 4 fails and the doublet stays
    res, _ = O.refine(build(partner=4), O.refine_opts())
    assert list(res.fate) == [0, 0, 0, 0, 0]


def test_oracle_doublet_needs_unrelated_partners():
    # p1 and p2 share a CDR3 with each other: no triple
    H, L, L2 = _chain("IGH", "H"), _chain("IGK", "L"), _chain("IGK", "L2")
    exacts = [dict(ncells=1, chains=[dict(H), dict(L), dict(L2)], entries=[(0, 1), (0, 2)]), _two(dict(H), dict(L), 9),
              _two(dict(H), dict(L2), 9)]
    res, _ = O.refine(R.pack_from_exacts(exacts, [((0, 0), (1, 0))]), O.refine_opts(signature=0, weak_chains=0))
    assert list(res.fate) == [0, 0, 0]


def test_oracle_signature_filter():
    # p = a one-cell three-chain exact subclonotype glued to two two-chain families of 12 and 9 cells: 21 >= 20 * 1
    H, L, HB, LB = _chain("IGH", "H"), _chain("IGK", "L"), _chain("IGH", "HB"), _chain("IGK", "LB")
    def build(n2):
        exacts = [_two(dict(H), dict(L), 12), _two(dict(HB), dict(LB), n2),
                  dict(ncells=1, chains=[dict(H), dict(L), dict(LB)], entries=[(0, 1), (0, 2)])]
        return R.pack_from_exacts(exacts, [((0, 0), (2, 0)), ((1, 0), (2, 1))])
    res, st = O.refine(build(9), O.refine_opts(doublet=0, weak_chains=0))
    assert list(res.fate) == [0, 0, R.SIGNATURE]
    assert list(res.clonotype_of) == [0, 1, NONE]             # and the two families come apart
    assert list(res.n_columns) == [2, 2, 0]
    res, st = O.refine(build(7), O.refine_opts(doublet=0, weak_chains=0))      # 19 < 20
    assert list(res.fate) == [0, 0, 0] and list(res.clonotype_of) == [0, 0, 0]
    # the raw join through the glue aligned H with HB (one column) and LB with LB: columns H|HB, L, LB
    assert list(res.n_columns) == [3, 3, 3]


def test_oracle_weak_chain_filter():
    H, L, LW = _chain("IGH", "H"), _chain("IGK", "L"), _chain("IGK", "LW")
    def build(n_main, n_weak):
        exacts = [_two(dict(H), dict(L), n_main), dict(ncells=n_weak, chains=[dict(H), dict(L), dict(LW)], entries=[(0, 1), (0, 2)])]
        return R.pack_from_exacts(exacts, [((0, 0), (1, 0))])
    res, _ = O.refine(build(15, 2), O.refine_opts(doublet=0, signature=0))      # 8 * 2 = 16 < 17
    assert list(res.fate) == [0, R.WEAK_CHAIN] and list(res.n_columns) == [2, 0]
    res, _ = O.refine(build(14, 2), O.refine_opts(doublet=0, signature=0))      # 16 < 16 is false
    assert list(res.fate) == [0, 0]
    res, _ = O.refine(build(400, 21), O.refine_opts(doublet=0, signature=0))    # more than 20 cells support the chain
    assert list(res.fate) == [0, 0]
    res, _ = O.refine(build(400, 21), O.refine_opts(doublet=0, signature=0, weak_max_cells=21))
    assert list(res.fate) == [0, R.WEAK_CHAIN]


def test_oracle_third_chain_rule_and_multiplicity():
    base = _seq("third", 300)
    def mutated(k):
        s = list(base)
        for i in range(k):
            s[7 * i + 3] = "ACGT"[("ACGT".index(s[7 * i + 3]) + 1) % 4]
        return "".join(s)
    H, L = _chain("IGH", "H"), _chain("IGK", "L")
    def build(k):
        a = dict(ncells=2, chains=[dict(H), dict(L), _chain("IGK", "T1", "ZTHIRD", seq=base)], entries=[(0, 1), (0, 2)])
        b = dict(ncells=2, chains=[dict(H), dict(L), _chain("IGK", "T2", "ZTHIRD", seq=mutated(k))], entries=[(0, 1), (0, 2)])
        return R.pack_from_exacts([a, b], [((0, 0), (1, 0))])
    res, _ = O.refine(build(10), O.refine_opts(doublet=0, signature=0, weak_chains=0))
    assert list(res.n_columns) == [3, 3] and list(res.column_of) == [0, 1, 2, 0, 1, 2]
    res, _ = O.refine(build(11), O.refine_opts(doublet=0, signature=0, weak_chains=0))
    assert list(res.n_columns) == [4, 4] and sorted(res.column_of[[2, 5]]) == [2, 3]
    # two chains of one exact subclonotype in one orbit (both identical to a onesie's chain) make two columns of it
    K = _chain("IGK", "K")
    exacts = [dict(ncells=1, chains=[dict(H), dict(K), dict(K)], entries=[(0, 1), (0, 2)]), dict(ncells=1, chains=[dict(K)], entries=[(0, None)])]
    res, _ = O.refine(R.pack_from_exacts(exacts, [], [((0, 0), (1, 0))]), O.refine_opts(doublet=0, signature=0, weak_chains=0))
    assert list(res.column_of) == [0, 1, 2, 1] and list(res.n_columns) == [3, 3]


def test_oracle_onesie_without_a_match_stands_alone():
    H, L = _chain("IGH", "H"), _chain("IGK", "L")
    exacts = [_two(dict(H), dict(L), 3), dict(ncells=1, chains=[_chain("IGK", "other")], entries=[(0, None)])]
    res, _ = O.refine(R.pack_from_exacts(exacts, [], [((0, 0), (1, 0))]), O.refine_opts())
    assert list(res.clonotype_of) == [0, 1]            # no shared column: the break-up takes it out
    assert list(res.n_columns) == [2, 1]


def test_oracle_on_synthetic_repertoire_is_consistent():
    pk = R.synth_refine_pack(400, seed=11)
    res, st = O.refine(pk, O.refine_opts())
    a = pk.a
    kept = res.fate == 0
    has_entry = np.zeros(pk.n_exact, bool); has_entry[a["exact_id"]] = True
    assert np.all(res.clonotype_of[kept & has_entry] != NONE) and np.all(res.clonotype_of[~kept] == NONE)
    assert st["n_doublet"] > 0 and st["n_signature"] > 0 and st["n_weak"] > 0
    # every kept chain has a column inside its clonotype's table, distinct within its exact subclonotype
    for x in np.nonzero(kept & has_entry)[0][:500]:
        cols = res.column_of[a["chain_off"][x]:a["chain_off"][x + 1]]
        assert len(set(cols)) == len(cols) and cols.max() < res.n_columns[x]
    # raw joins between kept exact subclonotypes keep them in one clonotype and in the same columns
    x1, x2 = a["exact_id"][a["join_k1"]], a["exact_id"][a["join_k2"]]
    both = kept[x1] & kept[x2]
    assert np.array_equal(res.clonotype_of[x1[both]], res.clonotype_of[x2[both]])
    c1 = res.column_of[a["chain_off"][x1] + a["exact_cols0"][a["join_k1"]]]
    c2 = res.column_of[a["chain_off"][x2] + a["exact_cols0"][a["join_k2"]]]
    assert np.array_equal(c1[both], c2[both])
    # running it again on its own output changes nothing more: filters off, same clonotypes and columns
    res2, _ = O.refine(pk, O.refine_opts(doublet=0, signature=0, weak_chains=0))
    assert np.all(res2.fate == 0)


def test_tiled_pack_is_independent_copies():
    pk = R.synth_refine_pack(60, seed=5)
    res, st = O.refine(pk, O.refine_opts())
    res3, st3 = O.refine(R.tile_pack(pk, 3), O.refine_opts())
    assert st3["n_doublet"] == 3 * st["n_doublet"] and st3["n_clonotypes_out"] == 3 * st["n_clonotypes_out"]
    assert np.array_equal(res3.fate, np.tile(res.fate, 3)) and np.array_equal(res3.column_of, np.tile(res.column_of, 3))


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="-3"):
        R.refine_clonotypes(R.synth_refine_pack(5, seed=1))


# ------------------------------------------------------------------------------------------------------------------
# GPU path (erefine_run through the C ABI) against the oracle
# ------------------------------------------------------------------------------------------------------------------
def _assert_same(pk, **kw):
    got, gst = R.refine_clonotypes(pk, **kw)
    want, wst = O.refine(pk, O.refine_opts(**kw))
    for k in ("fate", "clonotype_of", "n_columns", "column_of"):
        g, w = getattr(got, k), getattr(want, k)
        assert np.array_equal(g, w), (k, kw, np.nonzero(g != w)[0][:10], g[g != w][:10], w[g != w][:10])
    for k in ("n_clonotypes_in", "n_clonotypes_out", "n_pure", "n_doublet", "n_signature", "n_weak"):
        assert gst[k] == wst[k], (k, gst[k], wst[k])
    assert gst["gpu_launches"] > 0
    return got, gst


@pytest.mark.gpu
def test_gpu_reproduces_the_hand_cases():
    build = golden_170()
    for partner in (4, 5):
        _assert_same(build(partner))
        _assert_same(build(partner), doublet=False)
    got, _ = R.refine_clonotypes(build(5))
    assert list(got.fate) == [R.DOUBLET, 0, 0, 0, 0] and list(got.n_columns) == [0, 2, 2, 2, 2]


@pytest.mark.gpu
@pytest.mark.parametrize("seed,nfam", [(1, 50), (2, 400), (3, 1500), (4, 3000)])
def test_gpu_equals_oracle_on_synthetic_repertoires(seed, nfam):
    pk = R.synth_refine_pack(nfam, seed=seed)
    _assert_same(pk)
    _assert_same(pk, doublet=False)
    _assert_same(pk, signature=False, weak_chains=False)
    _assert_same(pk, doublet=False, signature=False, weak_chains=False)
    _assert_same(pk, doublet_mult=1, signature_mult=3, weak_max_cells=3, weak_mult=2, third_chain_diffs=30)


@pytest.mark.gpu
def test_gpu_equals_oracle_at_size_and_on_edge_cases():
    pk = R.tile_pack(R.synth_refine_pack(2500, seed=9, frac_doublet=0.08, frac_three=0.1), 12)
    got, st = _assert_same(pk)
    assert st["n_doublet"] > 100 and st["n_clonotypes_out"] > st["n_clonotypes_in"]
    # no raw joins at all; an exact subclonotype without an info entry; a single exact subclonotype
    H, L = _chain("IGH", "H"), _chain("IGK", "L")
    _assert_same(R.pack_from_exacts([_two(dict(H), dict(L), 1), _two(_chain("IGH", "H2"), dict(L), 2)], []))
    _assert_same(R.pack_from_exacts([_two(dict(H), dict(L), 1), dict(ncells=3, chains=[dict(H)], entries=[])], []))
    _assert_same(R.pack_from_exacts([_two(dict(H), dict(L), 1)], []))


@pytest.mark.gpu
def test_gpu_rejects_malformed_packs():
    pk = R.synth_refine_pack(20, seed=3)
    bad = dict(pk.a); bad["exact_id"] = pk.a["exact_id"].copy(); bad["exact_id"][3] = pk.n_exact + 7
    with pytest.raises(RuntimeError, match="-1"):
        R.refine_clonotypes(R.RefinePack(bad))
    bad = dict(pk.a); bad["join_k2"] = pk.a["join_k2"].copy(); bad["join_k2"][0] = pk.n_info
    with pytest.raises(RuntimeError, match="-1"):
        R.refine_clonotypes(R.RefinePack(bad))
    nine = dict(ncells=1, chains=[_chain("IGK", f"c{i}") for i in range(9)], entries=[(0, 1)])
    with pytest.raises(RuntimeError, match="-1"):
        R.refine_clonotypes(R.pack_from_exacts([nine], []))
