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


# ------------------------------------------------------------------------------------------------------------------
# A second, independent restatement in plain Python — written clonotype by clonotype, the way the reference works (a
# matrix per clonotype, literal enumeration of the doublet triples) — against the flat C oracle.
# ------------------------------------------------------------------------------------------------------------------
def _py_refine(exacts, joins, links, doublet=True, signature=True, weak=True, third=10, dmult=5, smult=20, wmax=20, wmult=8):
    import itertools
    nx = len(exacts)
    alive = [bool(e["entries"]) for e in exacts]
    fate = [0] * nx
    # clonotypes the join left: components of joins + links over exact subclonotypes
    lab = list(range(nx))

    def find(p, a):
        while p[a] != a:
            a = p[a]
        return a
    for (x1, _), (x2, _) in list(joins) + list(links):
        a, b = find(lab, x1), find(lab, x2)
        lab[max(a, b)] = min(a, b)
    label = [find(lab, x) if alive[x] else None for x in range(nx)]
    okey = lambda ch: ((("TRX" + ch["chain_type"][3:]) if ch["chain_type"].startswith("TRB") else ("TRY" + ch["chain_type"][3:]) if ch["chain_type"].startswith("TRA")
                        else ch["chain_type"]) + ":" + ch["cdr3_aa"], len(ch["seq"]))
    origins = sorted({tuple(e.get("origin", (0,))) for e in exacts})

    def stage():
        """per clonotype: columns of every chain (dict (x, m) -> col), new labels, number of columns"""
        col, ncols, newlab = {}, {}, {}
        for L in sorted({l for l in label if l is not None}):
            members = sorted((x for x in range(nx) if alive[x] and label[x] == L), key=lambda x: (origins.index(tuple(exacts[x].get("origin", (0,)))), x))
            chains = [(x, m) for x in members for m in range(len(exacts[x]["chains"]))]
            par = {c: c for c in chains}

            def fnd(c):
                while par[c] != c:
                    c = par[c]
                return c

            def join(a, b):
                a, b = fnd(a), fnd(b)
                if a != b:
                    par[max(a, b, key=chains.index)] = min(a, b, key=chains.index)
            for (x1, e1), (x2, e2) in joins:
                if not (alive[x1] and alive[x2] and label[x1] == L):
                    continue
                c1, c2 = exacts[x1]["entries"][e1], exacts[x2]["entries"][e2]
                for h in range(2):
                    if c1[h] is not None and c2[h] is not None:
                        join((x1, c1[h]), (x2, c2[h]))
                if len(exacts[x1]["chains"]) == 3 and len(exacts[x2]["chains"]) == 3 and None not in c1 and None not in c2 and c1[0] != c1[1] and c2[0] != c2[1]:
                    t1, t2 = 3 - c1[0] - c1[1], 3 - c2[0] - c2[1]
                    a, b = exacts[x1]["chains"][t1], exacts[x2]["chains"][t2]
                    heavy = lambda ch: ch["chain_type"][:3] in ("IGH", "TRB")
                    if heavy(a) == heavy(b) and len(a["seq"]) == len(b["seq"]) and sum(p != q for p, q in zip(a["seq"], b["seq"])) <= third:
                        join((x1, t1), (x2, t2))
            for c in chains:                                  # onesies: identical V..J
                if len(exacts[c[0]]["chains"]) == 1:
                    for d in chains:
                        if d != c and exacts[d[0]]["chains"][d[1]]["seq"] == exacts[c[0]]["chains"][0]["seq"]:
                            join(c, d)
            # the break-up: exact subclonotypes sharing a column
            xp = {x: x for x in members}

            def xf(x):
                while xp[x] != x:
                    x = xp[x]
                return x
            for c in chains:
                a, b = xf(c[0]), xf(fnd(c)[0])
                if a != b:
                    xp[max(a, b)] = min(a, b)
            for comp in sorted({xf(x) for x in members}):
                mem = [x for x in members if xf(x) == comp]
                cch = [c for c in chains if c[0] in mem]
                srt = sorted(cch, key=lambda c: (okey(exacts[c[0]]["chains"][c[1]]), members.index(c[0]), c[1]))
                orbits = {}
                for c in cch:
                    orbits.setdefault(fnd(c), []).append(c)
                order = sorted(orbits, key=lambda r: srt.index(min(orbits[r], key=cch.index)))
                base, run = {}, 0
                for r in order:
                    base[r] = run
                    run += max(sum(1 for c in orbits[r] if c[0] == x) for x in mem)
                for x in mem:
                    newlab[x] = comp; ncols[x] = run
                    for m in range(len(exacts[x]["chains"])):
                        r = fnd((x, m))
                        col[(x, m)] = base[r] + sum(1 for mm in range(m) if fnd((x, mm)) == r)
        return col, ncols, newlab

    def pures(col):
        groups = {}
        for x in range(nx):
            if alive[x]:
                groups.setdefault((label[x], tuple(sorted(col[(x, m)] for m in range(len(exacts[x]["chains"]))))), []).append(x)
        return groups

    def apply(dead, code):
        for x in dead:
            alive[x] = False; fate[x] = code
    col, ncols, newlab = stage()
    for x in range(nx):
        label[x] = newlab.get(x) if alive[x] else None
    if doublet:
        P = list(pures(col).values())
        cells = [sum(exacts[x]["ncells"] for x in p) for p in P]
        cd = [{ch["cdr3_dna"] for x in p for ch in exacts[x]["chains"]} for p in P]
        dead = []
        for i0 in range(len(P)):
            part = [j for j in range(len(P)) if j != i0 and cd[j] & cd[i0] and dmult * cells[i0] <= cells[j]]
            if any(not (cd[a] & cd[b]) for a, b in itertools.combinations(part, 2)):
                dead += P[i0]
        if dead:
            apply(dead, 1); col, ncols, newlab = stage()
            for x in range(nx):
                label[x] = newlab.get(x) if alive[x] else None
    if signature:
        G = pures(col)
        dead = []
        for (L, sig), p in G.items():
            if len(sig) < 2:
                continue
            tot = sum(sum(exacts[x]["ncells"] for x in q) for (L2, s2), q in G.items() if L2 == L and s2 != sig and len(s2) == 2 and set(s2) & set(sig))
            if tot >= smult * sum(exacts[x]["ncells"] for x in p):
                dead += p
        if dead:
            apply(dead, 2); col, ncols, newlab = stage()
            for x in range(nx):
                label[x] = newlab.get(x) if alive[x] else None
    if weak:
        dead = set()
        for L in {l for l in label if l is not None}:
            mem = [x for x in range(nx) if alive[x] and label[x] == L]
            if ncols[mem[0]] < 3:
                continue
            total = sum(exacts[x]["ncells"] for x in mem)
            for c in range(ncols[mem[0]]):
                sup = [x for x in mem if any(col[(x, m)] == c for m in range(len(exacts[x]["chains"])))]
                w = sum(exacts[x]["ncells"] for x in sup)
                if w <= wmax and wmult * w < total:
                    dead |= set(sup)
        if dead:
            apply(sorted(dead), 3); col, ncols, newlab = stage()
            for x in range(nx):
                label[x] = newlab.get(x) if alive[x] else None
    clon = [label[x] if alive[x] else NONE for x in range(nx)]
    cols = [col[(x, m)] if alive[x] else NONE for x in range(nx) for m in range(len(exacts[x]["chains"]))]
    return fate, clon, cols, [ncols[x] if alive[x] else 0 for x in range(nx)]


@pytest.mark.parametrize("seed", range(8))
def test_oracle_matches_an_independent_python_restatement(seed):
    kw = dict(frac_doublet=0.2, frac_three=0.3, frac_onesie=0.3, frac_weak=0.1, frac_sig=0.1, big=0.3)
    exacts, joins, links = R.synth_exacts(14, seed=100 + seed, **kw)
    pk = R.pack_from_exacts(exacts, joins, links)
    for opts in (dict(), dict(doublet=0), dict(signature=0, weak_chains=0), dict(doublet_mult=1, signature_mult=2, weak_max_cells=6, weak_mult=2, third_chain_diffs=20)):
        res, _ = O.refine(pk, O.refine_opts(**opts))
        fate, clon, cols, ncols = _py_refine(exacts, joins, links, doublet=opts.get("doublet", 1), signature=opts.get("signature", 1), weak=opts.get("weak_chains", 1),
                                             third=opts.get("third_chain_diffs", 10), dmult=opts.get("doublet_mult", 5), smult=opts.get("signature_mult", 20),
                                             wmax=opts.get("weak_max_cells", 20), wmult=opts.get("weak_mult", 8))
        assert list(res.fate) == fate, (seed, opts)
        assert list(res.clonotype_of) == clon, (seed, opts)
        assert list(res.n_columns) == ncols, (seed, opts)
        assert list(res.column_of) == cols, (seed, opts)
