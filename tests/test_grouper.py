"""Symmetric grouping: oracle/grouper_oracle.c against hand-derived answers (CPU), the GPU path against the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib
from enclone_b200 import group as G


def tiny_pack(heavy, light=None, cdr3=(3, 4), names=None):
    """one exact subclonotype per clonotype, chains: heavy[i] (and light[i])"""
    n = len(heavy)
    light = light or [None] * n
    coff, cex, xoff = [0], [], [0]
    left, seqs, starts, aas, col_off, cv, cj, cd, cl = [], [], [], [], [0], [], [], [], []
    for i in range(n):
        cex.append(i)
        for s, is_left in ((heavy[i], 1), (light[i], 0)):
            if s is None:
                continue
            left.append(is_left); seqs.append(s); starts.append(cdr3[0]); aas.append(G.translate(s[cdr3[0]:cdr3[0] + 3 * cdr3[1]]))
            nm = names[i] if names else (0, 0, 0)
            cv.append(nm[0] + (0 if is_left else 100)); cj.append(nm[1]); cd.append(nm[2]); cl.append(is_left)
        xoff.append(len(left)); coff.append(len(cex)); col_off.append(len(cv))
    so, ao, sa, aa = [], [], bytearray(), bytearray()
    for s, a in zip(seqs, aas):
        so.append(len(sa)); sa += s.encode(); ao.append(len(aa)); aa += a.encode()
    return G.GroupPack(dict(clono_exact_off=coff, clono_exact=cex, exact_chain_off=xoff, left=left, seq_off=so, seq_len=[len(s) for s in seqs],
                            seq_del_len=[len(s) for s in seqs], cdr3_start=starts, cdr3_aa_off=ao, cdr3_aa_len=[len(a) for a in aas],
                            seq_arena=np.frombuffer(bytes(sa), np.uint8), aa_arena=np.frombuffer(bytes(aa), np.uint8), col_off=col_off, col_vname=cv,
                            col_jname=cj, col_dname=cd, col_left=cl))


def opts(**kw):
    hf = kw.pop("hf_matrix", None)
    o, keep = G.make_opts(hf_matrix=hf, **kw)
    o._keep = keep
    return o


BASE = "GCTGCTGCAAAACCCGGGTTTGCTGCAGCTGCTGCAGCTGCTGCAGCTGCTGCAGCTGCT"        # 60 nt


def mutated(k):
    s = list(BASE)
    for i in range(k):
        s[5 + 7 * i] = "T" if s[5 + 7 * i] != "T" else "A"
    return "".join(s)


def test_oracle_heavy_pc_threshold():
    # distances 0, 3, 6 substitutions of 60: identities 1.0, 0.95, 0.90 against the first one; 3 vs 6 differ at 3 positions
    pk = tiny_pack([mutated(0), mutated(3), mutated(6)])
    g, ev = oracle_lib.group(pk, opts(heavy_pc=95.0))
    assert g.tolist() == [0, 0, 0]            # 0-1 at 95%, 1-2 at 95% (transitively)
    g, _ = oracle_lib.group(pk, opts(heavy_pc=96.0))
    assert g.tolist() == [0, 1, 2]
    pk = tiny_pack([mutated(0), mutated(6), mutated(0)[:57]])
    g, _ = oracle_lib.group(pk, opts(heavy_pc=95.0))
    assert g.tolist() == [0, 1, 0]            # a 3-base deletion: d = 3, r2 = 54/57 < 0.95 but r1 = 57/60 >= 0.95 (either ratio accepts)


def test_oracle_key_passes():
    names = [(1, 1, 1), (1, 1, 2), (2, 1, 1), (1, 1, 1)]
    pk = tiny_pack([BASE] * 4, [BASE] * 4, names=names)
    assert oracle_lib.group(pk, opts())[0].tolist() == [0, 0, 0, 0]
    assert oracle_lib.group(pk, opts(vj_refname=1))[0].tolist() == [0, 0, 2, 0]
    assert oracle_lib.group(pk, opts(vdj_refname=1))[0].tolist() == [0, 1, 2, 0]
    assert oracle_lib.group(pk, opts(v_heavy_refname=1))[0].tolist() == [0, 0, 2, 0]
    # a clonotype without a heavy chain is alone in a *_heavy_refname pass (grouper.rs:156-161)
    pk = tiny_pack([BASE, None, None], [BASE, BASE, BASE])
    assert oracle_lib.group(pk, opts(v_heavy_refname=1))[0].tolist() == [0, 1, 2]
    assert oracle_lib.group(pk, opts(vj_refname=1))[0].tolist() == [0, 1, 1]


def test_oracle_orbit_idiom():
    """grouper.rs:441-452 lists `ee.orbit(i)` for i < #orbits (not orbit(reps[i])): with classes {0,1} and {2,3} it lists
    {0,1} twice and never {2,3}, so 2 and 3 stay apart in the final partition.  The oracle does what the source does."""
    far = "".join({"A": "C", "C": "A", "G": "T", "T": "G"}[c] for c in BASE)
    pk = tiny_pack([BASE, mutated(1), far, far])
    assert oracle_lib.group(pk, opts(heavy_pc=95.0))[0].tolist() == [0, 0, 2, 3]
    # the same four with the far pair first: classes {0,1} {2,3} again, by symmetry the same outcome
    pk = tiny_pack([far, far, BASE, mutated(1)])
    assert oracle_lib.group(pk, opts(heavy_pc=95.0))[0].tolist() == [0, 0, 2, 3]
    # classes {0,2} and {1,3}: orbit(0) and orbit(1) list both
    pk = tiny_pack([BASE, far, mutated(1), far])
    assert oracle_lib.group(pk, opts(heavy_pc=95.0))[0].tolist() == [0, 1, 0, 1]


def test_oracle_cdr3_passes():
    # CDR3 = 4 codons from position 3: AAA ACC CGG GTT -> KTRV
    a = BASE[:3] + "AAAACCCGGGTT" + BASE[15:]
    b = BASE[:3] + "AAAACCCGGGTA" + BASE[15:]      # one nt differs inside the CDR3 (V -> V? GTA = V): aa identical
    c = BASE[:3] + "AAATCCCGGGTT" + BASE[15:]      # T -> S
    pk = tiny_pack([a, b, c])
    assert G.translate(b[3:15]) == "KTRV" and G.translate(c[3:15]) == "KSRV"
    # cdr3_heavy_pc 90%: d_max = floor(0.1 * 12) = 1 -> all within one substitution of a
    assert oracle_lib.group(pk, opts(cdr3_heavy_pc=90.0))[0].tolist() == [0, 0, 0]
    # 92%: d_max = floor(0.96) = 0 -> nothing joins
    assert oracle_lib.group(pk, opts(cdr3_heavy_pc=92.0))[0].tolist() == [0, 1, 2]
    # amino acids, 75%: d_max = floor(0.25 * 4) = 1
    assert oracle_lib.group(pk, opts(cdr3_aa_heavy_pc=75.0))[0].tolist() == [0, 0, 0]
    assert oracle_lib.group(pk, opts(cdr3_aa_heavy_pc=80.0))[0].tolist() == [0, 0, 2]
    # penalty matrix: T<->S costs 0.4 -> err = 0.1 over 4 residues
    m = np.zeros((20, 20)); i, j = G.HF_ORDER.index("T"), G.HF_ORDER.index("S"); m[i, j] = m[j, i] = 0.4
    assert oracle_lib.group(pk, opts(cdr3_heavy_pc_hf=89.0, hf_matrix=m))[0].tolist() == [0, 0, 0]
    # at 90 the source's own f64 expression decides: err = 0.1 > 1.0 - 0.9 = 0.09999999999999998
    assert oracle_lib.group(pk, opts(cdr3_heavy_pc_hf=90.0, hf_matrix=m))[0].tolist() == [0, 0, 2]
    # whole-chain amino acids: 20 residues, one differs in c -> 0.95
    assert oracle_lib.group(pk, opts(aa_heavy_pc=95.0))[0].tolist() == [0, 0, 0]
    assert oracle_lib.group(pk, opts(aa_heavy_pc=96.0))[0].tolist() == [0, 0, 2]


def test_oracle_synthetic_is_nontrivial():
    pk = G.synth_group_pack(300, seed=3)
    g, ev = oracle_lib.group(pk, opts(vj_refname=1, heavy_pc=96.0))
    ncls = len(set(g.tolist()))
    assert 10 < ncls < 300 and ev > 300


def test_symbols_exported():
    L = C.CDLL(G._lib.SO_PATH)
    for name in ("egroup_opts_default", "egroup_run", "egroup_last_error", "egroup_run_asym", "egroup_asym_free"):
        assert hasattr(L, name)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    pk = tiny_pack([BASE, mutated(1)])
    with pytest.raises(G._lib.EjoinError) as ei:
        G.group_clonotypes(pk, heavy_pc=95.0)
    assert ei.value.code == -3
    with pytest.raises(G._lib.EjoinError) as ei:
        G.group_asymmetric(pk, np.ones(2, np.uint8), top=1)
    assert ei.value.code == -3
    # key partitions alone are host work and need no device
    g, _ = G.group_clonotypes(tiny_pack([BASE] * 3, names=[(1, 1, 1), (2, 1, 1), (1, 1, 1)]), vj_refname=1)
    assert g.tolist() == [0, 1, 0]


OPTION_SETS = [
    dict(heavy_pc=96.0), dict(light_pc=97.0), dict(vj_refname=1, heavy_pc=95.0, light_pc=95.0), dict(aa_heavy_pc=94.0), dict(aa_light_pc=96.0),
    dict(cdr3_heavy_pc=85.0), dict(cdr3_light_pc=90.0), dict(cdr3_aa_heavy_pc=80.0), dict(cdr3_aa_light_pc=85.0), dict(cdr3_heavy_pc_hf=90.0, hf=True),
    dict(vdj_refname=1, cdr3_len=1, cdr3_aa_heavy_pc=70.0), dict(v_heavy_refname=1, vj_len=1), dict(vj_heavy_refname=1, cdr3_heavy_len=1, cdr3_heavy_pc=80.0),
    dict(vdj_heavy_refname=1, cdr3_light_len=1, heavy_pc=90.0, cdr3_aa_light_pc=75.0), dict(heavy_pc=98.0, aa_heavy_pc=90.0, cdr3_heavy_pc=70.0),
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(OPTION_SETS)))
def test_gpu_grouping_equals_oracle(case):
    kw = dict(OPTION_SETS[case])
    hf = G.hf_matrix_example() if kw.pop("hf", False) else None
    heavy = any(k in kw for k in ("heavy_pc", "light_pc", "aa_heavy_pc", "aa_light_pc"))          # whole-chain passes: the oracle's DP is slow
    keyed = any(k in G.KEY_FLAGS for k in kw)
    n = (1200 if keyed else 140) if heavy else 1500
    pk = G.synth_group_pack(n, seed=case, n_families=max(2, n // (4 + case)))
    want, ev = oracle_lib.group(pk, opts(hf_matrix=hf, **kw))
    got, st = G.group_clonotypes(pk, hf_matrix=hf, **kw)
    assert np.array_equal(got, want)
    assert 1 < len(set(want.tolist())) < n
    if any(k.endswith("_pc") or k.endswith("_hf") for k in kw):
        assert st["gpu_launches"] > 0 and st["pairs_evaluated"] > 0


@pytest.mark.gpu
def test_gpu_grouping_hand_cases():
    far = "".join({"A": "C", "C": "A", "G": "T", "T": "G"}[c] for c in BASE)
    pk = tiny_pack([BASE, mutated(1), far, far])
    assert G.group_clonotypes(pk, heavy_pc=95.0)[0].tolist() == [0, 0, 2, 3]
    pk = tiny_pack([mutated(0), mutated(6), mutated(0)[:57]])
    assert G.group_clonotypes(pk, heavy_pc=95.0)[0].tolist() == [0, 1, 0]
    assert G.group_clonotypes(tiny_pack([]), heavy_pc=95.0)[0].tolist() == []
    assert G.group_clonotypes(tiny_pack([BASE]), heavy_pc=95.0)[0].tolist() == [0]


@pytest.mark.gpu
def test_gpu_levenshtein_long_and_ragged():
    """lengths across the 64-row block boundaries of the bit-parallel kernel, with insertions and deletions"""
    rng = np.random.default_rng(5)
    heavy = []
    for n in (1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 256, 300, 383, 384, 385, 448, 511, 512):
        s = "".join("ACGT"[int(x)] for x in rng.integers(0, 4, n))
        heavy.append(s)
        t = list(s)
        for _ in range(max(1, n // 40)):
            p = int(rng.integers(0, len(t)))
            op = int(rng.integers(0, 3))
            if op == 0: t[p] = "ACGT"[int(rng.integers(0, 4))]
            elif op == 1 and len(t) > 1: del t[p]
            else: t.insert(p, "G")
        heavy.append("".join(t)[:512])
    pk = tiny_pack(heavy, cdr3=(0, 0))
    for pc in (90.0, 95.0, 97.0, 98.5, 99.5):
        want, _ = oracle_lib.group(pk, opts(heavy_pc=pc))
        got, _ = G.group_clonotypes(pk, heavy_pc=pc)
        assert np.array_equal(got, want), pc
        want, _ = oracle_lib.group(pk, opts(aa_heavy_pc=pc))
        got, _ = G.group_clonotypes(pk, aa_heavy_pc=pc)
        assert np.array_equal(got, want), pc


@pytest.mark.gpu
def test_gpu_grouping_too_long_is_an_error():
    s = "ACGT" * 200                                   # 800 nt > 512 on both sides
    pk = tiny_pack([s, s[:-1] + "A"], cdr3=(0, 0))
    with pytest.raises(G._lib.EjoinError):
        G.group_clonotypes(pk, heavy_pc=95.0)
    pk = tiny_pack([BASE, BASE[:10] + "R" + BASE[11:]])          # a byte outside ACGTN-
    with pytest.raises(G._lib.EjoinError):
        G.group_clonotypes(pk, heavy_pc=95.0)
    assert G.group_clonotypes(tiny_pack([BASE, BASE[:10] + "N" + BASE[11:], BASE[:10] + "-" + BASE[11:]]), heavy_pc=98.0)[0].tolist() == [0, 0, 0]


def test_kernel_levenshtein_arithmetic_on_host(tmp_path):
    """egroup_myers.h compiles for the host too: fuzz the kernel's bit-parallel distance against the DP (no GPU needed)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "myers_fuzz")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(root, "tests", "myers_fuzz.cpp")])
    out = subprocess.run([exe, "4000"], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout


# ---- pins on real enclone output: tests/golden/gtest_pins.json (made by tests/golden/make_gtest_pins.py from the reference's goldens) ----
def _pin_pack(pin):
    import json  # noqa: F401
    ids = {}

    def nid(s):
        return ids.setdefault(s, len(ids) + 1)

    coff, cex, xoff = [0], [], [0]
    left, seqs, starts, aas, col_off, cv, cj, cd, cl = [], [], [], [], [0], [], [], [], []
    for c in pin["clonotypes"]:
        nm = c["names"]
        hv, rest = nm[0], nm[1:]
        hd = rest.pop(0) if rest[0].startswith("IGHD") else "null"
        hj, lv, lj = rest
        cv += [nid(hv), nid(lv)]; cj += [nid(hj), nid(lj)]; cd += [nid(hd), nid("null")]; cl += [1, 0]
        col_off.append(len(cv))
        for e in c["exacts"]:
            cex.append(len(xoff) - 1)
            for cdr3, is_left in ((e["heavy_cdr3"], 1), (e["light_cdr3"], 0)):
                if cdr3 is None:
                    continue
                s = c["vj_seq1"] if (is_left and c.get("vj_seq1")) else "ACG" * (len(cdr3) + 20)
                left.append(is_left); seqs.append(s); starts.append(30); aas.append(cdr3)
            xoff.append(len(left))
        coff.append(len(cex))
    so, ao, sa, aa = [], [], bytearray(), bytearray()
    for s, a in zip(seqs, aas):
        so.append(len(sa)); sa += s.encode(); ao.append(len(aa)); aa += a.encode()
    return G.GroupPack(dict(clono_exact_off=coff, clono_exact=cex, exact_chain_off=xoff, left=left, seq_off=so, seq_len=[len(s) for s in seqs],
                            seq_del_len=[len(s) for s in seqs], cdr3_start=starts, cdr3_aa_off=ao, cdr3_aa_len=[len(a) for a in aas],
                            seq_arena=np.frombuffer(bytes(sa), np.uint8), aa_arena=np.frombuffer(bytes(aa), np.uint8), col_off=col_off, col_vname=cv,
                            col_jname=cj, col_dname=cd, col_left=cl))


def _pins():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gtest_pins.json")))


# options under which the pinned clonotypes must come apart again — derived from the golden's own data, not printed by enclone
SPLITS = {"gtest17": dict(vj_refname=1, heavy_pc=96.7),                    # distance 15 of 448: identity 0.96652
          "gtest5": dict(vj_refname=1, cdr3_aa_heavy_pc=80.0, cdr3_aa_light_pc=85.0),   # light CDR3s differ at 2 of 12: floor(0.15 * 12) = 1
          "gtest7": dict(cdr3_aa_heavy_pc=100.0, cdr3_aa_light_pc=100.0),
          "gtest19": dict(vj_refname=1, cdr3_light_len=1, cdr3_heavy_len=1)}


def test_oracle_reproduces_the_groups_of_enclones_goldens():
    pins = {k: v for k, v in _pins().items() if "opts" in v}
    assert set(pins) == {"gtest5", "gtest7", "gtest8", "gtest17", "gtest19"}
    for name, pin in pins.items():
        pk = _pin_pack(pin)
        g, _ = oracle_lib.group(pk, opts(**pin["opts"]))
        assert g.tolist() == [0] * pk.n_clono, (name, g.tolist())          # one group, as the golden prints it
        if name in SPLITS:
            g, _ = oracle_lib.group(pk, opts(**SPLITS[name]))
            assert len(set(g.tolist())) == pk.n_clono, (name, g.tolist())
    a, b = (c["vj_seq1"] for c in pins["gtest17"]["clonotypes"])
    assert len(a) == len(b) == 448


@pytest.mark.gpu
def test_gpu_reproduces_the_groups_of_enclones_goldens():
    for name, pin in _pins().items():
        if "opts" not in pin:
            continue
        pk = _pin_pack(pin)
        g, _ = G.group_clonotypes(pk, **pin["opts"])
        assert g.tolist() == [0] * pk.n_clono, (name, g.tolist())
        if name in SPLITS:
            g, _ = G.group_clonotypes(pk, **SPLITS[name])
            assert len(set(g.tolist())) == pk.n_clono, (name, g.tolist())


# ---- Case 2 (asymmetric grouping): distances printed by enclone_gtest10/11/12_output ----
DECOY = dict(names=["IGHV1-2", "IGHJ1", "IGKV1-5", "IGKJ1"], exacts=[dict(heavy_cdr3="CWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWWW", light_cdr3="CMMMMMMMMMMMMMMMMMMMMMMMMMW")], vj_seq1=None)


def _asym_cases():
    pins = _pins()
    p10, p11, p12 = pins["gtest10"], pins["gtest11"], pins["gtest12"]
    assert p12["clonotypes"][0]["exacts"] == p11["clonotypes"][0]["exacts"]          # gtest12 = gtest11's center under max=13
    out = []
    for pin, bound, want in ((p10, p10["asym"], [d["distance"] for d in p10["clonotypes"]]), (p11, p11["asym"], [d["distance"] for d in p11["clonotypes"]]),
                             (p11, p12["asym"], [0.0])):
        full = dict(clonotypes=pin["clonotypes"] + [DECOY, DECOY])
        out.append((_pin_pack(full), bound, want))
    return out


def test_oracle_reproduces_the_distances_of_enclones_agroup_goldens():
    for pk, bound, want in _asym_cases():
        cen = np.zeros(pk.n_clono, np.uint8); cen[0] = 1
        groups, ev = oracle_lib.group_asym(pk, cen, G.asym_opts(**bound))
        assert len(groups) == 1
        assert groups[0] == [(i, d) for i, d in enumerate(want)], groups     # center first, then the members at 13 / 14 in index order


def test_oracle_asym_walk_semantics():
    pk = G.synth_group_pack(80, seed=4)
    cen = np.zeros(80, np.uint8); cen[::7] = 1
    g_all, _ = oracle_lib.group_asym(pk, cen, G.asym_opts())
    assert len(g_all) == int(cen.sum()) and all(len(g) >= 80 - 1 for g in g_all)           # no bound: everybody, the center listed once
    g_top, _ = oracle_lib.group_asym(pk, cen, G.asym_opts(top=5))
    for g in g_top:
        assert 6 <= len(g) <= 7 and g[0][1] == 0.0                                        # positions 0..5 of the sorted row, the center skipped if among them
        assert [d for _, d in g[1:]] == sorted(d for _, d in g[1:])
    g_min, _ = oracle_lib.group_asym(pk, cen, G.asym_opts(max_dist=0.0, min_group=2))
    assert all(len(g) >= 2 and all(d == 0.0 for _, d in g) for g in g_min)


@pytest.mark.gpu
def test_gpu_reproduces_the_distances_of_enclones_agroup_goldens():
    for pk, bound, want in _asym_cases():
        cen = np.zeros(pk.n_clono, np.uint8); cen[0] = 1
        groups, st = G.group_asymmetric(pk, cen, **bound)
        assert groups == [[(i, d) for i, d in enumerate(want)]]


@pytest.mark.gpu
@pytest.mark.parametrize("bound", [dict(top=0), dict(top=3), dict(top=40), dict(top=5000), dict(max_dist=0.0), dict(max_dist=6.5), dict(max_dist=30.0, min_group=3),
                                   dict(max_dist=2e9), dict(max_dist=-1.0), dict(), dict(top=2, min_group=4)])
def test_gpu_asymmetric_grouping_equals_oracle(bound):
    for n, seed, step in ((300, 11, 9), (1200, 12, 1)):
        if step == 1 and not bound:
            continue                                            # no bound with every clonotype a center: n^2 members, pointless
        pk = G.synth_group_pack(n, seed=seed)
        cen = np.zeros(n, np.uint8); cen[::step] = 1
        want, ev = oracle_lib.group_asym(pk, cen, G.asym_opts(**bound))
        got, st = G.group_asymmetric(pk, cen, **bound)
        assert got == want
        if bound.get("max_dist", 0.0) >= 0.0:                   # a negative bound needs no distances at all
            assert st["gpu_launches"] >= 1 and st["pairs_evaluated"] == int(cen.sum()) * n


@pytest.mark.gpu
def test_gpu_asymmetric_edge_cases():
    pk = G.synth_group_pack(50, seed=3)
    assert G.group_asymmetric(pk, np.zeros(50, np.uint8), top=3)[0] == []                 # no center: no group, no device work
    pk0 = tiny_pack([])
    assert G.group_asymmetric(pk0, np.zeros(0, np.uint8), max_dist=3.0)[0] == []
    one = tiny_pack([BASE])
    assert G.group_asymmetric(one, np.ones(1, np.uint8), top=2)[0] == oracle_lib.group_asym(one, np.ones(1, np.uint8), G.asym_opts(top=2))[0]
