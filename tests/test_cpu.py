"""CPU-only tests: ABI surface, host logic, the product's p1 function against the oracle and
the golden vector, the generator, and the multi-rank host path on gloo."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O
from enclone_b200 import _lib, synth
from enclone_b200._abi import EDGE_DTYPE, EjoinEdge, EjoinParams, default_params
from enclone_b200.join import EquivRel, p1, plan_shards

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "enclone_join.h")).read()
    import re
    declared = set(re.findall(r"\b(ejoin_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ejoin_pack", "ejoin_params", "ejoin_edge", "ejoin_stats", "ejoin_result", "ejoin_raw_edge", "ejoin_handle"}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.load()
    for s in _lib.EXPORTS:
        assert hasattr(L, s), s
    ghdr = open(os.path.join(ROOT, "include", "enclone_group.h")).read()
    gdecl = set(re.findall(r"\b(egroup_[a-z0-9_]+)\s*\(", ghdr))
    assert gdecl == {"egroup_opts_default", "egroup_run", "egroup_last_error", "egroup_run_asym", "egroup_asym_free"}
    for s in gdecl:
        assert hasattr(L, s), s
    rhdr = open(os.path.join(ROOT, "include", "enclone_refine.h")).read()
    rdecl = set(re.findall(r"\b(erefine_[a-z0-9_]+)\s*\(", rhdr))
    assert rdecl == {"erefine_opts_default", "erefine_run", "erefine_last_error"}
    for s in rdecl:
        assert hasattr(L, s), s


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from enclone_b200.join import Joiner
    with pytest.raises(_lib.EjoinError) as ei:
        Joiner()
    assert ei.value.code == -3


def test_params_default_matches_python_defaults():
    L = _lib.load()
    p = EjoinParams()
    L.ejoin_params_default(C.byref(p))
    q = default_params()
    for name, _ in EjoinParams._fields_:
        if name != "reserved":
            assert getattr(p, name) == getattr(q, name), name


def test_edge_struct_layout():
    assert C.sizeof(EjoinEdge) == EDGE_DTYPE.itemsize == 88


def test_product_p1_hits_the_golden_vector():
    # enclone_test256_output:55-56
    assert p1(33, 5, 2373) == 1.6600942157683246e-6


def test_product_p1_agrees_with_oracle():
    rng = np.random.default_rng(3)
    worst = 0.0
    for _ in range(400):
        k = int(rng.integers(1, 400)); d = int(rng.integers(0, min(k, 40) + 1)); n = 3 * int(rng.integers(500, 1000))
        a, b = p1(k, d, n), O.p1(k, d, n)
        assert abs(a - b) <= 1e-29 + 4.5e-16 * b, (k, d, n, a, b)
        worst = max(worst, abs(a - b))
    assert p1(10, 0, 2400) == 1.0


def test_generator_is_deterministic_and_sorted():
    a, ta = synth.generate(n_cells=3000, n_donors=2, seed=9)
    b, tb = synth.generate(n_cells=3000, n_donors=2, seed=9)
    assert a.n_info == b.n_info and np.array_equal(ta, tb)
    assert np.array_equal(a.arrays["tig_arena"], b.arrays["tig_arena"])
    l0, l1 = a.arrays["len"][0].astype(np.int64), a.arrays["len"][1].astype(np.int64)
    key = l0 * 65536 + l1
    assert (np.diff(key) >= 0).all()           # lens is the first sort key
    bs = a.bucket_starts()
    for i, j in list(zip(bs[:-1], bs[1:]))[:200]:
        tigs = [a.tig(0, k) for k in range(i, j)]
        assert tigs == sorted(tigs)


def test_plan_shards_deals_balanced_chunks():
    pack, _ = synth.generate(n_cells=30000, n_donors=2, seed=4)
    l0, l1 = pack.arrays["len"][0].astype(np.int64), pack.arrays["len"][1].astype(np.int64)
    true_starts = np.concatenate([[0], np.flatnonzero(np.diff(l0 * 65536 + l1) != 0) + 1, [pack.n_info]])
    for world in (1, 2, 4, 8):
        owner, bstart = plan_shards(pack.struct, world)
        assert len(owner) == len(bstart) - 1 and bstart[0] == 0 and bstart[-1] == pack.n_info
        assert np.array_equal(bstart, true_starts)                 # the galloping scan finds every run of equal lens
        assert owner.max() <= world - 1
        # each rank owns at most SHARD_CHUNKS (4) runs of consecutive buckets, dealt in turn
        runs = 1 + int((np.diff(owner.astype(np.int64)) != 0).sum())
        assert runs <= 4 * world
        sz = np.diff(bstart).astype(np.float64)
        cost = sz * (sz - 1) / 2 + 64 * sz
        per = np.bincount(owner, weights=cost, minlength=world)
        assert per.max() <= 1.6 * cost.sum() / world + cost.max()


def test_equivrel_mirror():
    eq = EquivRel(np.array([0, 0, 2, 0, 2, 5]))
    assert eq.orbit_reps() == [0, 2, 5] and eq.orbit(2) == [2, 4] and eq.class_id(3) == 0 and eq.norbits() == 3


def test_oracle_two_cell_rule_and_skip_semantics():
    pack, _ = synth.generate(n_cells=4000, n_donors=1, seed=12)
    e_strict, c_strict, _ = O.join_exacts(pack.struct, default_params())
    e_easy, c_easy, _ = O.join_exacts(pack.struct, default_params(easy=1))
    assert len(e_easy) >= len(e_strict)
    # emitted edges form a forest: #classes = n - #edges - (same-exact merges)
    n = pack.n_info
    ex = pack.arrays["exact_id"]
    # oracle threads do not change the result
    e8, c8, _ = O.join_exacts(pack.struct, default_params(), threads=8)
    assert np.array_equal(e8, e_strict) and np.array_equal(c8, c_strict)
    # raw_joins are sorted by (k1,k2)
    key = e_strict["k1"].astype(np.int64) << 32 | e_strict["k2"]
    assert (np.diff(key) > 0).all()


def test_rust_ffi_crate_declares_the_same_symbols():
    """rust/enclone_join_ffi (written, not compilable here) must bind exactly the header's entry points."""
    import re
    rs = open(os.path.join(ROOT, "rust", "enclone_join_ffi", "src", "lib.rs")).read()
    bound = set(re.findall(r"pub fn (ejoin_[a-z0-9_]+)\(", rs))
    assert bound == set(_lib.EXPORTS), bound ^ set(_lib.EXPORTS)
    assert set(re.findall(r"pub fn (egroup_[a-z0-9_]+)\(", rs)) == {"egroup_opts_default", "egroup_run", "egroup_last_error", "egroup_run_asym", "egroup_asym_free"}
    assert set(re.findall(r"pub fn (erefine_[a-z0-9_]+)\(", rs)) == {"erefine_opts_default", "erefine_run", "erefine_last_error"}


def test_tier1_work_decomposition_arithmetic(tmp_path):
    """tests/seg_enum_test.cu (host code only, built with nvcc): the (segment, row) enumeration of
    ejoin_dev.cuh is a bijection, item_decode()'s loop inverts it, segments respect their length cap."""
    import shutil
    import subprocess
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "seg_enum_test")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", exe, os.path.join(root, "tests", "seg_enum_test.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_pack_compact_round_trip():
    """ejoin_pack_compact (host code): the 2-bit words decode back to the ASCII bases in the DnaString layout
    (base i of a chain in word i >> 5, bits 63-2(i&31), 62-2(i&31); A,C,G,T = 0..3), chains with a deletion marker or
    a non-ACGT byte stay ASCII, and every record field equals the parallel array it replaces."""
    from enclone_b200.join import CompactPack
    pack, _ = synth.generate(n_cells=3000, n_donors=2, seed=5, frac_indel=0.05)
    a = pack.arrays
    arena = a["tig_arena"]
    rng = np.random.default_rng(1)
    for k in rng.choice(pack.n_info, size=40, replace=False):          # a few non-ACGT bytes outside the CDR3s
        c = int(rng.integers(0, 2))
        cs, cl = int(a["cdr3_start"][c][k]), int(a["cdr3_len"][c][k])
        pos = int(rng.integers(0, max(1, cs)))
        arena[int(a["tig_off"][c][k]) + pos] = ord("N")
    for planes in (True, False):
        if not planes:
          os.environ["EJOIN_COMPACT_DNASTRING"] = "1"
        try:
          cp = CompactPack(pack, threads=3)
        finally:
          os.environ.pop("EJOIN_COMPACT_DNASTRING", None)
        st = cp.struct
        assert st.flags & 2 and st.flags & 4 and st.n_info == pack.n_info and bool(st.flags & 16) == planes
        words = np.ctypeslib.as_array(st.tig2_arena, shape=(int(st.tig2_arena_words),))
        asc = np.ctypeslib.as_array(st.tig_arena, shape=(max(1, int(st.tig_arena_bytes)),))
        n_ascii = 0
        for k in range(0, pack.n_info, 7):
            r = st.entry_rec[k]
            for c in range(2):
                ln = int(a["len"][c][k])
                src = pack.tig(c, k)
                plain = not a["has_del"][c][k] and all(ch in b"ACGT" for ch in src)
                assert bool(r.has_del[c]) == (not plain)
                if plain:
                    w = words[r.tig_off[c]: r.tig_off[c] + (ln + 31) // 32]
                    if planes:          # EJOIN_F_TIGS_PLANES: high-bit plane in bits 63..32, low-bit plane in bits 31..0, base i at bit 31-i
                        got = bytes(b"ACGT"[(((int(w[i >> 5]) >> (63 - (i & 31))) & 1) << 1) | ((int(w[i >> 5]) >> (31 - (i & 31))) & 1)] for i in range(ln))
                    else:
                        got = bytes(b"ACGT"[(int(w[i >> 5]) >> (62 - 2 * (i & 31))) & 3] for i in range(ln))
                else:
                    n_ascii += 1
                    got = asc[r.tig_off[c]: r.tig_off[c] + ln].tobytes()
                assert got == src
                assert r.cdr3_start[c] == int(a["cdr3_start"][c][k])
                for f in ("v_ref_id", "j_ref_id", "vs_seq", "js_seq", "vname_id", "vuni_seq", "dref_id", "utr_seq"):
                    v = int(a[f][c][k])
                    assert getattr(r, f)[c] == (0xFFFF if v == 0xFFFFFFFF else v), f
            assert r.hcomp == int(a["hcomp"][k]) and r.fr3_start == int(a["fr3_start"][k])
        assert n_ascii > 0
        cp.close()


def test_tetrahedron_contraction_is_the_cdr3_hamming_test():
    """The arithmetic k_sweep_mma rests on, checked in numpy: bases as tetrahedron vertices in {+1,-1}^3 give
    dot = 3L - 4d; with the two threshold slots ((16, 1) on the a side, (-thr // 16, -thr % 16) on the b side) the accumulator is
    dot - thr and `d <= cdmax` is exactly a clear sign bit; every value fits int8 and K = 3L + 2 fits the 288-byte row for L <= 95."""
    rng = np.random.default_rng(7)
    V = np.array([[1, 1, 1], [1, -1, -1], [-1, 1, -1], [-1, -1, 1]], np.int32)
    assert np.array_equal(V @ V.T, 4 * np.eye(4, dtype=np.int32) - 1)            # 3 on the diagonal, -1 elsewhere
    for L, cdmax in ((1, 0), (36, 5), (84, 12), (95, 14), (95, 0)):
        K = (3 * L + 2 + 31) // 32 * 32
        assert K <= 288
        thr = 3 * L - 4 * cdmax
        assert 0 < thr and thr // 16 <= 127
        a = rng.integers(0, 4, (64, L))
        b = a[rng.integers(0, 64, 64)].copy()                                     # relatives: a few substitutions
        for r in range(64):
            for p in rng.integers(0, L, rng.integers(0, cdmax + 6)):
                b[r, p] = rng.integers(0, 4)
        ra = np.zeros((64, K), np.int32); rb = np.zeros((64, K), np.int32)
        ra[:, :3 * L] = V[a].reshape(64, 3 * L); rb[:, :3 * L] = V[b].reshape(64, 3 * L)
        ra[:, 3 * L], ra[:, 3 * L + 1] = 16, 1
        rb[:, 3 * L], rb[:, 3 * L + 1] = -(thr // 16), -(thr % 16)
        assert np.abs(ra).max() <= 127 and np.abs(rb).max() <= 127
        acc = rb @ ra.T                                                           # rows = b (the patched B tile), columns = a
        d = (b[:, None, :] != a[None, :, :]).sum(-1)
        assert np.array_equal(acc, 3 * L - 4 * d - thr)
        assert np.array_equal(acc >= 0, d <= cdmax)
