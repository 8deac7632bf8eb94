"""GPU parity tests: the CUDA join (through the C ABI) against the CPU oracle on the same
inputs.  Bit-exact for edges, per-edge integers and the partition; p1/score within the
tolerance stated in tests/parity.py."""
import numpy as np
import pytest

import oracle_lib as O
from enclone_b200 import synth
from enclone_b200._abi import JoinPack, default_params
from enclone_b200.join import Joiner, join_exacts, p1
from kat import kat_pack
from parity import assert_same_join

pytestmark = pytest.mark.gpu


def _run_both(pack, **pk):
    params = default_params(**pk)
    g = join_exacts(pack, params)
    oe, oc, ost = O.join_exacts(pack.struct, params, threads=8)
    return g, oe, oc, ost


def test_kat256_through_the_gpu():
    pack, kat = kat_pack()
    e = kat["expected"]
    g, oe, oc, _ = _run_both(pack, mix_donors=1)
    assert_same_join(g, oe, oc, what="kat256")
    assert len(g.edges) == 1
    x = g.edges[0]
    assert (x["shares"][0], x["indeps"][0], x["cd"], x["k"], x["d"], x["n"]) == (5, 23, 2, 33, 5, 2373)
    assert list(x["shares_details"][0]) == e["shares_details"]
    assert x["p1"] == float(e["p1"]) and x["mult"] == float(e["mult"]) and x["score"] == float(e["score"])
    assert x["err"] == 1
    g2 = join_exacts(pack, default_params(mix_donors=0))
    assert len(g2.edges) == 0 and list(g2.class_of) == [0, 1]


def test_kat256_replicated_tile():
    pack, _ = kat_pack(extra_copies=40)
    g, oe, oc, _ = _run_both(pack, mix_donors=1)
    assert_same_join(g, oe, oc, what="kat256x41")


def test_device_p1_table_matches_host_function():
    # the device table is built by the same double-double code as ejoin_p1()
    pack, _ = kat_pack()
    g = join_exacts(pack, default_params(mix_donors=1))
    assert g.edges[0]["p1"] == p1(33, 5, 2373)


@pytest.mark.parametrize("name,scale", [("C1", 1.0), ("C2", 0.2), ("C3", 0.02), ("C5", 0.03)])
def test_synthetic_configs(name, scale):
    pack, _ = synth.generate_config(name, scale)
    g, oe, oc, ost = _run_both(pack)
    assert_same_join(g, oe, oc, what=name)
    assert g.stats["pairs_candidate"] == ost["pairs_candidate"]


def test_giant_bucket_small():
    pack, _ = synth.generate_config("C4", 0.02)      # 4000 entries in one bucket
    g, oe, oc, ost = _run_both(pack)
    assert_same_join(g, oe, oc, what="C4")
    assert g.stats["n_buckets"] == 1


@pytest.mark.parametrize("over", [dict(mix_donors=1), dict(easy=1), dict(max_score=1.0), dict(auto_share=1000000 if False else 100),
                                  dict(join_cdr3_ident=0.0, max_cdr3_diffs=15), dict(max_diffs=55), dict(old_light=1),
                                  dict(cdr3_mult=100), dict(comp_filt=1000000), dict(max_degradation=0), dict(ref_v_trim=0, ref_j_trim=0),
                                  dict(degradation_mode=1), dict(degradation_mode=1, max_degradation=0), dict(max_degradation=1), dict(max_degradation=40),
                                  dict(old_mult=1), dict(old_mult=1, max_score=1e7)])
def test_parameter_variants(over):
    pack, _ = synth.generate(n_cells=6000, n_donors=3, seed=11)
    g, oe, oc, _ = _run_both(pack, **over)
    assert_same_join(g, oe, oc, max_score=over.get("max_score", 100000.0), what=str(over))


@pytest.mark.parametrize("case", range(10))
def test_randomized_repertoires_and_parameters(case):
    """Random repertoire shapes (size, donors, BCR/TCR, naive / public / indel / multi-chain fractions, family size
    law) under random parameter settings: bit-exact against the oracle every time."""
    rng = np.random.default_rng(1000 + case)
    is_bcr = case % 3 != 2                                       # cases 2, 5, 8: TCR rules
    gen = dict(n_cells=int(rng.integers(1500, 20000)), n_donors=int(rng.integers(1, 5)), is_bcr=is_bcr, seed=int(rng.integers(1, 1 << 30)),
               frac_naive=float(rng.uniform(0.1, 0.8)), frac_public=float(rng.uniform(0.0, 0.15)), frac_indel=float(rng.uniform(0.0, 0.08)),
               frac_3chain=float(rng.uniform(0.0, 0.08)), frac_1chain=float(rng.uniform(0.0, 0.1)), family_alpha=float(rng.uniform(1.2, 2.5)))
    if rng.random() < 0.3:
        gen["max_family"] = int(rng.integers(200, 3000))         # a few very large families: multi-segment tile rows
    over = {}
    if rng.random() < 0.5: over["mix_donors"] = 1
    if rng.random() < 0.3: over["max_score"] = float(10.0 ** rng.uniform(0, 7))
    if rng.random() < 0.3: over["auto_share"] = int(rng.integers(3, 40))
    if rng.random() < 0.3: over["join_cdr3_ident"] = float(rng.uniform(70, 100))
    if rng.random() < 0.3: over["max_cdr3_diffs"] = int(rng.integers(0, 12))
    if rng.random() < 0.3: over["cdr3_mult"] = int(rng.integers(1, 8))
    if rng.random() < 0.3: over["comp_filt"] = int(rng.integers(0, 20))
    if rng.random() < 0.3: over["max_degradation"] = int(rng.integers(0, 5))
    if rng.random() < 0.4: over["degradation_mode"] = 1
    if rng.random() < 0.2: over["easy"] = 1
    if rng.random() < 0.2: over["fwr1_cdr12_delta"] = float(rng.uniform(0, 40))
    if rng.random() < 0.2: over["max_diffs"] = int(rng.integers(5, 120))
    pack, _ = synth.generate(**gen)
    g, oe, oc, ost = _run_both(pack, **over)
    assert_same_join(g, oe, oc, max_score=over.get("max_score", 100000.0), what=f"case {case}: {gen} {over}")
    assert g.stats["pairs_candidate"] == ost["pairs_candidate"]


def test_accept_graph_is_a_superset_forest():
    """The GPU scores ALL pairs; its emitted edges must be the lexicographically-first spanning
    forest of the oracle's full accept graph."""
    pack, _ = synth.generate(n_cells=3000, n_donors=1, seed=5)
    params = default_params()
    g = join_exacts(pack, params)
    bs = pack.bucket_starts()
    allowed = set()
    for i, j in zip(bs[:-1], bs[1:]):
        if j - i >= 2:
            for e in O.accept_graph(pack.struct, params, int(i), int(j)):
                allowed.add((int(e["k1"]), int(e["k2"])))
    assert set(g.raw_joins) <= allowed


def test_empty_and_degenerate_inputs():
    pack, _ = synth.generate(n_cells=2000, seed=3)
    empty = pack.subset([])
    g = join_exacts(empty)
    assert len(g.edges) == 0 and len(g.class_of) == 0
    one = pack.subset([5])
    g = join_exacts(one)
    assert len(g.edges) == 0 and list(g.class_of) == [0]
    # ragged: every entry in its own bucket
    bs = pack.bucket_starts()
    sub = pack.subset(bs[:-1])              # keep the owner alive: .struct points into its arrays
    g = join_exacts(sub)
    oe, oc, _ = O.join_exacts(sub.struct)
    assert_same_join(g, oe, oc, what="singletons")


def test_handle_reuse_and_resident_runs():
    pack, _ = synth.generate(n_cells=8000, n_donors=2, seed=21)
    j = Joiner(default_params())
    r1 = j.run(pack.struct)
    r2 = j.run(pack.struct)          # idempotent on a reused handle (tables cached)
    assert np.array_equal(r1.edges, r2.edges) and np.array_equal(r1.class_of, r2.class_of)
    j.upload(pack.struct)
    s1 = j.run_resident()
    s2 = j.run_resident()
    assert s1["pairs_accepted"] == s2["pairs_accepted"] == r1.stats["pairs_accepted"]
    assert s1["edges_after_prune"] == r1.stats["edges_after_prune"]
    # a TCR input on the same handle (tables keyed by is_bcr are rebuilt)
    tcr, _ = synth.generate(n_cells=5000, n_donors=2, is_bcr=False, seed=22)
    r3 = j.run(tcr.struct)
    oe, oc, _ = O.join_exacts(tcr.struct)
    assert_same_join(r3, oe, oc, what="tcr-on-reused-handle")
    r4 = j.run(pack.struct)
    assert np.array_equal(r1.edges, r4.edges)
    j.close()


def test_tables_are_built_on_demand_without_waiting_for_the_key_pass(monkeypatch):
    """A handle that has tables does not wait for the key pass any more: an input with lengths it has no tables for
    runs, finds them missing (ERR_TABLES_MISSING), builds them and reruns — same result as the waiting path."""
    small, _ = synth.generate(n_cells=1500, n_donors=1, seed=31)
    big, _ = synth.generate(n_cells=30000, n_donors=3, seed=32)            # many more (len0 + len1) and CDR3 length pairs
    j = Joiner(default_params())
    j.run(small.struct)                                                    # first run on the handle: the waiting path
    r_big = j.run(big.struct)                                              # optimistic run, tables missing, rebuilt, rerun
    oe, oc, _ = O.join_exacts(big.struct, threads=4)
    assert_same_join(r_big, oe, oc, what="tables-on-demand")
    assert r_big.stats["pairs_candidate"] > 0 and r_big.stats["n_buckets"] > 0      # the key-pass counters survive the merged sync
    r_again = j.run(big.struct)                                            # now fully optimistic
    assert np.array_equal(r_big.edges, r_again.edges) and np.array_equal(r_big.class_of, r_again.class_of)
    assert r_again.stats["pairs_candidate"] == r_big.stats["pairs_candidate"] and r_again.stats["n_buckets"] == r_big.stats["n_buckets"]
    j.close()
    monkeypatch.setenv("EJOIN_SYNC_TABLES", "1")
    j2 = Joiner(default_params())
    j2.run(small.struct)
    r_sync = j2.run(big.struct)
    assert np.array_equal(r_big.edges, r_sync.edges) and r_sync.stats["pairs_candidate"] == r_big.stats["pairs_candidate"]
    j2.close()


def test_sharded_run_equals_single():
    pack, _ = synth.generate(n_cells=12000, n_donors=2, seed=31)
    single = join_exacts(pack)
    for world in (2, 3):
        finals = []
        for rank in range(world):
            j = Joiner(default_params())
            raw, st = j.run_shard(pack.struct, rank, world)
            finals.append(j.finish(pack.struct, raw))
            j.close()
        edges = np.concatenate([f.edges for f in finals])
        order = np.lexsort((edges["k2"], edges["k1"]))
        assert np.array_equal(edges[order], single.edges)
        from enclone_b200.dist import partition
        cls, ncls = partition(pack.struct, edges["k1"], edges["k2"])
        assert np.array_equal(cls, single.class_of) and ncls == len(np.unique(single.class_of))


def test_shard_final_and_gpu_partition():
    pack, _ = synth.generate(n_cells=12000, n_donors=2, seed=31)
    single = join_exacts(pack)
    world = 3
    parts = []
    for rank in range(world):
        j = Joiner(default_params())
        mode, fin, raw, st = j.run_shard_final(pack.struct, rank, world)
        assert mode == 0 and raw is None
        parts.append(fin.edges)
        j.close()
    edges = np.concatenate(parts)                     # each rank's list is in raw_joins order; the ranks' bucket ranges interleave
    edges = edges[np.lexsort((edges["k2"], edges["k1"]))]
    assert np.array_equal(edges, single.edges)
    j = Joiner(default_params())
    cls, ncls = j.partition(pack.struct, edges["k1"], edges["k2"])
    assert np.array_equal(cls, single.class_of) and ncls == len(np.unique(single.class_of))
    j.close()


def test_stride_mode_giant_bucket():
    pack, _ = synth.generate_config("C4", 0.01)
    single = join_exacts(pack)
    world = 2
    raws = []
    js = []
    for rank in range(world):
        j = Joiner(default_params())
        raw, st = j.run_shard(pack.struct, rank, world)
        raws.append(raw)
        js.append(j)
    allraw = np.concatenate(raws)
    fin = js[0].finish(pack.struct, allraw)
    assert np.array_equal(fin.edges, single.edges) and np.array_equal(fin.class_of, single.class_of)
    for j in js:
        j.close()


def test_ambiguous_bases_in_reads_and_germlines():
    """'N' and lower-case bytes: in the V..J reads (outside the CDR3s) they send the entry to the bytewise path;
    in the germline table they count as a difference from every read base (the N plane of k_pack_ref)."""
    pack, _ = synth.generate(n_cells=6000, n_donors=2, seed=77)
    rng = np.random.default_rng(5)
    a = pack.arrays
    n = pack.n_info
    arena = a["tig_arena"]
    for k in rng.choice(n, size=n // 40, replace=False):
        for c in range(2):
            ln, cs, cl = int(a["len"][c][k]), int(a["cdr3_start"][c][k]), int(a["cdr3_len"][c][k])
            if ln < 60:
                continue
            pos = int(rng.integers(0, ln))
            if cs <= pos < cs + cl:                       # CDR3 characters must stay ACGT (documented limit)
                continue
            arena[int(a["tig_off"][c][k]) + pos] = ord("N") if rng.random() < 0.7 else ord("a")
    ref = a["refseq_arena"]
    for i in rng.choice(len(a["refseq_len"]), size=max(1, len(a["refseq_len"]) // 5), replace=False):
        ln = int(a["refseq_len"][i])
        for pos in rng.integers(0, ln, size=3):
            b = ref[int(a["refseq_off"][i]) + int(pos)]
            ref[int(a["refseq_off"][i]) + int(pos)] = ord("N") if rng.random() < 0.5 else (b | 0x20)   # lower case: same 2-bit code, different byte
    g, oe, oc, _ = _run_both(pack)
    assert_same_join(g, oe, oc, what="ambiguous bases")
    assert g.stats["pairs_generic"] > 0


def test_narrow_trims_make_windows_overlap():
    """REF_V_TRIM / REF_J_TRIM of 2: the V and J windows of short light chains overlap, positions in the overlap are
    counted against both germlines (bytewise path), everything else shifts its window ends."""
    pack, _ = synth.generate(n_cells=6000, n_donors=2, seed=78)
    base = join_exacts(pack)
    g, oe, oc, _ = _run_both(pack, ref_v_trim=2, ref_j_trim=2)
    assert_same_join(g, oe, oc, what="trims 2/2")
    g0, oe0, oc0, _ = _run_both(pack, ref_v_trim=0, ref_j_trim=0)
    assert_same_join(g0, oe0, oc0, what="trims 0/0")
    assert g0.stats["pairs_generic"] >= base.stats["pairs_generic"]


def test_stride_mode_segmented_rows_three_ranks():
    """A 10,000-entry bucket: rows of up to 79 tiles are cut into 10 segments (extra slot rows, capacity
    retry) and dealt to three ranks row by row; the gathered edges must replay to the single-GPU join."""
    pack, _ = synth.generate_config("C4", 0.05)
    single = join_exacts(pack)
    oe, oc, _ = O.join_exacts(pack.struct)
    assert_same_join(single, oe, oc, what="C4x0.05 single")
    world = 3
    raws, js = [], []
    for rank in range(world):
        j = Joiner(default_params())
        raw, st = j.run_shard(pack.struct, rank, world)
        raws.append(raw)
        js.append(j)
    fin = js[0].finish(pack.struct, np.concatenate(raws))
    assert np.array_equal(fin.edges, single.edges) and np.array_equal(fin.class_of, single.class_of)
    for j in js:
        j.close()


def test_upload_shard_resident_matches_shard_run():
    pack, _ = synth.generate(n_cells=12000, n_donors=2, seed=31)
    world = 3
    total = 0
    for rank in range(world):
        j = Joiner(default_params())
        mode, fin, raw, st = j.run_shard_final(pack.struct, rank, world)
        j.upload_shard(pack.struct, rank, world)
        s = j.run_resident()
        assert s["pairs_accepted"] == st["pairs_accepted"] and s["edges_after_prune"] == len(fin.edges)
        total += len(fin.edges)
        j.close()
    assert total == len(join_exacts(pack).edges)


@pytest.mark.parametrize("name,scale,over", [("C3", 1.0, {}), ("C3", 1.0, dict(mix_donors=1)), ("C5", 1.0, {}), ("C4", 0.1, {}), ("C4", 1.0, {}),
                                             ("C2", 1.0, {})])
def test_full_size_configs_against_oracle(name, scale, over):
    """BASELINE.json's configurations at full size: bit-exact edges, per-edge integers and partition against the
    multithreaded CPU oracle, plus size-independent properties.  The 200,000-entry bucket of C4 is scored by the
    oracle's row-parallel mode (tests/test_oracle_kat.py checks that mode against the sequential scan)."""
    import os
    sp, _ = synth.generate_config(name, scale, copy=False)
    params = default_params(**over)
    j = Joiner(params)
    g = j.run(sp.struct)
    g2 = j.run(sp.struct)                                    # idempotent on a warm handle
    assert np.array_equal(g.edges, g2.edges) and np.array_equal(g.class_of, g2.class_of)
    j.close()
    # properties that need no oracle
    key = g.edges["k1"].astype(np.int64) << 32 | g.edges["k2"]
    assert (np.diff(key) > 0).all()                          # raw_joins strictly ascending in (k1,k2)
    assert (g.edges["k1"] < g.edges["k2"]).all()
    assert (g.class_of[g.edges["k1"]] == g.class_of[g.edges["k2"]]).all()
    assert (g.class_of <= np.arange(len(g.class_of))).all()  # labels are min members
    assert (g.class_of[g.class_of] == g.class_of).all()
    if name == "C4" and scale >= 0.5:
        O.set_parallel_bucket(50000, os.cpu_count() or 1)
    try:
        oe, oc, ost = O.join_exacts(sp.struct, params, threads=os.cpu_count() or 1)
    finally:
        O.set_parallel_bucket(0, 1)
    assert_same_join(g, oe, oc, what=f"{name} x{scale} {over}")
    assert g.stats["pairs_candidate"] == ost["pairs_candidate"]
    # the GPU compares every pair of a comparison group; the oracle only those join_core did not skip
    assert g.stats["pairs_compared"] >= ost["pairs_same_cdr3_len"]
    if over.get("mix_donors"):
        assert g.edges["err"].sum() > 0            # cross-donor joins are kept and flagged (JOIN ERROR)


def test_candidate_buffer_overflow_retry(monkeypatch):
    """A deliberately tiny candidate buffer: the pipeline notices the overflow (the counts stay exact),
    re-sizes and re-runs; the result is unchanged."""
    pack, _ = synth.generate(n_cells=6000, n_donors=2, seed=41)
    ref = join_exacts(pack)
    monkeypatch.setenv("EJOIN_CAND_CAP", "64")
    got = join_exacts(pack)
    assert np.array_equal(ref.edges, got.edges) and np.array_equal(ref.class_of, got.class_of)


def test_unsupported_inputs_fail_loudly():
    from enclone_b200._lib import EjoinError
    pack, _ = synth.generate(n_cells=500, seed=42)
    # AUTO_SHARES beyond the p1 table depth is refused at handle creation
    with pytest.raises(EjoinError) as ei:
        Joiner(default_params(auto_share=100000))
    assert ei.value.code == -4
    # a CDR3 with a character outside ACGT
    a = {k: (list(v) if isinstance(v, list) else v) for k, v in pack.arrays.items()}
    arena = a["cdr3_arena"].copy()
    arena[int(a["cdr3_off"][0][3])] = ord("N")
    a["cdr3_arena"] = arena
    bad = JoinPack(a, is_bcr=True)
    with pytest.raises(EjoinError) as ei:
        join_exacts(bad)
    assert ei.value.code == -4 and "ACGT" in str(ei.value)
    # a germline index outside the table
    a = {k: (list(v) if isinstance(v, list) else v) for k, v in pack.arrays.items()}
    vs = a["vs_seq"][0].copy(); vs[5] = 10 ** 6; a["vs_seq"] = [vs, a["vs_seq"][1]]
    with pytest.raises(EjoinError) as ei:
        join_exacts(JoinPack(a, is_bcr=True))
    assert ei.value.code == -1
    # abi version mismatch
    good = JoinPack({k: v for k, v in pack.arrays.items()}, is_bcr=True)
    good.struct.abi_version = 99
    with pytest.raises(EjoinError) as ei:
        join_exacts(good)
    assert ei.value.code == -1


def test_zero_copy_in_place_arena(monkeypatch):
    """Pinned, padded arenas flagged EJOIN_F_TIGS_IN_PLACE are read in place (only the needed entries
    cross PCIe); the result must equal the uploaded-arena path and the oracle."""
    import os
    sp, _ = synth.generate(n_cells=60000, n_donors=3, seed=77, copy=False, pinned=True)
    assert sp.struct.flags & 1
    params = default_params()
    j = Joiner(params)
    g = j.run(sp.struct)
    moved_in_place = g.stats["h2d_bytes"]
    j.close()
    oe, oc, _ = O.join_exacts(sp.struct, params, threads=os.cpu_count() or 1)
    assert_same_join(g, oe, oc, what="zero-copy")
    monkeypatch.setenv("EJOIN_NO_ZERO_COPY", "1")
    j = Joiner(params)
    g2 = j.run(sp.struct)
    j.close()
    assert np.array_equal(g.edges, g2.edges) and np.array_equal(g.class_of, g2.class_of)
    assert moved_in_place < g2.stats["h2d_bytes"]          # fewer bytes crossed PCIe


def test_finish_rejects_edges_outside_the_resident_range():
    """ejoin_finish on a handle that holds one contiguous shard must refuse gathered edges of other shards
    (EJOIN_E_ARG), not index its resident rows out of bounds."""
    from enclone_b200._lib import EjoinError
    from enclone_b200._abi import RAW_EDGE_DTYPE
    pack, _ = synth.generate(n_cells=12000, n_donors=2, seed=31)
    j0, j1 = Joiner(default_params()), Joiner(default_params())
    raw0, _ = j0.run_shard(pack.struct, 0, 2)
    raw1, _ = j1.run_shard(pack.struct, 1, 2)
    assert len(raw0) and len(raw1)
    fin = j0.finish(pack.struct, raw0)                       # its own edges: fine
    assert len(fin.edges) > 0
    with pytest.raises(EjoinError) as ei:
        j0.finish(pack.struct, np.concatenate([raw0, raw1]))
    assert ei.value.code == -1 and "outside" in str(ei.value)
    bad = np.zeros(1, RAW_EDGE_DTYPE)
    bad["k1"], bad["k2"] = 5, 5                              # k1 >= k2
    with pytest.raises(EjoinError):
        j0.finish(pack.struct, bad)
    j0.close(); j1.close()


def test_resident_run_after_in_place_run_is_refused():
    from enclone_b200._lib import EjoinError
    sp, _ = synth.generate(n_cells=20000, n_donors=2, seed=78, copy=False, pinned=True)
    j = Joiner(default_params())
    j.run(sp.struct)
    with pytest.raises(EjoinError) as ei:
        j.run_resident()
    assert ei.value.code == -1
    j.upload(sp.struct)
    assert j.run_resident()["pairs_candidate"] > 0
    j.close()


def test_null_pack_arrays_are_an_argument_error():
    from enclone_b200._lib import EjoinError
    import ctypes as C
    pack, _ = synth.generate(n_cells=500, seed=42)
    for field in ("hcomp", "ncells", "refseq_len", "c_ref_id_light"):
        good = JoinPack({k: v for k, v in pack.arrays.items()}, is_bcr=True)
        setattr(good.struct, field, C.cast(None, type(getattr(good.struct, field))))
        with pytest.raises(EjoinError) as ei:
            join_exacts(good)
        assert ei.value.code == -1 and "null array" in str(ei.value)
    a = {k: (list(v) if isinstance(v, list) else v) for k, v in pack.arrays.items()}
    ex = a["exact_id"].copy(); ex[7] = 10 ** 7; a["exact_id"] = ex
    with pytest.raises(EjoinError) as ei:
        join_exacts(JoinPack(a, is_bcr=True))
    assert ei.value.code == -1 and "exact_id" in str(ei.value)


# ---- compact input (EJOIN_F_TIGS_2BIT | EJOIN_F_ENTRY_RECS): the form a caller holding CloneInfo.tigsp passes ----
def _compact_vs_everything(pack, what, **over):
    from enclone_b200.join import CompactPack
    params = default_params(**over)
    cp = CompactPack(pack)
    j = Joiner(params)
    g = j.run(cp.struct)                                   # 2-bit words read in place over PCIe, 64-byte records
    a = j.run(pack.struct if hasattr(pack, "struct") else pack)      # the ASCII / parallel-array form on the same handle
    j.close()
    assert np.array_equal(g.edges, a.edges) and np.array_equal(g.class_of, a.class_of), f"{what}: compact != ascii"
    oe, oc, _ = O.join_exacts(pack.struct, params, threads=8)
    assert_same_join(g, oe, oc, max_score=over.get("max_score", 100000.0), what=what)
    cp.close()
    return g, a


@pytest.mark.parametrize("name,scale", [("C1", 1.0), ("C2", 0.3), ("C3", 0.03), ("C5", 0.05), ("C4", 0.03)])
def test_compact_input_matches_ascii_and_oracle(name, scale):
    pack, _ = synth.generate_config(name, scale)
    g, a = _compact_vs_everything(pack, f"compact {name}")
    assert g.stats["h2d_bytes"] < 0.6 * a.stats["h2d_bytes"]


def test_compact_input_bytewise_paths():
    """Pairs the bit planes cannot represent are scored from bytes rebuilt out of the plane words (2-bit chains) or read
    from the small ASCII arena (chains with a deletion marker / non-ACGT byte): indels, N bases, overlapping windows,
    different germlines under both MAX_DEGRADATION readings."""
    pack, _ = synth.generate(n_cells=9000, n_donors=3, seed=91, frac_indel=0.06, frac_public=0.1)
    a = pack.arrays
    rng = np.random.default_rng(9)
    arena = a["tig_arena"]
    for k in rng.choice(pack.n_info, size=pack.n_info // 30, replace=False):
        c = int(rng.integers(0, 2))
        ln, cs, cl = int(a["len"][c][k]), int(a["cdr3_start"][c][k]), int(a["cdr3_len"][c][k])
        if ln < 60:
            continue
        pos = int(rng.integers(0, ln))
        if not (cs <= pos < cs + cl):
            arena[int(a["tig_off"][c][k]) + pos] = ord("N")
    g, _ = _compact_vs_everything(pack, "compact generic", mix_donors=1)
    assert g.stats["pairs_generic"] > 0
    _compact_vs_everything(pack, "compact trims", ref_v_trim=2, ref_j_trim=2)
    _compact_vs_everything(pack, "compact degradation 1", degradation_mode=1, max_degradation=1, mix_donors=1)
    _compact_vs_everything(pack, "compact degradation 0", max_degradation=1, mix_donors=1)


def test_compact_input_sharded_and_uploaded(monkeypatch):
    from enclone_b200.join import CompactPack
    pack, _ = synth.generate(n_cells=15000, n_donors=2, seed=33, frac_indel=0.03)
    cp = CompactPack(pack)
    single = join_exacts(pack)
    world = 3
    parts, raws, js = [], [], []
    for rank in range(world):
        j = Joiner(default_params())
        mode, fin, raw, st = j.run_shard_final(cp.struct, rank, world)
        if mode == 0:
            parts.append(fin.edges)
            j.upload_shard(cp.struct, rank, world)               # resident form of the same shard
            s = j.run_resident()
            assert s["edges_after_prune"] == len(fin.edges)
        else:
            raws.append(raw)
        js.append(j)
    if parts:
        assert not raws
        allp = np.concatenate(parts)
        assert np.array_equal(allp[np.lexsort((allp["k2"], allp["k1"]))], single.edges)
    else:
        fin = js[0].finish(cp.struct, np.concatenate(raws))
        assert np.array_equal(fin.edges, single.edges) and np.array_equal(fin.class_of, single.class_of)
    for j in js:
        j.close()
    # stride mode on a giant bucket
    big, _ = synth.generate_config("C4", 0.02)
    cb = CompactPack(big)
    ref = join_exacts(big)
    raws, js = [], []
    for rank in range(2):
        j = Joiner(default_params())
        raw, st = j.run_shard(cb.struct, rank, 2)
        raws.append(raw); js.append(j)
    fin = js[0].finish(cb.struct, np.concatenate(raws))
    assert np.array_equal(fin.edges, ref.edges) and np.array_equal(fin.class_of, ref.class_of)
    for j in js:
        j.close()
    # no zero-copy: the 2-bit arena is uploaded
    monkeypatch.setenv("EJOIN_NO_ZERO_COPY", "1")
    g2 = join_exacts(cp)
    assert np.array_equal(g2.edges, single.edges) and np.array_equal(g2.class_of, single.class_of)
    cp.close(); cb.close()


def test_compact_full_size_c3():
    import os
    from enclone_b200.join import CompactPack
    sp, _ = synth.generate_config("C3", 1.0, copy=False)
    cp = CompactPack(sp)
    params = default_params()
    j = Joiner(params)
    g = j.run(cp.struct)
    j.close()
    oe, oc, _ = O.join_exacts(sp.struct, params, threads=os.cpu_count() or 1)
    assert_same_join(g, oe, oc, what="compact C3")
    assert g.stats["h2d_bytes"] <= 350e6
    cp.close()


@pytest.mark.parametrize("name,scale,stride,compact,mma", [("C3", 0.2, 0, 1, 0), ("C3", 0.2, 0, 0, 0), ("C4", 0.05, 1, 1, 0), ("C3", 0.05, 1, 0, 0), ("C5", 0.05, 0, 1, 0),
                                                          ("C4", 0.1, 1, 1, 1), ("C3", 0.2, 0, 1, 1)])
def test_two_rank_nccl_sharded_join(name, scale, stride, compact, mma):
    """Real NCCL, one process per GPU (torchrun, 2 ranks): contiguous bucket ranges and the shared giant bucket (stride
    mode), the exchange inside the library; rank 0 checks the gathered result against the one-GPU join and the oracle.
    mma = 1: tier 1 forced onto the tensor cores for every group that qualifies (in stride mode each rank sweeps its rows of the
    one group with k_sweep_mma)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    here = os.path.dirname(os.path.abspath(__file__))
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(here, "mgpu_worker.py"), name, str(scale), str(stride), str(compact)]
    env = dict(os.environ)
    if mma:
        env["EJOIN_MMA_MIN_TILES"], env["EJOIN_MMA_CAP_TILES"] = "1", "20000"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "MGPU_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_compact_input_dnastring_words(monkeypatch):
    """The 2-bit words exactly as a DnaString holds them (interleaved; no EJOIN_F_TIGS_PLANES): the packing kernel transposes them
    itself.  Same joins as the plane-split form and as the oracle."""
    from enclone_b200.join import CompactPack
    pack, _ = synth.generate(n_cells=9000, n_donors=2, seed=92, frac_indel=0.04)
    monkeypatch.setenv("EJOIN_COMPACT_DNASTRING", "1")
    cp = CompactPack(pack)
    monkeypatch.delenv("EJOIN_COMPACT_DNASTRING")
    assert cp.struct.flags & 2 and not cp.struct.flags & 16
    cp2 = CompactPack(pack)
    assert cp2.struct.flags & 16
    g, g2 = join_exacts(cp), join_exacts(cp2)
    assert np.array_equal(g.edges, g2.edges) and np.array_equal(g.class_of, g2.class_of)
    oe, oc, _ = O.join_exacts(pack.struct, default_params(), threads=8)
    assert_same_join(g, oe, oc, what="compact, DnaString words")
    cp.close(); cp2.close()


@pytest.mark.parametrize("name,scale,min_tiles", [("C2", 0.3, 1), ("C3", 0.03, 1), ("C4", 0.06, 1), ("C4", 0.06, 3), ("C5", 0.05, 1)])
def test_tensor_core_sweep_equals_popcount_sweep_and_oracle(monkeypatch, name, scale, min_tiles):
    """Tier 1 on the tensor cores (k_sweep_mma: int8 tetrahedron rows, tcgen05.mma, survivor masks read back from TMEM) forced onto every
    group it can take — from one partial tile up, diagonal tiles, several segments, groups mixed with popcount-swept ones — against
    the popcount sweep (tensor-core path off) and the oracle: identical joins, payloads and partition, identical survivor counts."""
    pack, _ = synth.generate_config(name, scale)
    monkeypatch.setenv("EJOIN_MMA_MIN_TILES", "0")
    ref = join_exacts(pack)
    monkeypatch.setenv("EJOIN_MMA_MIN_TILES", str(min_tiles))
    monkeypatch.setenv("EJOIN_MMA_CAP_TILES", "20000" if name != "C5" else "40")       # C5: most groups find no slot and stay with the popcount sweep
    got = join_exacts(pack)
    assert np.array_equal(ref.edges, got.edges) and np.array_equal(ref.class_of, got.class_of)
    for k in ("pairs_candidate", "pairs_compared", "pairs_tier2"):
        assert ref.stats[k] == got.stats[k], k
    oe, oc, ost = O.join_exacts(pack.struct, default_params(), threads=8)
    assert_same_join(got, oe, oc, what=f"{name} mma>={min_tiles}")
