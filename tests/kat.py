"""Builds the two-entry JoinPack of the reference's known-answer vector (regression test 256)
from tests/golden/kat256.json (made by tests/golden/make_kat256.py)."""
import json
import os

import numpy as np

from enclone_b200._abi import EJOIN_NONE, JoinPack

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    return json.load(open(os.path.join(HERE, "golden", "kat256.json")))


def kat_pack(extra_copies=0):
    """Entries 0,1 = the golden pair (k1 = flaky8b cell, k2 = flaky8a cell).  With
    extra_copies > 0 the pair is replicated (same bucket) to give kernels a fuller tile."""
    kat = load()
    cells = [kat["k1"], kat["k2"]] * (1 + extra_copies)
    n = len(cells)
    refs = [kat["germline"]["heavy"]["v"], kat["germline"]["heavy"]["j"],
            kat["germline"]["light"]["v"], kat["germline"]["light"]["j"]]
    ref_off, arena = [], bytearray()
    for r in refs:
        ref_off.append(len(arena))
        arena += r.encode()
        while len(arena) % 16:
            arena.append(0)
    tig, cdr3 = bytearray(), bytearray()
    a = {k: [[], []] for k in ("len", "tig_off", "has_del", "cdr3_len", "cdr3_off", "cdr3_start", "v_ref_id", "j_ref_id",
                               "dref_id", "vs_seq", "js_seq", "vname_id", "vuni_seq", "utr_seq")}
    for cell in cells:
        for c, ch in enumerate(("heavy", "light")):
            s = cell[ch]
            a["len"][c].append(len(s["seq"])); a["tig_off"][c].append(len(tig)); tig += s["seq"].encode()
            a["has_del"][c].append(0); a["cdr3_len"][c].append(len(s["cdr3"])); a["cdr3_off"][c].append(len(cdr3))
            cdr3 += s["cdr3"].encode(); a["cdr3_start"][c].append(s["cdr3_start"])
            # golden l.33-34: "ref v segs = 107=..., 219=..." ; ids 108|IGHV3-23, 55|IGHJ4, 220|IGKV2-30, 171|IGKJ1
            a["v_ref_id"][c].append(108 if c == 0 else 220); a["j_ref_id"][c].append(55 if c == 0 else 171)
            a["dref_id"][c].append(EJOIN_NONE); a["vs_seq"][c].append(0 if c == 0 else 2); a["js_seq"][c].append(1 if c == 0 else 3)
            a["vname_id"][c].append(c); a["vuni_seq"][c].append(0 if c == 0 else 2); a["utr_seq"][c].append(EJOIN_NONE)
    reg = kat["regions"]
    arrays = dict(a)
    arrays.update(
        exact_id=np.arange(n), c_ref_id_light=np.full(n, 170), hcomp=np.full(n, 3),
        fr1_start=np.full(n, reg["fr1_start"]), cdr1_start=np.full(n, reg["cdr1_start"]),
        fr2_start=np.full(n, reg["fr2_start"]), cdr2_start=np.full(n, reg["cdr2_start"]),
        fr3_start=np.full(n, reg["fr3_start"]),
        nchains=np.full(n, 2), ncells=np.full(n, 1),
        donor_lo=np.array([1, 0] * (1 + extra_copies)), donor_hi=np.array([1, 0] * (1 + extra_copies)),
        refseq_off=np.array(ref_off), refseq_len=np.array([len(r) for r in refs]),
        tig_arena=np.frombuffer(bytes(tig), np.uint8), cdr3_arena=np.frombuffer(bytes(cdr3), np.uint8),
        refseq_arena=np.frombuffer(bytes(arena), np.uint8))
    return JoinPack(arrays, is_bcr=True), kat
