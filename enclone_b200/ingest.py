"""all_contig_annotations.json -> JoinPack: the steps of enclone in front of the join, as far as the join's inputs need them.

What it restates (host code, Python; the reference does this in enclone_stuff / enclone_core, un-vendored):
  * input data (enclone_help/src/help1.rs:229-240): productive contigs only, each with an annotated V and J segment; the V..J
    sequence is `sequence[V.contig_match_start : J.contig_match_end]` (SURVEY App. B checked this against the coordinates the
    golden of regression test 256 prints);
  * exact subclonotypes (help1.rs:242-245): cells with the same number of chains, identical V..J sequences and identical C
    segment assignments;
  * `info`: one CloneInfo per (heavy/TRB chain, light/TRA chain) pair of an exact subclonotype, sorted with
    `lens` first, then the V..J bytes.
What it cannot restate: the VDJ reference is not in the repository, so the germline used for shared-mutation
counting is a **consensus** of the dataset's own contigs per V gene and per J gene (the same idea as enclone's donor-reference
derivation, help1.rs:247-264, without its evidence thresholds).  Joins computed on such a pack exercise the whole path on real
contig sequences, CDR3s, indels and chain multiplicities; they are not claimed to equal enclone's on the same data.
FWR / CDR1 / CDR2 bounds are not in the JSON (enclone derives them from reference motifs): they are passed as None.
"""
import json
import re
from collections import Counter, defaultdict

import numpy as np

from ._abi import EJOIN_NONE, EJOIN_NONE16, JoinPack

HEAVY = {"IGH", "TRB", "TRD"}


def _cigar_has_deletion(cigar):
    return "D" in cigar


def _contig_chain(c):
    """(chain type, V..J bytes, cdr3 nt, cdr3 start in V..J coordinates, V gene, J gene, C gene, 5'UTR gene, has_del) or None"""
    ann = {}
    for a in c.get("annotations", []):
        ann.setdefault(a["feature"]["region_type"], a)
    v, j = ann.get("L-REGION+V-REGION"), ann.get("J-REGION")
    if not c.get("productive") or v is None or j is None or not c.get("cdr3_seq"):
        return None
    s, e = v["contig_match_start"], j["contig_match_end"]
    seq = c["sequence"][s:e]
    cs = c["cdr3_start"] - s
    if cs < 0 or cs + len(c["cdr3_seq"]) > len(seq) or seq[cs:cs + len(c["cdr3_seq"])] != c["cdr3_seq"]:
        return None
    cg = ann.get("C-REGION")
    ug = ann.get("5'UTR")
    return dict(chain=v["feature"]["chain"], seq=seq, cdr3=c["cdr3_seq"], cdr3_start=cs, vgene=v["feature"]["gene_name"], vid=v["feature"]["feature_id"],
                jgene=j["feature"]["gene_name"], jid=j["feature"]["feature_id"], cid=cg["feature"]["feature_id"] if cg else None,
                uid=ug["feature"]["feature_id"] if ug else None, vlen=v["annotation_length"], jlen=j["annotation_length"],
                vstart=v["annotation_match_start"], jend_gap=j["annotation_length"] - j["annotation_match_end"],
                has_del=_cigar_has_deletion(v.get("cigar", "")))


def _consensus(rows, length, align_right=False):
    cols = [Counter() for _ in range(length)]
    for r in rows:
        r = r[-length:] if align_right else r[:length]
        off = length - len(r) if align_right else 0
        for i, b in enumerate(r):
            cols[off + i][b] += 1
    return "".join((c.most_common(1)[0][0] if c else "A") for c in cols)


def load_datasets(paths, is_bcr=True, donors=None):
    """paths: list of `.../outs/all_contig_annotations.json` (one per dataset); donors: donor index per dataset (default: all 0).
    Returns (JoinPack, info): info[k] = dict(dataset, exact, barcodes, cdr3_aa-free description) for reporting."""
    donors = donors or [0] * len(paths)
    cells = defaultdict(list)                      # (dataset, barcode) -> chains
    for ds, p in enumerate(paths):
        for c in json.load(open(p)):
            ch = _contig_chain(c)
            if ch is not None and c.get("is_cell", True):
                cells[(ds, c["barcode"])].append(ch)
    # exact subclonotypes: same chains (type, V..J, C gene), heavy chains first
    exacts = defaultdict(list)
    for (ds, bc), chains in cells.items():
        chains.sort(key=lambda x: (x["chain"] not in HEAVY, x["seq"]))
        key = tuple((x["chain"], x["seq"], x["cid"]) for x in chains)
        exacts[key].append((ds, bc, chains))
    ex_list = sorted(exacts.items(), key=lambda kv: kv[0])
    # consensus germlines per V / J gene over the exact subclonotypes (one vote each), V left-aligned, J right-aligned
    vrows, jrows, vlen, jlen = defaultdict(list), defaultdict(list), {}, {}
    for key, members in ex_list:
        for x in members[0][2]:
            vl = min(x["vlen"] - x["vstart"], len(x["seq"]))
            jl = min(x["jlen"] - x["jend_gap"], len(x["seq"]))
            vrows[x["vgene"]].append(x["seq"][:vl]); vlen[x["vgene"]] = max(vlen.get(x["vgene"], 0), vl)
            jrows[x["jgene"]].append(x["seq"][-jl:]); jlen[x["jgene"]] = max(jlen.get(x["jgene"], 0), jl)
    refs, ref_index = [], {}

    def ref_id(s):
        if s not in ref_index:
            ref_index[s] = len(refs); refs.append(s)
        return ref_index[s]

    vcons = {g: _consensus(r, vlen[g]) for g, r in vrows.items()}
    jcons = {g: _consensus(r, jlen[g], align_right=True) for g, r in jrows.items()}
    vname = {g: i for i, g in enumerate(sorted(set(re.sub(r"\*.*$", "", g) for g in vcons)))}
    # info entries
    entries = []
    for e, (key, members) in enumerate(ex_list):
        chains = members[0][2]
        heavy = [x for x in chains if x["chain"] in HEAVY]
        light = [x for x in chains if x["chain"] not in HEAVY]
        pairs = [(h, l) for h in heavy for l in light] or [(h, None) for h in heavy] or [(l, None) for l in light[:1]]
        for h, l in pairs:
            entries.append(dict(exact=e, chains=[h, l] if l is not None else [h]))
    entries.sort(key=lambda en: ([len(c["seq"]) for c in en["chains"]] + [0] * (2 - len(en["chains"])), [c["seq"] for c in en["chains"]]))
    n = len(entries)
    a = {k: [np.zeros(n, dt), np.zeros(n, dt)] for k, dt in (("len", np.uint16), ("tig_off", np.uint64), ("has_del", np.uint8), ("cdr3_len", np.uint16),
                                                                ("cdr3_off", np.uint64), ("cdr3_start", np.uint16))}
    for k in ("v_ref_id", "j_ref_id", "dref_id", "vname_id", "utr_seq"):
        a[k] = [np.full(n, EJOIN_NONE, np.uint32), np.full(n, EJOIN_NONE, np.uint32)]
    for k in ("vs_seq", "js_seq", "vuni_seq"):
        a[k] = [np.zeros(n, np.uint32), np.zeros(n, np.uint32)]
    tig, cdr3 = bytearray(), bytearray()
    exact_id, c_light = np.zeros(n, np.uint32), np.full(n, EJOIN_NONE, np.uint32)
    for k, en in enumerate(entries):
        exact_id[k] = en["exact"]
        for c, x in enumerate(en["chains"]):
            a["len"][c][k] = len(x["seq"]); a["tig_off"][c][k] = len(tig); tig += x["seq"].encode()
            a["has_del"][c][k] = int(x["has_del"]); a["cdr3_len"][c][k] = len(x["cdr3"]); a["cdr3_off"][c][k] = len(cdr3)
            cdr3 += x["cdr3"].encode(); a["cdr3_start"][c][k] = x["cdr3_start"]
            a["v_ref_id"][c][k] = x["vid"]; a["j_ref_id"][c][k] = x["jid"]
            a["vs_seq"][c][k] = a["vuni_seq"][c][k] = ref_id(vcons[x["vgene"]]); a["js_seq"][c][k] = ref_id(jcons[x["jgene"]])
            a["vname_id"][c][k] = vname[re.sub(r"\*.*$", "", x["vgene"])]
            if c == 1 and x["cid"] is not None:
                c_light[k] = x["cid"]
    nx = len(ex_list)
    nchains = np.array([len(m[0][2]) for _, m in ex_list], np.uint32) if nx else np.zeros(0, np.uint32)
    ncells = np.array([len(m) for _, m in ex_list], np.uint32) if nx else np.zeros(0, np.uint32)
    dlo = np.array([min(donors[ds] for ds, _, _ in m) for _, m in ex_list], np.uint32) if nx else np.zeros(0, np.uint32)
    dhi = np.array([max(donors[ds] for ds, _, _ in m) for _, m in ex_list], np.uint32) if nx else np.zeros(0, np.uint32)
    ref_off, arena = [], bytearray()
    for r in refs:
        ref_off.append(len(arena)); arena += r.encode()
        while len(arena) % 16:
            arena.append(0)
    arrays = dict(a)
    none16 = np.full(n, EJOIN_NONE16, np.uint16)
    arrays.update(exact_id=exact_id, c_ref_id_light=c_light, hcomp=np.zeros(n, np.uint16), fr1_start=none16, cdr1_start=none16, fr2_start=none16,
                  cdr2_start=none16, fr3_start=none16, nchains=nchains, ncells=ncells, donor_lo=dlo, donor_hi=dhi,
                  refseq_off=np.array(ref_off, np.uint64), refseq_len=np.array([len(r) for r in refs], np.uint32),
                  tig_arena=np.frombuffer(bytes(tig), np.uint8).copy() if tig else np.zeros(0, np.uint8),
                  cdr3_arena=np.frombuffer(bytes(cdr3), np.uint8).copy() if cdr3 else np.zeros(0, np.uint8),
                  refseq_arena=np.frombuffer(bytes(arena), np.uint8).copy() if arena else np.zeros(0, np.uint8))
    info = [dict(exact=en["exact"], barcodes=[bc for _, bc, _ in ex_list[en["exact"]][1]], datasets=sorted(set(ds for ds, _, _ in ex_list[en["exact"]][1])),
                 cdr3s=[c["cdr3"] for c in en["chains"]], vgenes=[c["vgene"] for c in en["chains"]]) for en in entries]
    return JoinPack(arrays, is_bcr=is_bcr), info
