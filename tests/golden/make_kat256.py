#!/usr/bin/env python
"""Build tests/golden/kat256.json — the one known-answer vector the reference holds for join_one.

Source (read-only, only present in the build container):
  /root/reference/enclone_exec/testx/inputs/flaky8{a,b}/outs/all_contig_annotations.json   (inputs)
  /root/reference/enclone_exec/testx/inputs/outputs/enclone_test256_output                 (golden, lines 29-90)
  command: enclone_testlist/src/main_testlist.rs:625-629

The two cells' V..J sequences, CDR3s and annotation coordinates are taken verbatim from the
JSON.  The V(D)J germline reference is NOT in the repository, so the germline
used for shares/indeps is reconstructed to be consistent with everything the golden prints:
share positions 69,146,147,148,223 on the heavy V, per-contig V/J mismatch counts (16/17 and
6/6 heavy, 6/5 and 0/1 light), and the doubly-different position 205.
"""
import json
import os
import re

REF = "/root/reference/enclone_exec/testx/inputs"
HERE = os.path.dirname(os.path.abspath(__file__))


def cell(ds, barcode):
    out = {}
    for c in json.load(open(f"{REF}/{ds}/outs/all_contig_annotations.json")):
        if c["barcode"] != barcode or not c["productive"]:
            continue
        ann = {a["feature"]["region_type"]: a for a in c["annotations"]}
        v, j = ann["L-REGION+V-REGION"], ann["J-REGION"]
        chain = v["feature"]["chain"]
        s, e = v["contig_match_start"], j["contig_match_end"]
        out["heavy" if chain == "IGH" else "light"] = dict(
            seq=c["sequence"][s:e], cdr3=c["cdr3_seq"], cdr3_start=c["cdr3_start"] - s,
            vlen=v["annotation_match_end"] - v["annotation_match_start"],
            jlen=j["annotation_match_end"] - j["annotation_match_start"],
            vgene=v["feature"]["gene_name"], jgene=j["feature"]["gene_name"], cgene=ann["C-REGION"]["feature"]["gene_name"])
    return out


def other(b, *avoid):
    for x in "ACGT":
        if x != b and x not in avoid:
            return x


def main():
    c1 = cell("flaky8b", "AATCCAGGTACAGACG-1")   # entry k1 ("flaky8b.clono1")
    c2 = cell("flaky8a", "ACGAGCCGTCGGGTCT-1")   # entry k2 ("flaky8a.clono0")
    gold = re.sub(r"\x1b\[[0-9;]*m", "", open(f"{REF}/outputs/enclone_test256_output").read())
    exp = dict(
        shares=int(re.search(r"shares = (\d+)", gold).group(1)),
        indeps=int(re.search(r"indeps = (\d+)", gold).group(1)),
        shares_details=[int(x) for x in re.search(r"breakdown = V(\d+), J(\d+); V(\d+), J(\d+)", gold).groups()],
        share_pos_v=[int(x) for x in re.search(r"share pos v = ([\d,]+)", gold).group(1).split(",")],
        cd=int(re.search(r"CDR3 nucleotide diffs = (\d+)", gold).group(1)),
        fwr1_diffs=int(re.search(r"FWR1 diffs = (\d+)", gold).group(1)),
        cdr1_diffs=int(re.search(r"CDR1 diffs = (\d+)", gold).group(1)),
        cdr2_diffs=int(re.search(r"CDR2 diffs = (\d+)", gold).group(1)),
        cdr3h_diffs=int(re.search(r"heavy chain CDR3 diffs = (\d+)", gold).group(1)),
        fwr1_ident=re.search(r"FWR1 = ([\d.]+)%", gold).group(1),
        cdr12_ident=re.search(r"CDR1-2 = ([\d.]+)%", gold).group(1),
        p1=re.search(r"by accident = ([\d.]+)", gold).group(1),
        k=int(re.search(r"using k = (\d+)", gold).group(1)),
        d=int(re.search(r"d = (\d+), n", gold).group(1)),
        n=int(re.search(r"n = (\d+)", gold).group(1)),
        mult=re.search(r"cd as usize\) = ([\d.]+)", gold).group(1),
        score=re.search(r"score = p1 \* mult = ([\d.]+)", gold).group(1),
        join_error="JOIN ERROR" in gold,
        vmis_heavy=[16, 17], jmis_heavy=[6, 6], vmis_light=[6, 5], jmis_light=[0, 1],
    )
    germ = {}
    for ch, shares_pos, dbl in (("heavy", exp["share_pos_v"], [205]), ("light", [], [])):
        t1, t2 = c1[ch]["seq"], c2[ch]["seq"]
        assert len(t1) == len(t2)
        n, vl, jl = len(t1), c1[ch]["vlen"], c1[ch]["jlen"]
        v = list(t1[:vl])
        # shared mutations: both contigs carry the same non-germline base
        for p in shares_pos:
            assert t1[p] == t2[p]
            v[p] = other(t1[p])
        for p in dbl:
            assert t1[p] != t2[p]
            v[p] = other(t1[p], t2[p])
        diff = [p for p in range(vl) if t1[p] != t2[p] and p not in dbl]
        vm1, vm2 = exp["vmis_" + ch]
        base = len(shares_pos) + len(dbl)
        # a = differing positions where contig 1 is the mutated one; e = mismatches common to
        # both inside the last 15 V bases (outside the window): vm1 = base+a+e, vm2 = base+(len(diff)-a)+e
        e2 = vm1 + vm2 - 2 * base - len(diff)
        assert e2 >= 0 and e2 % 2 == 0, (ch, e2)
        e = e2 // 2
        a = vm1 - base - e
        assert 0 <= a <= len(diff)
        for i, p in enumerate(diff):
            v[p] = t2[p] if i < a else t1[p]
        tail = [p for p in range(vl - 15, vl) if t1[p] == t2[p]][-e:] if e else []
        for p in tail:
            v[p] = other(t1[p])
        # J: aligned to the end of V..J.  Mismatches sit in the first 15 J bases (outside the window).
        j = list(t1[n - jl:])
        jm1, jm2 = exp["jmis_" + ch]
        jdiff = [q for q in range(jl) if t1[n - jl + q] != t2[n - jl + q]]
        assert all(q < 15 for q in jdiff)
        com = (jm1 + jm2 - len(jdiff)) // 2
        a = jm1 - com
        for i, q in enumerate(jdiff):
            j[q] = t2[n - jl + q] if i < a else t1[n - jl + q]
        free = [q for q in range(15) if q not in jdiff][:com]
        assert len(free) == com
        for q in free:
            j[q] = other(t1[n - jl + q])
        germ[ch] = dict(v="".join(v), j="".join(j))
        assert sum(x != y for x, y in zip(germ[ch]["v"], t1[:vl])) == vm1
        assert sum(x != y for x, y in zip(germ[ch]["v"], t2[:vl])) == vm2
        assert sum(x != y for x, y in zip(germ[ch]["j"], t1[n - jl:])) == jm1
        assert sum(x != y for x, y in zip(germ[ch]["j"], t2[n - jl:])) == jm2
    # heavy-chain region starts, V..J coordinates: FWR1 is 75 nt and CDR1+CDR2 39 nt in the golden
    # (97.3% = 1-2/75, 92.3% = 1-3/39); the leader of IGHV3-23 is 57 nt.
    regions = dict(fr1_start=57, cdr1_start=132, fr2_start=153, cdr2_start=204, fr3_start=222)
    out = dict(source="enclone_test256_output + flaky8a/flaky8b all_contig_annotations.json",
               k1=c1, k2=c2, germline=germ, regions=regions, expected=exp)
    json.dump(out, open(os.path.join(HERE, "kat256.json"), "w"), indent=1)
    # the join's log block, verbatim (ANSI stripped): "JOIN ERROR" .. last line of "difference pattern for chain 2"
    lines = gold.split("\n")
    i0 = lines.index("JOIN ERROR")
    i1 = next(i for i in range(i0, len(lines)) if lines[i].startswith("chain 1, tig 1"))
    open(os.path.join(HERE, "kat256_join_log.txt"), "w").write("\n".join(lines[i0:i1]) + "\n")
    print("wrote kat256.json; n =", 3 * (len(c1["heavy"]["seq"]) + len(c1["light"]["seq"])))


if __name__ == "__main__":
    main()
