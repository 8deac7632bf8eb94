"""Extracts known-answer pins for the symmetric grouping from the reference's own golden outputs
(enclone_exec/testx/inputs/outputs/enclone_gtest*_output; commands at enclone_exec/tests/enclone_test2.rs:176-270).
Each golden prints ONE group selected by CDR3=: the pins are "these clonotypes end in one group under these options".
Run in the build container (reads /root/reference); writes tests/golden/gtest_pins.json."""
import json
import os
import re

OUT = "/root/reference/enclone_exec/testx/inputs/outputs"
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def clonotypes(t):
    """[(header cells, [rows of (chain1 text, chain2 text)])] for every [g.c] CLONOTYPE block of a golden"""
    txt = ANSI.sub("", open(os.path.join(OUT, f"enclone_gtest{t}_output")).read())
    blocks = re.split(r"\n\[\d+\.\d+\] CLONOTYPE = \d+ CELLS\n", txt)[1:]
    res = []
    for b in blocks:
        rows, extra = [], []
        for ln in b.split("\n"):
            m = re.match(r"^│(\d+) .*?│(.*)│(.*)│$", ln)
            if m:
                rows.append((m.group(2), m.group(3)))
            elif re.fullmatch(r"[ACGT]{100,}", ln.strip()):
                extra.append(ln.strip())
        hdr = [ln for ln in b.split("\n") if "|IG" in ln]
        names = re.findall(r"\|(IG[A-Z0-9/-]+)", hdr[0]) if hdr else []
        m = re.search(r"^distance = ([0-9.]+)$", b, re.M)
        res.append(dict(rows=rows, seqs=extra, names=names, distance=float(m.group(1)) if m else None))
    return res


def cdr3(cell):
    m = re.search(r"\b(C[A-Z]{4,40}[FW])\b", cell)
    return m.group(1) if m else None


pins = {}
for t, opts, src in ((5, dict(vj_refname=1, cdr3_aa_heavy_pc=80.0, cdr3_aa_light_pc=80.0), "enclone_test2.rs:180"),
                     (7, dict(cdr3_aa_heavy_pc=100.0), "enclone_test2.rs:191"), (8, dict(cdr3_aa_light_pc=100.0), "enclone_test2.rs:197"),
                     (17, dict(vj_refname=1, heavy_pc=96.6), "enclone_test2.rs:251"), (19, dict(vj_refname=1, cdr3_light_len=1), "enclone_test2.rs:265")):
    cl = clonotypes(t)
    pins[f"gtest{t}"] = dict(source=src, golden=f"enclone_gtest{t}_output", opts=opts,
                             clonotypes=[dict(names=c["names"], exacts=[dict(heavy_cdr3=cdr3(h), light_cdr3=cdr3(l)) for h, l in c["rows"]],
                                              vj_seq1=c["seqs"][0] if c["seqs"] else None) for c in cl])
# Case 2 (AGROUP): the goldens print the distance of every member from the center (the first clonotype)
for t, bound, src in ((10, dict(top=2), "enclone_test2.rs:210"), (11, dict(max_dist=14.0), "enclone_test2.rs:216"), (12, dict(max_dist=13.0), "enclone_test2.rs:222")):
    cl = clonotypes(t)
    pins[f"gtest{t}"] = dict(source=src, golden=f"enclone_gtest{t}_output", asym=bound,
                             clonotypes=[dict(names=c["names"], exacts=[dict(heavy_cdr3=cdr3(h), light_cdr3=cdr3(l)) for h, l in c["rows"]],
                                              vj_seq1=None, distance=c["distance"]) for c in cl])
json.dump(pins, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gtest_pins.json"), "w"), indent=1)
for k, v in pins.items():
    print(k, v.get("opts") or v.get("asym"), [(c["names"], [(e["heavy_cdr3"], e["light_cdr3"]) for e in c["exacts"]], len(c["vj_seq1"] or ""), c.get("distance")) for c in v["clonotypes"]])
