"""The C++ host mirror of the reference interface (enclone_b200/host/join_exacts.hpp):
compiles everywhere (CPU test), runs the known-answer pair through join_exacts() on a GPU."""
import json
import os
import subprocess

import pytest

from kat import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "host_mirror_test")


def _build():
    src = os.path.join(ROOT, "tests", "host_mirror_test.cpp")
    lib = os.path.join(ROOT, "enclone_b200")
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(lib, "host", "join_exacts.hpp")),
                                                                  os.path.getmtime(os.path.join(ROOT, "include", "enclone_join.h")), os.path.getmtime(os.path.join(lib, "libenclone_join.so"))):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", BIN, src, "-L" + lib, "-lenclone_join", "-Wl,-rpath," + lib])
    return BIN


def test_host_mirror_compiles_and_links():
    _build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
def test_host_mirror_runs_the_known_answer_pair(tmp_path):
    kat = load()
    g = kat["germline"]
    reg = kat["regions"]
    refs = [("IGHV3-23*01", g["heavy"]["v"]), ("IGHJ4", g["heavy"]["j"]), ("IGKV2-30", g["light"]["v"]), ("IGKJ1", g["light"]["j"])]
    lines = ["6 4 1"] + [f"{n} {s}" for n, s in refs]
    cells = [kat["k1"], kat["k2"]] * 3
    for i, cell in enumerate(cells):
        lines.append(str(1 if i % 2 == 0 else 0))                      # donors differ inside each pair
        for c, ch in enumerate(("heavy", "light")):
            s = cell[ch]
            vid, jid = (0, 1) if c == 0 else (2, 3)
            regs = [reg["fr1_start"], reg["cdr1_start"], reg["fr2_start"], reg["cdr2_start"], reg["fr3_start"]] if c == 0 else [0, -1, -1, -1, -1]
            lines.append(" ".join(map(str, [s["seq"], s["cdr3"], s["cdr3_start"], vid, jid, refs[vid][1], refs[jid][1], 170] + regs + [3])))
    p = tmp_path / "in.txt"
    p.write_text("\n".join(lines) + "\n")
    out = json.loads(subprocess.check_output([_build(), str(p)]))
    e = kat["expected"]
    assert out["n"] == 6
    # every copy of k1 joins every copy of k2 and identical copies join each other: one orbit
    assert out["norbits"] == 1 and len(out["raw_joins"]) == 5
    j01 = [j for rj, j in zip(out["raw_joins"], out["joins"]) if rj == [0, 1]][0]
    assert (j01["shares"], j01["indeps"], j01["cd"]) == (e["shares"], e["indeps"], e["cd"])
    assert j01["p1"] == float(e["p1"]) and j01["mult"] == float(e["mult"]) and j01["score"] == float(e["score"]) and j01["err"] == 1
