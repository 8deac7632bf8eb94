#!/usr/bin/env python
"""Build tests/golden/flaky_packs.npz: JoinPacks made by enclone_b200/ingest.py from the reference's own small datasets
(/root/reference/enclone_exec/testx/inputs/flaky*/outs/all_contig_annotations.json; used by regression tests 244-261,
enclone_testlist/src/main_testlist.rs:589-639).  The JSON files only exist in the build container, so the packs are committed."""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from enclone_b200.ingest import load_datasets  # noqa: E402

REF = "/root/reference/enclone_exec/testx/inputs"


def flatten(pack):
    out = {}
    for k, v in pack.arrays.items():
        if isinstance(v, list):
            out[k + "__0"], out[k + "__1"] = v[0], v[1]
        else:
            out[k] = v
    return out


def main():
    sets = {}
    for d in sorted(glob.glob(REF + "/flaky*")):
        name = os.path.basename(d)
        p = os.path.join(d, "outs", "all_contig_annotations.json")
        if os.path.exists(p) and name not in ("flaky8a", "flaky8b"):
            sets[name] = ([p], [0], name != "flaky6")          # flaky6 is the TCR dataset (main_testlist.rs:605)
    sets["flaky8ab"] = ([REF + "/flaky8b/outs/all_contig_annotations.json", REF + "/flaky8a/outs/all_contig_annotations.json"], [1, 0], True)
    out = {}
    for name, (paths, donors, is_bcr) in sets.items():
        pack, info = load_datasets(paths, is_bcr=is_bcr, donors=donors)
        for k, v in flatten(pack).items():
            out[f"{name}/{k}"] = v
        out[f"{name}/is_bcr"] = np.array([int(is_bcr)])
        print(name, "entries", pack.n_info, "exact subclonotypes", len(pack.arrays["nchains"]), "germlines", len(pack.arrays["refseq_len"]))
    np.savez_compressed(os.path.join(HERE, "flaky_packs.npz"), **out)


if __name__ == "__main__":
    main()
