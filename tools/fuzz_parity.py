#!/usr/bin/env python3
"""Randomized GPU-vs-oracle parity sweep (the loop behind tests/test_gpu_parity.py::test_randomized_repertoires_and_parameters,
with more cases).  usage: fuzz_parity.py [first_case] [n_cases]   — prints one line per case, exits 1 on the first mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from enclone_b200 import synth  # noqa: E402
from enclone_b200._abi import default_params  # noqa: E402
from enclone_b200.join import Joiner  # noqa: E402
from parity import assert_same_join  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    for case in range(first, first + count):
        rng = np.random.default_rng(5000 + case)
        gen = dict(n_cells=int(rng.integers(500, 40000)), n_donors=int(rng.integers(1, 6)), is_bcr=case % 3 != 2, seed=int(rng.integers(1, 1 << 30)),
                   frac_naive=float(rng.uniform(0.05, 0.9)), frac_public=float(rng.uniform(0.0, 0.2)), frac_indel=float(rng.uniform(0.0, 0.1)),
                   frac_3chain=float(rng.uniform(0.0, 0.1)), frac_1chain=float(rng.uniform(0.0, 0.1)), family_alpha=float(rng.uniform(1.1, 2.8)))
        if rng.random() < 0.35:
            gen["max_family"] = int(rng.integers(200, 6000))
        if rng.random() < 0.15:
            gen["mode"] = 1                                   # one giant bucket
            gen["n_cells"] = int(rng.integers(500, 6000))
        over = {}
        if rng.random() < 0.5: over["mix_donors"] = 1
        if rng.random() < 0.3: over["max_score"] = float(10.0 ** rng.uniform(0, 7))
        if rng.random() < 0.3: over["auto_share"] = int(rng.integers(3, 40))
        if rng.random() < 0.3: over["join_cdr3_ident"] = float(rng.uniform(70, 100))
        if rng.random() < 0.3: over["max_cdr3_diffs"] = int(rng.integers(0, 12))
        if rng.random() < 0.3: over["cdr3_mult"] = int(rng.integers(1, 8))
        if rng.random() < 0.3: over["comp_filt"] = int(rng.integers(0, 20))
        if rng.random() < 0.3: over["max_degradation"] = int(rng.integers(0, 5))
        if rng.random() < 0.2: over["easy"] = 1
        if rng.random() < 0.2: over["fwr1_cdr12_delta"] = float(rng.uniform(0, 40))
        if rng.random() < 0.2: over["max_diffs"] = int(rng.integers(5, 120))
        if rng.random() < 0.2: over["ref_v_trim"] = int(rng.integers(0, 20)); over["ref_j_trim"] = int(rng.integers(0, 20))
        pack, _ = synth.generate(**gen)
        params = default_params(**over)
        j = Joiner(params)
        g1 = j.run(pack.struct)
        g2 = j.run(pack.struct)                               # same handle, second run: must be identical
        j.close()
        oe, oc, ost = O.join_exacts(pack.struct, params, threads=os.cpu_count() or 8)
        assert_same_join(g1, oe, oc, max_score=over.get("max_score", 100000.0), what=f"case {case}: {gen} {over}")
        assert np.array_equal(g1.edges, g2.edges) and np.array_equal(g1.class_of, g2.class_of), f"case {case}: not deterministic"
        print(f"case {case}: n={pack.n_info} pairs={ost['pairs_candidate']} scored={g1.stats['pairs_tier2']} joins={len(g1.edges)} "
              f"generic={g1.stats['pairs_generic']} ok", flush=True)


if __name__ == "__main__":
    main()
