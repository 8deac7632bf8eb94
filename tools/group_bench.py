"""Times the symmetric grouping (egroup_run) on a synthetic repertoire and the CPU oracle on a bounded sample of it.
usage: python tools/group_bench.py [--n 20000] [--families 400] [--opts heavy_pc=96,vj_refname=1] [--cpu-sample 150]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

import numpy as np

from enclone_b200 import group as G


def asym_main(a):
    pk = G.synth_group_pack(a.n, seed=1, n_families=a.families or max(2, a.n // 50))
    cen = np.zeros(a.n, np.uint8); cen[::a.center_step] = 1
    k, v = a.asym.split("=")
    bound = dict(top=int(v)) if k == "top" else dict(max_dist=float(v))
    best, wall = None, []
    for _ in range(a.reps):
        groups, st = G.group_asymmetric(pk, cen, **bound)
        wall.append(round(st["ms_total"], 2))
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    out = dict(asym=bound, n_clono=a.n, n_center=int(cen.sum()), ms_total_per_rep=wall, stats=best, members=int(sum(len(g) for g in groups)),
               pairs_per_s=round(best["pairs_candidate"] / (best["ms_device"] * 1e-3), 1))
    if a.cpu_sample:
        import oracle_lib
        sub = G.synth_group_pack(a.cpu_sample, seed=1, n_families=max(2, a.cpu_sample // 50))
        c2 = np.zeros(a.cpu_sample, np.uint8); c2[::a.center_step] = 1
        t0 = time.time()
        want, ev = oracle_lib.group_asym(sub, c2, G.asym_opts(**bound))
        dt = time.time() - t0
        got, _ = G.group_asymmetric(sub, c2, **bound)
        out["cpu_oracle"] = dict(n_clono=a.cpu_sample, n_center=int(c2.sum()), s=round(dt, 3), pairs_per_s=round(int(c2.sum()) * a.cpu_sample / dt, 1), parity=bool(got == want))
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--families", type=int, default=0)
    ap.add_argument("--opts", default="vj_refname=1,heavy_pc=96")
    ap.add_argument("--cpu-sample", type=int, default=150)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--asym", default="", help="asymmetric grouping instead: 'top=10' or 'max=6'; every --center-step-th clonotype is a center")
    ap.add_argument("--center-step", type=int, default=10)
    a = ap.parse_args()
    if a.asym:
        return asym_main(a)
    kw = {}
    for t in a.opts.split(","):
        k, v = t.split("=")
        kw[k] = float(v) if k in G.DIST_OPTS else int(v)
    t0 = time.time()
    pk = G.synth_group_pack(a.n, seed=1, n_families=a.families or max(2, a.n // 50))
    t_synth = time.time() - t0
    best, wall = None, []
    for _ in range(a.reps):
        g, st = G.group_clonotypes(pk, **kw)
        wall.append(round(st["ms_total"], 2))
        if best is None or st["ms_device"] < best["ms_device"]:
            best = st
    mean_len = float(pk.a["seq_len"].mean())
    out = dict(n_clono=a.n, opts=kw, ms_total_per_rep=wall, synth_s=round(t_synth, 2), stats=best, n_groups=int(len(set(g.tolist()))),
               mean_chain_len=round(mean_len, 1),
               gcups=round(best["chain_compares"] * mean_len * mean_len / (best["ms_device"] * 1e-3) / 1e9, 2) if best["ms_device"] > 0 else None)
    if a.cpu_sample:
        import oracle_lib
        sub = G.synth_group_pack(a.cpu_sample, seed=1, n_families=max(2, a.cpu_sample // 50))
        o, _ = G.make_opts(**kw)
        t0 = time.time()
        want, ev = oracle_lib.group(sub, o)
        dt = time.time() - t0
        got, st2 = G.group_clonotypes(sub, **kw)
        out["cpu_oracle"] = dict(n_clono=a.cpu_sample, s=round(dt, 3), chain_compares=int(ev), compares_per_s=round(ev / dt, 1), parity=bool(np.array_equal(got, want)),
                                 gpu_ms_same_sample=round(st2["ms_device"], 3))
        out["gpu_compares_per_s"] = round(best["chain_compares"] / (best["ms_device"] * 1e-3), 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
