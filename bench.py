#!/usr/bin/env python
"""bench.py — join pairs scored per second on synthetic repertoires (BASELINE.json's metric).

One "step" = one pass of the join stage over the workload.
  value : device-resident step (inputs already in HBM): keys -> sort -> groups -> pack ->
          tier 1 -> tier 2 -> prune -> sort of kept edges, wall clock around K steps.
  e2e   : the same workload through the C-ABI call ejoin_run() with HOST buffers: H2D of
          the flat JoinPack (compact form: V..J bases 2 bits per base as the reference holds
          them in CloneInfo.tigsp, one 64-byte record per entry), the device pipeline (forest,
          two-cell rule, payload, union-find all on the device), D2H of edges and labels —
          the call a user of the library makes.  e2e_variants adds the ASCII / parallel-array
          form of the same input, e2e_cold a fresh process (context, tables, first call).
  pairs : candidate pairs = sum over lens buckets of n_b (n_b - 1) / 2.
Workload: BASELINE.json configs[2] (the ~1.6M-cell pooled BCR run the metric's target
sentence names), which fits one GPU; N > 1 shards its buckets over the ranks (strong scaling).

  python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C3] [--scale S]
With --workload C4 (one 200,000-entry bucket) tier 1 runs on the tensor cores and the roofline entry is k_sweep_mma's (bound: tensor).
Two more paths, each with its own metric (not the north star): --path group (symmetric grouping) and --path refine
(post-join refinement); both take --impl reference for their CPU restatements.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    "C1": "synthetic 123085-like BCR, 1,654 cells, 1 donor",
    "C2": "synthetic 100k-cell BCR repertoire, 1 donor",
    "C3": "synthetic 1.6M-cell 8-donor pooled BCR repertoire",
    "C4": "adversarial single giant bucket, 200k exact subclonotypes",
    "C5": "synthetic 1M-cell 4-donor TCR repertoire",
}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(name, scale, pinned):
    """Generate the synthetic repertoire; with pinned=True the big arenas come from
    ejoin_host_alloc and the smaller arrays are pinned in place (cudaHostRegister)."""
    from enclone_b200 import synth
    L = None
    if pinned:
        from enclone_b200 import _lib
        L = _lib.load()
    cfg = dict(synth.CONFIGS[name])
    cfg["n_cells"] = max(16, int(cfg["n_cells"] * scale))
    t0 = time.time()
    sp, truth = synth.generate(copy=False, pinned=pinned, **cfg)
    gen_s = time.time() - t0
    registered = []
    if pinned:
        st = sp.struct
        n, nx, nr = st.n_info, st.n_exact, st.n_refseq
        from enclone_b200._abi import _PACK_ARRAYS
        import numpy as np
        counts = {"nchains": nx, "ncells": nx, "donor_lo": nx, "donor_hi": nx, "refseq_off": nr, "refseq_len": nr}
        for fname, dt, per_chain in _PACK_ARRAYS:
            cnt = counts.get(fname, n)
            ptrs = [getattr(st, fname)[c] for c in range(2)] if per_chain else [getattr(st, fname)]
            for p in ptrs:
                addr = C.cast(p, C.c_void_p).value
                nbytes = cnt * np.dtype(dt).itemsize
                if addr and nbytes >= (1 << 16) and L.ejoin_host_register(addr, nbytes) == 0:
                    registered.append(addr)
    return sp, gen_s, (L, registered)


def cpu_baseline(sp, params, target_s=12.0, steps=1, warmup=0, best=False):
    """The CPU restatement of the reference join (oracle/), all host threads, on a bounded
    representative sample: every stride-th lens bucket.  Returns (pairs/s, info)."""
    import oracle_lib as O
    threads = os.cpu_count() or 1
    O.lib().oracle_prepare(700)
    # One giant bucket (configuration C4): the reference gives a bucket to ONE rayon task, which at 2e10 pairs would run for hours, so
    # the sample is the first rows of that bucket's triangle, scored by all threads (the oracle's row-parallel mode), reported as such.
    import numpy as np
    from enclone_b200.join import plan_shards
    _, bs = plan_shards(sp.struct, 1)
    sizes = np.diff(bs.astype(np.int64))
    if len(sizes) and sizes.max() >= 50000 and (sizes.max() * (sizes.max() - 1) // 2) > 0.9 * int((sizes * (sizes - 1) // 2).sum()):
        nb = int(sizes.max())
        rows = 256
        try:
            O.set_parallel_bucket(50000, threads)
            while True:
                O.lib().oracle_set_row_limit(rows)
                t0 = time.time()
                O.join_exacts(sp.struct, params, threads=1)
                dt = time.time() - t0
                if dt >= 0.5 * target_s or rows >= nb:
                    break
                rows = min(nb, int(rows * max(2.0, 0.8 * target_s / max(dt, 1e-3))))
        finally:
            O.lib().oracle_set_row_limit(0)
            O.set_parallel_bucket(0, 1)
        rows = min(rows, nb)
        pairs = rows * (nb - 1) - rows * (rows - 1) // 2
        info = {"cores": threads, "kind": "port", "pairs": pairs, "seconds": dt, "stride": 1,
                "sample": f"the first {rows} of {nb} rows of the one bucket's pair triangle ({pairs} candidate pairs), scored row-parallel on {threads} threads "
                          f"in {dt:.2f} s (the reference's own decomposition, one task per bucket, would run this bucket on a single core)"}
        return pairs / dt, info
    # pilot to size the sample
    stride = 64
    t0 = time.time()
    _, _, st = O.join_exacts(sp.struct, params, threads=threads, stride=stride, phase=1)
    dt = max(time.time() - t0, 1e-3)
    est_full = dt * stride
    stride = max(1, int(est_full / target_s))           # the pilot overestimates (few buckets, poor balance): round down
    times, pairs = [], 0
    for it in range(warmup + steps):
        t0 = time.time()
        _, _, st = O.join_exacts(sp.struct, params, threads=threads, stride=stride, phase=0)
        dt = time.time() - t0
        if it >= warmup:
            times.append(dt)
        pairs = st["pairs_candidate"]
    sec = min(times) if best else sum(times) / len(times)     # best-of: the box's host cores are shared, a pass can be disturbed
    info = {"cores": threads, "kind": "port",
            "sample": f"every {stride}-th lens bucket of the workload ({st['n_buckets']} buckets scanned, {pairs} candidate pairs, "
                      f"{st['pairs_evaluated']} evaluated after skip-if-joined, {sec:.2f} s per pass"
                      + (f", best of {len(times)} passes)" if best and len(times) > 1 else ")"),
            "pairs": pairs, "seconds": sec, "stride": stride}
    return pairs / sec, info


def group_main(args):
    """The symmetric grouping (include/enclone_group.h) on a synthetic repertoire: one step = one egroup_run() call with HOST
    buffers (key partitions on the host, uploads, distance passes on the GPU, labels back).  value = candidate clonotype pairs
    per second of device time (CUDA events inside the library around the passes), e2e = per second of wall clock around the call.
    One GPU; not the north-star metric (that is the default --path join)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    from enclone_b200 import group as G
    kw = {}
    for t in args.group_opts.split(","):
        k, v = t.split("=")
        kw[k] = float(v) if k in G.DIST_OPTS else int(v)
    W, K = max(args.warmup, 0), max(args.steps, 1)
    config = {"workload": f"synthetic repertoire of {args.group_n} clonotypes in lineage families (enclone_b200.group.synth_group_pack, seed 1)", "group_opts": kw}
    pk = G.synth_group_pack(args.group_n, seed=1, n_families=max(2, args.group_n // 50))
    mean_len = float(pk.a["seq_len"].mean())
    if args.impl == "reference":
        # grouper.rs is in the reference tree but its crates are not buildable here: the CPU restatement (oracle/grouper_oracle.c,
        # scalar DP distance, one thread) on a bounded sample of the same generator.  The reference itself uses triple_accel's
        # SIMD distance and rayon over groups: this arm understates it by the SIMD and core factors.
        import oracle_lib
        sub = G.synth_group_pack(min(args.group_n, 1500), seed=1, n_families=max(2, min(args.group_n, 1500) // 50))
        o, keep = G.make_opts(**kw)
        ts = []
        for i in range(W + K):
            t0 = time.perf_counter(); g, ev = oracle_lib.group(sub, o); ts.append(time.perf_counter() - t0)
        dt = sum(ts[W:]) / K
        pairs = _group_pairs(sub, g, kw, G)
        val = pairs / dt
        print(json.dumps({"impl": "reference", "path": "group", "metric": "clonotype pairs grouped/sec", "value": val, "unit": "pairs/s", "n_gpus": 1, "steps": K, "warmup": W,
                          "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 bit-vectors + f64 ratio", "data": "synthetic",
                          "config": config, "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": 1, "kind": "port", "sample": f"{sub.n_clono} clonotypes of the same generator, {ev} chain comparisons"},
                          "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the grouping's distance passes have no CPU fallback)")
    sampler = ClockSampler(0)
    for _ in range(W):
        G.group_clonotypes(pk, **kw)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, stats = 0.0, None
    for _ in range(K):
        g, stats = G.group_clonotypes(pk, **kw)
        dev_ms += stats["ms_device"]
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    pairs = stats["pairs_candidate"]
    h2d = int(sum(v.nbytes for k, v in pk.a.items() if not k.startswith("col_")))
    line = {"path": "group", "metric": "clonotype pairs grouped/sec", "value": pairs * K / (dev_ms * 1e-3), "unit": "pairs/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": wall / K * 1e3, "ms_device_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 bit-vectors + f64 ratio", "data": "synthetic", "config": config,
            "e2e": {"value": pairs * K / wall, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * pk.n_clono},
            "gpu_launches": int(stats["gpu_launches"]) * K, "stats": {k: (round(v, 3) if isinstance(v, float) else int(v)) for k, v in stats.items()},
            # the distance kernel is integer-ALU work (about 20 64-bit operations per 64 DP cells), not HBM or tensor work: no
            # hardware ceiling is claimed, the figure is the DP cells its comparisons stand for per second
            "roofline": {"bound": "int_alu", "achieved": stats["chain_compares"] * mean_len * mean_len / (stats["ms_device"] * 1e-3) / 1e9 if stats["ms_device"] > 0 else None,
                         "peak": None, "unit": "GCUPS (cells of the full DP matrices the comparisons stand for)", "frac": None, "traffic": None, "kernel": "k_group_pairs"},
            "clocks": clocks}
    try:
        if args.group_n == 20000 and args.group_opts == "vj_refname=1,heavy_pc=96":
            line["roofline"]["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get("k_group_pairs")
            line["roofline"]["traffic_source"] = "ncu --set full capture of this workload, profiles/r02_traffic.json (mostly the per-thread match vectors spilling through L2: local memory)"
    except Exception:
        pass
    if not args.no_cpu_baseline:
        import oracle_lib
        sub = G.synth_group_pack(min(args.group_n, 1500), seed=1, n_families=max(2, min(args.group_n, 1500) // 50))
        o, keep = G.make_opts(**kw)
        t0 = time.perf_counter(); want, ev = oracle_lib.group(sub, o); dt = time.perf_counter() - t0
        got, st2 = G.group_clonotypes(sub, **kw)
        line["cpu_baseline"] = {"value": _group_pairs(sub, want, kw, G) / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                "sample": f"{sub.n_clono} clonotypes of the same generator, {ev} chain comparisons in {dt:.2f} s (scalar DP distance; the reference uses a SIMD one)",
                                "parity_on_sample": bool(np.array_equal(got, want))}
    print(json.dumps(line))


def refine_main(args):
    """The post-join refinement (include/enclone_refine.h) on a synthetic repertoire: one step = one erefine_run()
    call with HOST buffers (uploads, chain grouping + the three filters with their break-ups on the GPU, results back).
    value = exact subclonotypes per second of device time (CUDA events inside the library), e2e = per second of wall clock
    around the call.  One GPU; not the north-star metric (that is the default --path join)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np
    from enclone_b200 import refine as R
    W, K = max(args.warmup, 0), max(args.steps, 1)
    base = R.synth_refine_pack(args.refine_families, seed=1)
    pk = R.tile_pack(base, args.refine_tiles, pinned=args.impl != "reference")      # page-locked tables and result arrays, as the join's bench uses
    config = {"workload": f"{args.refine_tiles} independent copies of a synthetic repertoire of {args.refine_families} clonal families "
                          f"(enclone_b200.refine.synth_refine_pack, seed 1): {pk.n_exact} exact subclonotypes, {pk.n_chain} chains, {pk.n_info} info entries, "
                          f"{len(pk.a['join_k1'])} raw joins", "filters": "doublet, signature, weak chains (enclone defaults)"}
    in_bytes = int(sum(v.nbytes for v in pk.a.values()))
    out_bytes = int(pk.n_exact * 9 + pk.n_chain * 4)
    if args.impl == "reference":
        import oracle_lib
        sub = R.tile_pack(base, max(1, args.refine_tiles // 4))
        ts = []
        for i in range(W + K):
            t0 = time.perf_counter(); res, st = oracle_lib.refine(sub, oracle_lib.refine_opts()); ts.append(time.perf_counter() - t0)
        dt = sum(ts[W:]) / K
        val = sub.n_exact / dt
        print(json.dumps({"impl": "reference", "path": "refine", "metric": "exact subclonotypes refined/sec", "value": val, "unit": "exacts/s", "n_gpus": 1, "steps": K,
                          "warmup": W, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 / u64 integer", "data": "synthetic",
                          "config": config, "cpu_baseline": {"value": val, "unit": "exacts/s", "cores": 1, "kind": "port",
                                                             "sample": f"{max(1, args.refine_tiles // 4)} of the {args.refine_tiles} copies ({sub.n_exact} exact subclonotypes); the reference itself runs one rayon task per clonotype"},
                          "e2e": {"value": val, "unit": "exacts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the refinement has no CPU fallback)")
    sampler = ClockSampler(0)
    out = R.RefineResult(pk.n_exact, pk.n_chain, pinned=True)
    for _ in range(W):
        R.refine_clonotypes(pk, result=out)
    sampler.start()
    t0 = time.perf_counter()
    dev_ms, stats, res = 0.0, None, None
    for _ in range(K):
        res, stats = R.refine_clonotypes(pk, result=out)
        dev_ms += stats["ms_device"]
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6561.6))
    stages = 1 + sum(1 for k in ("n_doublet", "n_signature", "n_weak") if stats[k])
    alg = (int(stats["h2d_bytes"]) + out_bytes) * stages   # every stage reads the tables once and rewrites labels / columns once
    line = {"path": "refine", "metric": "exact subclonotypes refined/sec", "value": pk.n_exact * K / (dev_ms * 1e-3), "unit": "exacts/s", "n_gpus": 1, "steps": K, "warmup": W,
            "ms_per_step": wall / K * 1e3, "ms_device_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 / u64 integer", "data": "synthetic", "config": config,
            "e2e": {"value": pk.n_exact * K / wall, "unit": "exacts/s", "h2d_bytes_per_step": int(stats["h2d_bytes"]), "d2h_bytes_per_step": out_bytes,
                    "pack_bytes": in_bytes, "note": "tables and result arrays in page-locked memory; of the pack's V..J arena (pageable) only the third chains the rule compares are gathered and uploaded"},
            "gpu_launches": int(stats["gpu_launches"]) * K, "stats": {k: (round(v, 3) if isinstance(v, float) else int(v)) for k, v in stats.items()},
            "roofline": {"bound": "hbm", "achieved": alg / (dev_ms / K * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": alg / (dev_ms / K * 1e-3) / 1e9 / hbm, "traffic": None,
                         "kernel": "whole refinement (no kernel holds a quarter of it: ~20 CUB radix sorts of 64-bit keys and ~70 small kernels)",
                         "algorithmic_bytes_per_step": alg, "note": f"pack + result bytes x {stages} stages (one per filter that deleted something, plus the first); sorts re-read their keys once per 8-bit digit, so the traffic is several times this"},
            "clocks": clocks}
    if not args.no_cpu_baseline:
        import oracle_lib
        sub = R.tile_pack(base, max(1, args.refine_tiles // 4))
        t0 = time.perf_counter(); want, wst = oracle_lib.refine(sub, oracle_lib.refine_opts()); dt = time.perf_counter() - t0
        got, gst = R.refine_clonotypes(sub)
        line["cpu_baseline"] = {"value": sub.n_exact / dt, "unit": "exacts/s", "cores": 1, "kind": "port",
                                "sample": f"{max(1, args.refine_tiles // 4)} of the {args.refine_tiles} copies ({sub.n_exact} exact subclonotypes) in {dt:.2f} s",
                                "parity_on_sample": bool(got.same_as(want))}
    print(json.dumps(line))


def _group_pairs(pk, g, kw, G):
    """candidate pairs of the first distance pass = pairs inside the groups the key partitions leave"""
    keys = {k: v for k, v in kw.items() if k in G.KEY_FLAGS}
    lab, _ = G.group_clonotypes(pk, **keys)                      # key partitions only: host work, no device
    import numpy as np
    _, cnt = np.unique(lab, return_counts=True)
    return int((cnt.astype(np.int64) * (cnt - 1) // 2).sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-from-structs", action="store_true")
    ap.add_argument("--path", default="join", choices=["join", "group", "refine"],
                    help="join: the north-star path (default); group: the symmetric grouping, SURVEY §8(f)3; refine: the post-join refinement, §8(f)2")
    ap.add_argument("--refine-families", type=int, default=3000)
    ap.add_argument("--refine-tiles", type=int, default=60)
    ap.add_argument("--group-n", type=int, default=20000)
    ap.add_argument("--group-opts", default="vj_refname=1,heavy_pc=96")
    args = ap.parse_args()
    if args.path == "group":
        return group_main(args)
    if args.path == "refine":
        return refine_main(args)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = max(args.warmup, 0)
    K = max(args.steps, 1)
    from enclone_b200._abi import default_params
    params = default_params()
    config = {"workload": f"{args.workload}: {WORKLOADS[args.workload]}" + (f" (scale {args.scale})" if args.scale != 1.0 else ""),
              "generator": "tools/synth.cpp, seed 20261019+config id", "join_params": "enclone defaults (SURVEY.md table P)"}

    if args.impl == "reference":
        # The reference's own join is Rust in an un-vendored dependency and cannot be built in this
        # image, so this arm times the CPU restatement (oracle/) on the host cores.  Rank 0 only.
        if rank != 0:
            return
        sp, gen_s, _ = make_workload(args.workload, args.scale, pinned=False)
        per_step = max(3.0, min(20.0, 150.0 / (K + W)))      # whole run within a few minutes; the full workload when it fits
        val, info = cpu_baseline(sp, params, target_s=per_step, steps=K, warmup=W)
        config.update({"n_info": sp.n_info})
        line = {"impl": "reference", "metric": "join pairs scored/sec", "value": val, "unit": "pairs/s", "n_gpus": args.gpus, "steps": K,
                "warmup": W, "ms_per_step": info["seconds"] * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u32 popcount + f64 score", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from enclone_b200.join import Joiner
    from enclone_b200._abi import RAW_EDGE_DTYPE
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the join has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sp, gen_s, pin = make_workload(args.workload, args.scale, pinned=True)
    from enclone_b200.join import CompactPack
    t0 = time.perf_counter()
    cp = CompactPack(sp)                       # what a caller holding CloneInfo.tigsp passes (ejoin_pack_compact, host threads)
    compact_s = time.perf_counter() - t0
    j = Joiner(params, device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- e2e: host buffers -> result, through the C ABI ----------------
    def e2e_step(pack_struct=None):
        pack_struct = pack_struct if pack_struct is not None else cp.struct
        if world == 1:
            return j.run(pack_struct, copy=False)     # views into the library's pinned result buffer
        from enclone_b200.dist import sharded_join
        return sharded_join(j, pack_struct, rank, world, device=dev)   # includes the device-to-device gather of the joins to rank 0

    for _ in range(W):
        res = e2e_step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(K):
        res = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_stats = res.stats
    n_joins = len(res.edges)
    if world > 1:
        nj = torch.tensor([n_joins], device=dev)
        dist.broadcast(nj, 0)
        n_joins = int(nj[0])
    # parity of the sharded result: rank 0 repeats the join on its one GPU and compares raw_joins and the partition
    parity_ok = None
    if world > 1:
        ok = 1
        if rank == 0:
            edges_n = np.array(res.edges, copy=True)
            cls_n = np.array(res.class_of, copy=True)
            j1 = Joiner(params, device=local_rank)
            one = j1.run(cp.struct)
            ok = int(np.array_equal(one.edges, edges_n) and np.array_equal(one.class_of, cls_n))     # every record field, in order
            j1.close()
        okt = torch.tensor([ok], device=dev)
        dist.broadcast(okt, 0)
        parity_ok = bool(int(okt[0]))
    # the same input in its ASCII / parallel-array form (ABI 1 layout), for comparison
    e2e_ascii_s = None
    if world == 1:
        for _ in range(2):
            e2e_step(sp.struct)
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            ra = e2e_step(sp.struct)
        barrier()
        e2e_ascii_s = time.perf_counter() - t0
        ascii_h2d = int(ra.stats["h2d_bytes"])
    # ---------------- value: inputs resident in HBM ----------------
    if world == 1:
        j.upload(cp.struct)
    else:
        j.upload_shard(cp.struct, rank, world)        # this rank's bucket range, resident in HBM
    for _ in range(W):
        st = j.run_resident()
    barrier()
    t0 = time.perf_counter()
    acc, stage_acc = {}, {}
    for _ in range(K):
        st = j.run_resident()
        for k2, v in st.items():
            if k2.startswith("ms_"):
                acc[k2] = acc.get(k2, 0.0) + v
    barrier()
    dev_s = time.perf_counter() - t0
    clocks = sampler.stop()
    # the same step again with a CUDA event between the stages (outside the timed region: the events themselves cost time)
    j.set_stage_timing(True)
    NS = 5
    for _ in range(NS):
        j.run_resident()
        for name, v in j.stage_times():
            stage_acc[name] = stage_acc.get(name, 0.0) + v
    j.set_stage_timing(False)
    tt = torch.tensor([e2e_s, dev_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s, dev_s = float(tt[0]), float(tt[1])
    # pairs of the whole job (all ranks): sum over the lens buckets of the full input of n_b (n_b - 1) / 2
    from enclone_b200.join import plan_shards
    _, bstart = plan_shards(sp.struct, 1)
    m = np.diff(bstart.astype(np.int64))
    pairs_total = int((m * (m - 1) // 2).sum())
    if rank != 0:
        j.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    ms = {k2: v / K for k2, v in acc.items()}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    # Every pipeline stage is timed live with CUDA events on the launching stream (ejoin_stage_time); the roofline entry is for the
    # longest single-kernel stage.  k_sweep is bound by the popcount (XU) pipe, the pack and scoring kernels by memory.
    stages = {k2: v / NS for k2, v in stage_acc.items()}
    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
    xu_peak = 148 * 16 * sm_clock                        # POPC lanes per SM and clock x SMs
    cand = {}
    if stages.get("k_sweep"):
        t = stages["k_sweep"] * 1e-3
        popc = st["alg_bytes_tier1"] / 16.0              # one POPC per 32-base word and compared pair
        cand["k_sweep"] = {"bound": "xu_popc", "kernel": "k_sweep", "achieved": popc / t, "peak": xu_peak, "unit": "POPC/s", "frac": popc / t / xu_peak,
                           "algorithmic_popc_per_launch": popc, "kernel_ms": stages["k_sweep"],
                           "hbm_equivalent": {"achieved": st["alg_bytes_tier1"] / t / 1e9, "peak": peak_hbm, "unit": "GB/s",
                                              "note": "2 x 8W bytes per compared pair as if streamed: rows are staged in shared memory by TMA and reused ~n_group "
                                                      "times on chip, so this exceeds the HBM peak; it only shows HBM is not the binding roof"},
                           "note": "all-pairs CDR3 XOR/popcount; ncu: XU pipe 75 %, ALU 71 %, issue slots 66 % busy, DRAM 2 % (profiles/r02z_ncu_metrics.json)"}
    if stages.get("k_sweep_mma") and st.get("mma_macs"):
        # the large groups of short CDR3s: tier 1 as an int8 contraction on the tensor cores (tcgen05.mma kind::i8).  MEASURED_PEAKS.json has
        # no int8 figure: the roof used is twice its measured dense bf16 rate (the int8 : bf16 ratio of the B200 tensor cores)
        t = stages["k_sweep_mma"] * 1e-3
        peak_i8 = 2.0 * float(peaks.get("bf16_tflops", 1655.3))
        tops = 2.0 * st["mma_macs"] / t / 1e12
        cand["k_sweep_mma"] = {"bound": "tensor", "kernel": "k_sweep_mma", "achieved": tops, "peak": peak_i8, "unit": "TOP/s (int8, dense)", "frac": tops / peak_i8,
                               "algorithmic_ops_per_launch": 2.0 * st["mma_macs"], "kernel_ms": stages["k_sweep_mma"], "pairs": int(st["mma_pairs"]),
                               "pairs_per_s": st["mma_pairs"] / t,
                               "note": "CDR3 Hamming distance as an int8 contraction: 3 values in {+1,-1} per base (tetrahedron vertices), dot = 3L - 4d, exact in s32; "
                                       "128 x 128 x K tile pairs by tcgen05.mma.cta_group::1.kind::i8 into TMEM, survivor masks read back with tcgen05.ld; "
                                       "the time does not depend on K (DESIGN.md 13): the epilogue's 128 compares per thread and tile pair bound it, not the MMA"}
    if stages.get("k_pack_t2"):
        t = stages["k_pack_t2"] * 1e-3
        cand["k_pack_t2"] = {"bound": "hbm", "kernel": "k_pack_t2_2bit (+ k_pack_t2 for the entries with an ASCII chain)", "achieved": st["alg_bytes_pack_t2"] / t / 1e9,
                             "peak": peak_hbm, "unit": "GB/s", "frac": st["alg_bytes_pack_t2"] / t / 1e9 / peak_hbm,
                             "algorithmic_bytes_per_launch": int(st["alg_bytes_pack_t2"]), "kernel_ms": stages["k_pack_t2"],
                             "note": "2-bit V..J words in (8 B per 32 bases), H/L/M plane rows out (12 B per 32 bases), needed entries only; ncu: issue 36 %, "
                                     "19 of 32 lanes, DRAM 16 %: bound by the bit transpose and the dependent meta -> offset -> words loads, not by HBM"}
    dom = max(cand, key=lambda k2: cand[k2]["kernel_ms"]) if cand else None
    roofline = dict(cand[dom]) if dom else {"bound": "hbm", "achieved": 0.0, "peak": peak_hbm, "unit": "GB/s", "frac": None}
    roofline["peak_source"] = ("MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)") if roofline["bound"] == "hbm" \
        else ("2 x MEASURED_PEAKS.json bf16_tflops (burst): int8 dense runs at twice the bf16 rate" if roofline["bound"] == "tensor"
              else "16 POPC lanes per SM and clock x 148 SMs x the SM clock sampled under load")
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tj.get("workload") == args.workload and args.scale == 1.0 and world == 1:
            traffic = tj.get(dom)
    except Exception:
        pass
    roofline["traffic"] = traffic
    roofline["traffic_source"] = "ncu --set full capture, profiles/r02_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum per launch)" if traffic else None
    roofline["other_kernels"] = {k2: {kk: v for kk, v in c.items() if kk in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms")} for k2, c in cand.items() if k2 != dom}
    roofline["stage_ms"] = {k2: round(v, 4) for k2, v in stages.items()}
    t2ms = max(ms.get("ms_kernel_tier2", 0), 1e-9)
    roofline["scoring_group"] = {"kernels": "k_phase_a, k_phaseb_split, k_tier2 x2, k_tier2_generic x2", "ms": ms.get("ms_kernel_tier2", 0),
                                 "algorithmic_bytes": int(st["alg_bytes_tier2"]), "achieved_gbs": st["alg_bytes_tier2"] / t2ms / 1e6,
                                 "frac_of_hbm": st["alg_bytes_tier2"] / t2ms / 1e6 / peak_hbm,
                                 "note": "latency-bound gathers of cold rows (two 128-byte records + two t2 rows per scored pair)"}
    roofline["device_ms_per_step"] = ms.get("ms_device")
    config.update({"n_info": sp.n_info})
    workload_stats = {"candidate_pairs": pairs_total, "pairs_compared_tier1": int(st["pairs_compared"]),
                      "pairs_tier2": int(st["pairs_tier2"]), "pairs_accepted": int(st["pairs_accepted"]), "joins_emitted": int(n_joins),
                      "l2": "inputs (%.0f MB resident) exceed the 126 MB L2; no flush" % (e2e_stats["h2d_bytes"] / 1e6),
                      "host_buffers": "pinned; compact pack (EJOIN_F_TIGS_2BIT | EJOIN_F_ENTRY_RECS | EJOIN_F_TIGS_IN_PLACE): the 2-bit V..J "
                                      "words are read in place over PCIe for the entries that need them, h2d_bytes_per_step counts those bytes",
                      "generate_s": round(gen_s, 2), "pack_compact_s": round(compact_s, 3), "pack_compact_threads": os.cpu_count(),
                      "parallelism": "1 GPU" if world == 1 else f"{world} GPUs, cost-balanced bucket ranges, edges gathered to rank 0 over NCCL"}
    line = {"metric": "join pairs scored/sec", "value": pairs_total * K / dev_s, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 popcount + f64 score", "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": pairs_total * K / e2e_s, "unit": "pairs/s", "ms_per_step": e2e_s / K * 1e3,
                    "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]), "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]),
                    "breakdown_ms": {k2: round(e2e_stats[k2], 3) for k2 in ("ms_h2d", "ms_device", "ms_d2h", "ms_host_post", "ms_total")
                                     if k2 in e2e_stats}},
            "gpu_launches": int(st["gpu_launches"]) * K, "roofline": roofline, "workload_stats": workload_stats,
            "pairs_compared_per_s": int(st["pairs_compared"]) * K / dev_s}
    if parity_ok is not None:
        line["parity_ok"] = parity_ok
    if e2e_ascii_s is not None:
        line["e2e_variants"] = {"ascii_parallel_arrays": {"ms_per_step": e2e_ascii_s / K * 1e3, "value": pairs_total * K / e2e_ascii_s,
                                                          "h2d_bytes_per_step": ascii_h2d}}
    # the same join from the caller's array-of-struct inputs, and the cold first call of a fresh process (tools/from_structs.cpp)
    fs = os.path.join(ROOT, "tools", "from_structs")
    if world == 1 and os.path.exists(fs) and not args.no_from_structs:
        try:
            out = subprocess.run([fs, args.workload, str(args.scale), "3", "1"], capture_output=True, text=True, timeout=240)
            r = json.loads(out.stdout.strip().splitlines()[-1])
            line["e2e_from_structs"] = {"ms_per_step": r["warm"]["ms_best"], "value": pairs_total / (r["warm"]["ms_best"] * 1e-3), "unit": "pairs/s",
                                        "breakdown_ms": {"flatten (Vec<CloneInfo> -> compact pack)": r["warm"]["ms_pack"], "ejoin_run": r["warm"]["ms_run"],
                                                         "raw_joins + EquivRel": r["warm"]["ms_unpack"]},
                                        "host_threads": r["threads"], "note": "enclone_b200/host/join_exacts.hpp: AoS structs in, raw_joins + EquivRel out; handle "
                                        "created beforehand; best of 3"}
            line["e2e_cold"] = {"ms_total": r["cold"]["ms_total"], "breakdown_ms": {"ejoin_create (CUDA context, streams)": r["cold"]["ms_create"],
                                                                                   "flatten incl. first pinned allocations": r["cold"]["ms_pack"],
                                                                                   "first ejoin_run (p1 / mult tables, device arenas)": r["cold"]["ms_run"],
                                                                                   "raw_joins + EquivRel": r["cold"]["ms_unpack"]},
                                "value": pairs_total / (r["cold"]["ms_total"] * 1e-3), "unit": "pairs/s",
                                "note": "fresh process, first call: join_exacts runs once per enclone process, so this is the latency a user sees unless the "
                                        "handle is created at program start (INTEGRATION.md)"}
        except Exception as e:                       # the tool is optional evidence, never a reason to lose the bench line
            line["e2e_from_structs"] = {"unavailable": str(e)[:200]}
    if not args.no_cpu_baseline:
        val, info = cpu_baseline(sp, params, steps=2, best=True)
        line["cpu_baseline"] = {"value": val, "unit": "pairs/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]}
    print(json.dumps(line), flush=True)
    j.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
