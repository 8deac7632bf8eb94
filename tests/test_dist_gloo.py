"""World-size-2 (and 3) host-side test of the multi-rank path on gloo: bucket ownership,
the variable-length all-gather of edge lists and the rank-0-style merge (finish_join).
The per-rank edge producer here is the CPU oracle restricted to the rank's buckets — the GPU
version of the same flow is tests/test_gpu_parity.py::test_sharded_run_equals_single."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    from enclone_b200 import synth
    from enclone_b200._abi import default_params
    from enclone_b200.dist import merge_rank_edges
    from enclone_b200.join import plan_shards
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pack, _ = synth.generate(n_cells=9000, n_donors=2, seed=77)
    owner, bstart = plan_shards(pack.struct, world)
    mine = np.flatnonzero(owner == rank)
    edges = np.zeros((0, 2), np.uint32)
    if len(mine):
        # the rank's buckets come as a few runs of consecutive buckets (chunks dealt in turn, ejoin_plan_shards)
        cuts = np.flatnonzero(np.diff(mine) != 1) + 1
        parts = []
        for run in np.split(mine, cuts):
            e, _, _ = O.join_exacts(pack.struct, default_params(), bucket_lo=int(run[0]), bucket_hi=int(run[-1]) + 1)
            parts.append(np.stack([e["k1"], e["k2"]], axis=1).astype(np.uint32))
        edges = np.concatenate(parts)
    allp, cls, ncls = merge_rank_edges(pack.struct, edges)
    if rank == 0:
        fe, fc, _ = O.join_exacts(pack.struct, default_params())
        ok = np.array_equal(allp[:, 0], fe["k1"]) and np.array_equal(allp[:, 1], fe["k2"]) and np.array_equal(cls, fc)
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_merge_on_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    ok = out.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
    assert ok and all(p.exitcode == 0 for p in procs)
