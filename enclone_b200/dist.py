"""Multi-GPU decomposition of the join: one process per GPU.

Buckets (runs of equal V..J lengths) are independent, so ranks own contiguous, cost-balanced
bucket ranges and score them with no communication.  The one exchange step is the all-gather of
the per-rank edge lists (torch.distributed: NCCL on GPUs, gloo in the CPU tests); every rank
then runs finish_join over the gathered edges.  When a single bucket dominates the work (the
adversarial giant bucket) ranks instead take all entries and an interleaved share of the tier-1
tiles ("stride mode"), and the ordered replay runs on the gathered raw edges.
"""
import ctypes as C

import numpy as np

from . import _lib
from .join import plan_shards


def stride_mode(pack_struct, world):
    """True when one bucket holds more than 1.25/world of all candidate pairs (same rule as
    ejoin_run_shard)."""
    if world <= 1:
        return False
    _, bstart = plan_shards(pack_struct, world)
    m = np.diff(bstart).astype(np.float64)
    cost = m * (m - 1) / 2
    return bool(cost.max() > 1.25 * cost.sum() / world) if len(cost) else False


def all_gather_rows(arr, device=None, group=None):
    """All-gather of a 2-D uint32 array with a different row count per rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    arr = np.ascontiguousarray(arr, dtype=np.uint32)
    cols = arr.shape[1]
    t = torch.from_numpy(arr.astype(np.int64).reshape(-1))
    if device is not None:
        t = t.to(device)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    buf = torch.zeros(mx, dtype=torch.int64, device=t.device)
    buf[:t.numel()] = t
    outs = [torch.empty(mx, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    parts = [o[:s].cpu().numpy().astype(np.uint32).reshape(-1, cols) for o, s in zip(outs, sizes)]
    return np.concatenate(parts) if parts else np.zeros((0, cols), np.uint32)


def partition(pack_struct, k1, k2):
    """finish_join on the host (ejoin_partition): min-member labels over all info entries."""
    L = _lib.load()
    k1 = np.ascontiguousarray(k1, dtype=np.uint32)
    k2 = np.ascontiguousarray(k2, dtype=np.uint32)
    cls = np.zeros(max(int(pack_struct.n_info), 1), np.uint32)
    ncls = C.c_uint32()
    p = C.POINTER(C.c_uint32)
    rc = L.ejoin_partition(C.byref(pack_struct), k1.ctypes.data_as(p), k2.ctypes.data_as(p), len(k1), cls.ctypes.data_as(p), C.byref(ncls))
    if rc != 0:
        raise _lib.EjoinError(rc, "ejoin_partition")
    return cls[:int(pack_struct.n_info)], int(ncls.value)


def merge_rank_edges(pack_struct, my_pairs, device=None, group=None):
    """Gather every rank's emitted (k1,k2) lists, order them as raw_joins (bucket order =
    ascending k1,k2) and build the global partition."""
    allp = all_gather_rows(my_pairs, device=device, group=group)
    if len(allp):
        order = np.lexsort((allp[:, 1], allp[:, 0]))
        allp = allp[order]
    cls, ncls = partition(pack_struct, allp[:, 0] if len(allp) else np.zeros(0, np.uint32),
                          allp[:, 1] if len(allp) else np.zeros(0, np.uint32))
    return allp, cls, ncls


def all_gather_pairs(pairs, device=None, group=None):
    """All-gather of per-rank (k1,k2) uint32 lists as one byte tensor per rank (two collectives:
    sizes, then the padded payload)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    t = torch.from_numpy(pairs.view(np.uint8).reshape(-1))
    if device is not None:
        t = t.to(device, non_blocking=True)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = torch.zeros(world, dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(sizes, n, group=group)
    sizes = sizes.cpu().tolist()
    mx = max(max(sizes), 8)
    buf = torch.zeros(mx, dtype=torch.uint8, device=t.device)
    buf[:t.numel()] = t
    out = torch.empty(world * mx, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    host = out.cpu().numpy()
    parts = [host[r * mx:r * mx + s].view(np.uint32).reshape(-1, 2) for r, s in enumerate(sizes)]
    return np.concatenate(parts) if parts else np.zeros((0, 2), np.uint32)


def sharded_join(joiner, pack_struct, rank, world, device=None, group=None, gpu_partition=True):
    """The whole join stage on `world` GPUs, one call per rank.  Returns (raw_joins [m,2] for the
    whole input in the reference's order, class_of, this rank's shard stats, this rank's JoinResult)."""
    from ._abi import RAW_EDGE_DTYPE
    mode, fin, raw, st = joiner.run_shard_final(pack_struct, rank, world, copy=False)
    if mode == 1:
        # a giant bucket is shared between ranks: gather the accepted edges, replay once
        rows = np.stack([raw["k1"], raw["k2"], raw["cd"].astype(np.uint32) | (raw["d"].astype(np.uint32) << 16), raw["flags"]], axis=1) \
            if len(raw) else np.zeros((0, 4), np.uint32)
        allrows = all_gather_rows(rows, device=device, group=group)
        allraw = np.zeros(len(allrows), RAW_EDGE_DTYPE)
        if len(allrows):
            allraw["k1"], allraw["k2"] = allrows[:, 0], allrows[:, 1]
            allraw["cd"], allraw["d"], allraw["flags"] = allrows[:, 2] & 0xFFFF, allrows[:, 2] >> 16, allrows[:, 3]
        fin = joiner.finish(pack_struct, allraw)
        pairs = np.stack([fin.edges["k1"], fin.edges["k2"]], axis=1) if len(fin.edges) else np.zeros((0, 2), np.uint32)
        return pairs, fin.class_of, st, fin
    # contiguous bucket ranges: ranks are in bucket order, so concatenating in rank order IS raw_joins order
    mine = np.stack([fin.edges["k1"], fin.edges["k2"]], axis=1) if len(fin.edges) else np.zeros((0, 2), np.uint32)
    allp = all_gather_pairs(mine, device=device, group=group)
    if gpu_partition:
        cls, _ = joiner.partition(pack_struct, allp[:, 0], allp[:, 1])
    else:
        cls, _ = partition(pack_struct, allp[:, 0], allp[:, 1])
    return allp, cls, st, fin
