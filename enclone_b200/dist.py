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


def init_comm(joiner, rank, world, device=None, group=None):
    """Give `joiner` the library's own communicator: rank 0 draws the id, torch.distributed hands it round."""
    import torch
    import torch.distributed as dist
    from .join import comm_id
    t = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        t.copy_(torch.frombuffer(bytearray(comm_id()), dtype=torch.uint8))
    dist.broadcast(t, 0, group=group)
    joiner.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)


def sharded_join(joiner, pack_struct, rank, world, device=None, group=None, copy=False):
    """The whole join stage on `world` GPUs, one call per rank: score this rank's bucket ranges, gather the emitted joins
    to rank 0 device to device inside the library (ejoin_run_sharded), finish_join there.  Returns the JoinResult of the
    whole input on rank 0 (views into the handle's pinned result buffer unless copy=True), an empty one elsewhere."""
    if getattr(joiner, "comm", None) != (rank, world):
        init_comm(joiner, rank, world, device=device, group=group)
    return joiner.run_sharded(pack_struct, copy=copy)
