"""Worker of tests/test_gpu_parity.py::test_two_rank_nccl_*: run under torchrun, one process per GPU.
Every rank joins its share (ejoin_run_sharded: NCCL inside the library); rank 0 compares the gathered result with
the same join on its one GPU and with the CPU oracle.  Prints MGPU_OK on success."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from enclone_b200 import synth
    from enclone_b200._abi import default_params
    from enclone_b200.dist import sharded_join, stride_mode
    from enclone_b200.join import CompactPack, Joiner
    from parity import assert_same_join
    name, scale, want_stride, compact = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    pack, _ = synth.generate_config(name, scale)
    src = CompactPack(pack) if compact else pack
    assert stride_mode(pack.struct, world) == bool(want_stride), "the test input does not exercise the intended mode"
    params = default_params()
    j = Joiner(params, device=local)
    for it in range(2):                                   # twice: a warm handle must give the same answer
        res = sharded_join(j, src.struct, rank, world, device=dev, copy=True)
        if rank == 0:
            j1 = Joiner(params, device=local)
            one = j1.run(src.struct)
            j1.close()
            assert np.array_equal(one.edges, res.edges), "gathered joins differ from the one-GPU join"
            assert np.array_equal(one.class_of, res.class_of), "partition differs from the one-GPU join"
            if it == 0:
                oe, oc, _ = O.join_exacts(pack.struct, params, threads=8)
                assert_same_join(res, oe, oc, what=f"{name} x{scale} on {world} GPUs")
        else:
            assert len(res.edges) == 0
    # an input without entries: every rank is empty, the exchange still has to line up
    empty = pack.subset([])
    res = sharded_join(j, empty.struct, rank, world, device=dev, copy=True)
    assert len(res.edges) == 0 and len(res.class_of) == 0
    # and one with a single two-entry bucket (fewer buckets than ranks x chunks: some ranks get nothing)
    tiny = pack.subset([0, 1])
    res = sharded_join(j, tiny.struct, rank, world, device=dev, copy=True)
    if rank == 0:
        j1 = Joiner(params, device=local)
        one = j1.run(tiny.struct)
        j1.close()
        assert np.array_equal(one.edges, res.edges) and np.array_equal(one.class_of, res.class_of)
    dist.barrier()
    j.close()
    if rank == 0:
        print(f"MGPU_OK {name} x{scale} world={world} stride={want_stride} compact={compact} joins={len(res.edges)}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
