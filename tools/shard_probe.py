#!/usr/bin/env python3
"""Per-rank device timing of the sharded join on ONE GPU: uploads each rank's bucket range in turn
(ejoin_upload_shard) and runs it resident, printing the library's per-kernel times.  Run with
EJOIN_DEBUG=1 to get the per-kernel marks on stderr.  usage: shard_probe.py [world] [workload]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enclone_b200 import synth  # noqa: E402
from enclone_b200.join import Joiner  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    name = sys.argv[2] if len(sys.argv) > 2 else "C3"
    sp, _ = synth.generate_config(name)
    j = Joiner()
    for rank in range(world):
        j.upload_shard(sp.struct, rank, world)
        for _ in range(3):
            st = j.run_resident()
        print(f"rank {rank}/{world}: device {st['ms_device']:.3f} ms  sweep {st['ms_kernel_tier1']:.3f}  pack_t2 {st['ms_kernel_pack_t2']:.3f}  "
              f"scoring {st['ms_kernel_tier2']:.3f}  other {st['ms_kernel_other']:.3f}  compared {st['pairs_compared']}  "
              f"tier2 {st['pairs_tier2']}  accepted {st['pairs_accepted']}", flush=True)
        if os.environ.get("STAGES") and rank in (0, world - 1):
            j.set_stage_timing(True)
            j.run_resident(); j.run_resident()
            print("    stages: " + ", ".join(f"{n} {ms:.3f}" for n, ms in j.stage_times()), flush=True)
            j.set_stage_timing(False)
    j.close()


if __name__ == "__main__":
    main()
