#!/usr/bin/env python3
"""The device-resident join step on a workload, three times — the short command the ncu captures wrap.
usage: prof_step.py [workload] [runs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enclone_b200 import synth  # noqa: E402
from enclone_b200.join import CompactPack, Joiner  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    runs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    sp, _ = synth.generate_config(name, 1.0, copy=False)
    cp = CompactPack(sp)
    j = Joiner()
    j.upload(cp.struct)
    for _ in range(runs):
        st = j.run_resident()
    print(f"{name}: device {st['ms_device']:.3f} ms, sweep {st['ms_kernel_tier1']:.3f}, pack_t2 {st['ms_kernel_pack_t2']:.3f}, scoring {st['ms_kernel_tier2']:.3f}, "
          f"launches {st['gpu_launches']}, joins {st['edges_after_prune']}")
    j.close()


if __name__ == "__main__":
    main()
