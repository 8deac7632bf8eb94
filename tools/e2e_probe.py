#!/usr/bin/env python3
"""Per-step wall time and library breakdown of ejoin_run() on a workload, compact and ASCII forms.
usage: e2e_probe.py [workload] [steps]   (EJOIN_DEBUG=1 adds the per-kernel marks on stderr)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enclone_b200 import synth  # noqa: E402
from enclone_b200.join import CompactPack, Joiner  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C3"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    sp, _ = synth.generate_config(name, 1.0, copy=False, pinned=True)
    t0 = time.perf_counter()
    cp = CompactPack(sp)
    print(f"pack_compact {1e3 * (time.perf_counter() - t0):.1f} ms ({os.cpu_count()} threads)")
    j = Joiner()
    for label, st in (("compact", cp.struct), ("ascii", sp.struct)):
        for i in range(steps):
            t0 = time.perf_counter()
            r = j.run(st, copy=False)
            wall = 1e3 * (time.perf_counter() - t0)
            s = r.stats
            print(f"{label} step {i}: wall {wall:.3f} ms  total {s['ms_total']:.3f}  h2d {s['ms_h2d']:.3f} ({s['h2d_bytes'] / 1e6:.1f} MB)  "
                  f"device {s['ms_device']:.3f}  post {s['ms_host_post']:.3f}  d2h {s['d2h_bytes'] / 1e6:.1f} MB  launches {s['gpu_launches']}", flush=True)
    j.close()


if __name__ == "__main__":
    main()
