#!/usr/bin/env python
"""Stages n pairs of one workload and runs the DP launch `steps` times on the warp-per-pair register kernel: the
command ncu captures for bsw_warp16_kernel.  python scripts/warp_run.py <config> <pairs> [steps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import genomicsbench_b200 as gb  # noqa: E402

name, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pairs, ref, qer = gb.gen_pairs(gb.gen_named_config(name), 0, n)
with gb.Engine(warp_max_pairs=1 << 20) as eng:
    eng.stage(pairs, ref, qer, 100)
    for _ in range(steps):
        eng.run_staged()
        st = eng.stats()
    print(name, n, "pairs: kernel ms %.3f, DP launches %d, effective cells %d" %
          (st["ms_kernel"], st["kernel_launches"], st["cells_effective"]))
