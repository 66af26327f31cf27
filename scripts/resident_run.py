#!/usr/bin/env python
"""Stages one workload and runs its DP kernels `steps` times (nothing else): the command ncu captures.
python scripts/resident_run.py <workload> [pairs] [steps]   (workload = a key of bench.WORKLOADS)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import genomicsbench_b200 as gb  # noqa: E402
from bench import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "short8"
idx, w, zdrop, desc, n0 = WORKLOADS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 and int(sys.argv[2]) > 0 else min(n0, 1_000_000)
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pairs, ref, qer = gb.gen_pairs(gb.gen_named_config(idx), 0, n)
with gb.Engine(zdrop=zdrop) as eng:
    eng.stage(pairs, ref, qer, w)
    for _ in range(steps):
        eng.run_staged()
        st = eng.stats()
    print(name, n, "pairs, w", w, "zdrop", zdrop, ": kernel ms %.3f, DP launches per step %d, effective cells %d" %
          (st["ms_kernel"], st["kernel_launches"], st["cells_effective"]))
