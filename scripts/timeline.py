"""BSW_TIMELINE=1 python scripts/timeline.py [workload] [n]: per-chunk pipeline timeline of bsw_extend on pinned buffers."""
import os, sys, time
os.environ.setdefault("BSW_TIMELINE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import genomicsbench_b200 as gb
from bench import WORKLOADS
name = sys.argv[1] if len(sys.argv) > 1 else "short8"
idx, w, zdrop, desc, default_n = WORKLOADS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
pairs, ref, qer = gb.gen_pairs(gb.gen_named_config(idx), 0, n)
sw = gb.BandedPairWiseSW(6, 1, 6, 1, zdrop, 5, None, 1, 4, 1, devices=[0])
pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
for rep in range(4):
    t0 = time.perf_counter()
    sw.getScores16(pp, pr, pq, n, 1, w)
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.3f} ms", file=sys.stderr)
sw.close()
