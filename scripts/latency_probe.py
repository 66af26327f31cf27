"""Per-call latency of getScores16 for small batches (the reference driver's -b 512 habit, scripts/run-cpu.sh:30)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, genomicsbench_b200 as gb
cfg = gb.gen_named_config("small")
allp, ref, qer = gb.gen_pairs(cfg, 0, 65536)
pr, pq = gb.pinned_copy(ref), gb.pinned_copy(qer)
sw = gb.BandedPairWiseSW(6, 1, 6, 1, 100, 5, None, 1, 4, 1, devices=[0])
for n in (512, 4096, 16384, 65536):
    pairs = allp[:n].copy(); pp = gb.pinned_copy(pairs)
    for label, args in (("pageable", (pairs, ref, qer)), ("pinned", (pp, pr, pq))):
        for _ in range(5): sw.getScores16(*args, n, 1, 100)
        t0 = time.perf_counter(); reps = 30
        for _ in range(reps): sw.getScores16(*args, n, 1, 100)
        dt = (time.perf_counter() - t0) / reps
        print(f"n={n:6d} {label:8s} {dt*1e3:7.3f} ms per call  {n/dt/1e6:7.2f} M pairs/s")
sw.close()
