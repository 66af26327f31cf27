import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, genomicsbench_b200 as gb
for name, n in (("small", 10000), ("small", 2000), ("small", 40000), ("short8", 20000)):
    cfg = gb.gen_named_config(name); pairs, ref, qer = gb.gen_pairs(cfg, 0, n)
    for kw in ({}, {"long_min_qlen": 1}):
        with gb.Engine(**kw) as eng:
            a = pairs.copy(); eng.stage(a, ref, qer, 100)
            ts = []
            for _ in range(6):
                eng.run_staged(); ts.append(eng.stats()["ms_kernel"])
            pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
            e = []
            for _ in range(6):
                t0 = time.perf_counter(); eng.extend(pp, pr, pq, 100); e.append((time.perf_counter() - t0) * 1e3)
            print(name, n, kw, "kernel ms %.3f" % min(ts[1:]), "e2e pinned ms %.3f" % min(e[1:]))
