"""Warp-per-pair register kernel (bsw_warp16.cuh) against the thread-per-pair kernel on calls that do not fill the GPU:
resident kernel time (bsw_run_staged) and the page-locked end-to-end call per batch size, and the latency route
(bsw_params.tiny_batch) with and without it.  Usage: python scripts/warp_probe.py [out.json]"""
import json
import sys
import time

sys.path.insert(0, '/root/repo')
import numpy as np
import genomicsbench_b200 as gb

out = {"rows": [], "latency_route": []}
for name, n in (("small", 512), ("small", 1024), ("small", 2048), ("small", 4096), ("small", 8192), ("small", 16384),
                ("short8", 2048), ("short8", 8192), ("large", 2048), ("long16", 2048)):
    cfg = gb.gen_named_config(name)
    pairs, ref, qer = gb.gen_pairs(cfg, 0, n)
    want = None
    for wmax in (-1, 1 << 20):
        with gb.Engine(warp_max_pairs=wmax) as eng:
            a = pairs.copy()
            eng.stage(a, ref, qer, 100)
            ts = []
            for _ in range(8):
                eng.run_staged()
                ts.append(eng.stats()["ms_kernel"])
            eng.fetch(a)
            res = np.stack([a[f] for f in gb.RESULT_FIELDS], axis=1)
            if want is None:
                want = res
            same = bool(np.array_equal(res, want))
            pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
            e = []
            for _ in range(8):
                t0 = time.perf_counter()
                eng.extend(pp, pr, pq, 100)
                e.append((time.perf_counter() - t0) * 1e3)
            cells = eng.stats()["cells_effective"]
            row = {"config": name, "pairs": n, "kernel": "thread-per-pair" if wmax < 0 else "warp-per-pair",
                   "kernel_ms": round(min(ts[2:]), 4), "e2e_pinned_ms": round(min(e[2:]), 4),
                   "gcups_eff": round(cells / min(ts[2:]) / 1e6, 1), "same_results": same}
            out["rows"].append(row)
            print(row, flush=True)
cfg = gb.gen_named_config("small")
pairs, ref, qer = gb.gen_pairs(cfg, 0, 1536)
for n in (128, 512, 1536):
    for wmax in (-1, 0):
        with gb.Engine(tiny_batch=1536, warp_max_pairs=wmax) as eng:
            a = pairs[:n].copy()
            e = []
            for _ in range(40):
                t0 = time.perf_counter()
                eng.extend(a, ref, qer, 100)
                e.append((time.perf_counter() - t0) * 1e3)
            st = eng.stats()
            row = {"pairs": n, "kernel": "warp-per-pair, rows in shared memory (32-bit)" if st["n_long"] else "warp-per-pair, rows in registers",
                   "call_ms_pageable": round(float(np.median(e[5:])), 4), "kernel_launches": st["kernel_launches"]}
            out["latency_route"].append(row)
            print(row, flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
