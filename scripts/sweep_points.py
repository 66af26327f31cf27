#!/usr/bin/env python
"""BASELINE configs[4]: the band / z-drop sweep -- w in {32, 100, 500} x zdrop in {100, off (32767)} on the
10 %-error pairs: parity of a 20 k-pair sample against the oracle and kernel-only throughput of a 1 M-pair shard."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import genomicsbench_b200 as gb
from oracle.pyoracle import Oracle, make_params

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cfg = gb.gen_named_config("sweep")
pairs, ref, qer = gb.gen_pairs(cfg, 0, n)
O = Oracle()
out = {}
peak = None
for zdrop in (100, 32767):
    with gb.Engine(zdrop=zdrop) as eng:
        if peak is None:
            peak = eng.measure_int_peak()
        for w in (32, 100, 500):
            a = pairs.copy()
            eng.stage(a, ref, qer, w)
            ts = []
            for _ in range(4):
                eng.run_staged(); st = eng.stats(); ts.append(st["ms_kernel"])
            eng.fetch(a)
            want = pairs[:20000].copy()
            O.batch(make_params(zdrop=zdrop), want, ref, qer, w)
            bad = int(sum((a[f][:20000] != want[f]).sum() for f in gb.RESULT_FIELDS))
            ms = min(ts[1:])
            nominal = float((pairs["len1"].astype(np.int64) * pairs["len2"]).sum())
            out[f"w{w}_z{'off' if zdrop == 32767 else zdrop}"] = {
                "ms_kernel": ms, "gcups_eff": st["cells_effective"] / ms / 1e6, "gcups_nom": nominal / ms / 1e6,
                "roofline_frac": st["cells_effective"] * 10 / (ms * 1e-3) / peak, "mismatching_fields_in_20k": bad}
            print(f"w={w} zdrop={zdrop}: {ms:.2f} ms, {out[list(out)[-1]]['gcups_eff']:.0f} GCUPS eff, frac {out[list(out)[-1]]['roofline_frac']:.2f}, mismatches {bad}", flush=True)
json.dump({"pairs": n, "int_peak": peak, "points": out}, open(os.path.join(ROOT, "gpurun_out", "sweep_points.json"), "w"), indent=1)
