"""Quick GPU sanity run: parity of every named config (sampled) against the oracle, the measured
integer peak, and kernel-only timing.  Usage: python scripts/gpu_check.py [n_parity] [n_timing] [short_variant]"""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import genomicsbench_b200 as gb
from oracle.pyoracle import Oracle, make_params

n_par = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n_tim = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # 0 = packed 16-bit short kernel, 1 = 32-bit
orc = Oracle()
P = make_params()
eng = gb.Engine(short_variant=variant)
peak = eng.measure_int_peak()
print("int peak lane-ops/s: %.4g" % peak, flush=True)
out = {"int_peak": peak, "short_variant": variant, "configs": {}}
for name in ["small", "short8", "long16", "large", "sweep"]:
    cfg = gb.gen_named_config(name)
    pairs, r, q = gb.gen_pairs(cfg, 0, min(n_par, cfg.n_pairs))
    a = pairs.copy(); b = pairs.copy()
    orc.batch(P, a, r, q, 100)
    eng.extend(b, r, q, 100)
    bad = {f: int((a[f] != b[f]).sum()) for f in gb.RESULT_FIELDS}
    nbad = int(np.any([a[f] != b[f] for f in gb.RESULT_FIELDS], axis=0).sum())
    print(name, "parity mismatching pairs:", nbad, bad, flush=True)
    if nbad:
        idx = np.nonzero(np.any([a[f] != b[f] for f in gb.RESULT_FIELDS], axis=0))[0][:5]
        for k in idx:
            print("  pair", k, "len1", pairs['len1'][k], "len2", pairs['len2'][k], "h0", pairs['h0'][k],
                  "oracle", [int(a[f][k]) for f in gb.RESULT_FIELDS], "gpu", [int(b[f][k]) for f in gb.RESULT_FIELDS])
    # timing
    nt = min(n_tim, cfg.n_pairs)
    pairs, r, q = gb.gen_pairs(cfg, 0, nt)
    t0 = time.time(); eng.stage(pairs, r, q, 100); t_stage = time.time() - t0
    best = 1e30
    for rep in range(4):
        eng.run_staged()
        st = eng.stats()
        best = min(best, st["ms_kernel"])
    t0 = time.time(); eng.fetch(pairs); t_fetch = time.time() - t0
    eng.extend(pairs, r, q, 100)
    t0 = time.time(); eng.extend(pairs, r, q, 100); t_e2e = time.time() - t0
    st2 = eng.stats()
    pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(r), gb.pinned_copy(q)
    eng.extend(pp, pr, pq, 100)
    t0 = time.time(); eng.extend(pp, pr, pq, 100); t_pin = time.time() - t0
    st3 = eng.stats()
    same = all(np.array_equal(pp[f], pairs[f]) for f in gb.RESULT_FIELDS)
    cells = st["cells_effective"]
    rec = dict(n=nt, ms_kernel=best, gcups_eff=cells / best / 1e6, gcups_nom=st["cells_nominal"] / best / 1e6,
               mpairs_s=nt / best / 1e3, roofline_frac=(cells * 10 / (best * 1e-3)) / peak if peak else None,
               launches=st["kernel_launches"], stage_ms=t_stage * 1e3, fetch_ms=t_fetch * 1e3, e2e_ms=t_e2e * 1e3,
               e2e_pinned_ms=t_pin * 1e3, pinned_equals_pageable=bool(same), pinned_launches=st3["kernel_launches"],
               e2e_stats={k: st2[k] for k in ("ms_sort", "ms_pack", "ms_h2d", "ms_kernel", "ms_d2h", "ms_scatter", "ms_total")})
    out["configs"][name] = rec
    print(name, json.dumps(rec), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/gpu_check_v{variant}.json").write_text(json.dumps(out, indent=1))
