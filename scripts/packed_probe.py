#!/usr/bin/env python
"""End-to-end time of the packed route (bsw_extend_packed) next to the SeqPair route (bsw_extend) and the resident
kernels, per workload.  python scripts/packed_probe.py [workload ...] [--n PAIRS] [--steps K]"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import genomicsbench_b200 as gb  # noqa: E402

W = {"small": (0, 10_000), "short8": (1, 1_000_000), "long16": (2, 1_000_000), "large": (3, 1_000_000), "sweep": (4, 1_000_000)}


def timeit(fn, steps, warm=3):
    for _ in range(warm):
        fn()
    t = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t)), float(np.min(t))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="*", default=["short8"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--w", type=int, default=100)
    ap.add_argument("--only-packed", action="store_true")
    ap.add_argument("--only-resident", action="store_true")
    ap.add_argument("--long", action="store_true", help="route every pair to the warp-per-pair kernel")
    args = ap.parse_args()
    res = {}
    for name in args.workloads:
        idx, n0 = W[name]
        n = args.n or n0
        cfg = gb.gen_named_config(idx)
        pairs, ref, qer = gb.gen_pairs(cfg, 0, n)
        nominal = float((pairs["len1"].astype(np.int64) * pairs["len2"]).sum())
        t0 = time.perf_counter()
        b = gb.PackedBatch.from_pairs(pairs, ref, qer, pinned=True)
        t_pack = (time.perf_counter() - t0) * 1e3
        pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
        out = gb.pinned_empty(n, gb.OUTSCORE_DTYPE)
        out16 = gb.pinned_empty(n, gb.SCORE16_DTYPE)
        if args.only_resident:
            with gb.Engine(**(dict(long_min_qlen=1) if args.long else {})) as eng:
                eng.stage(pairs, ref, qer, args.w)
                ks = []
                for _ in range(args.steps + 3):
                    eng.run_staged()
                    ks.append(eng.stats()["ms_kernel"])
                print(name, "resident kernel ms median %.3f min %.3f launches %d" % (np.median(ks[3:]), np.min(ks[3:]), eng.stats()["kernel_launches"]), flush=True)
            continue
        if args.only_packed:
            with gb.Engine() as eng:
                ms_packed, mn_packed = timeit(lambda: eng.extend_packed(b, args.w, out=out), args.steps)
                print(name, "packed ms median %.3f min %.3f" % (ms_packed, mn_packed), flush=True)
            continue
        with gb.Engine() as eng:
            eng.stage(pairs, ref, qer, args.w)
            ms_res, _ = timeit(lambda: eng.run_staged(), args.steps)
            ms_k = eng.stats()["ms_kernel"]
            ms_packed, mn_packed = timeit(lambda: eng.extend_packed(b, args.w, out=out), args.steps)
            st = eng.stats()
            ms_p16, mn_p16 = timeit(lambda: eng.extend_packed(b, args.w, out=out16, compact=True), args.steps)
            ms_ext, mn_ext = timeit(lambda: eng.extend(pp, pr, pq, args.w), args.steps)
            st2 = eng.stats()
        same = all(np.array_equal(out[f], pp[f]) for f in gb.RESULT_FIELDS)
        res[name] = dict(pairs=n, ms_kernel_resident=ms_k, ms_packed=ms_packed, ms_packed_min=mn_packed, ms_packed16=ms_p16,
                         ms_extend_pinned=ms_ext, ms_extend_min=mn_ext,
                         gcups_packed=nominal / ms_packed / 1e6, gcups_extend=nominal / ms_ext / 1e6, gcups_resident=nominal / ms_k / 1e6,
                         bytes_per_pair_packed=(st["h2d_bytes"] + st["d2h_bytes"]) / n,
                         bytes_per_pair_extend=(st2["h2d_bytes"] + st2["d2h_bytes"]) / n, host_pack_ms=t_pack,
                         launches_packed=st["kernel_launches"], results_equal=bool(same))
        print(name, json.dumps(res[name]), flush=True)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "packed_probe.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
