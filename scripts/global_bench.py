#!/usr/bin/env python
"""Throughput of bsw_global (banded global alignment + CIGAR) next to the reference's own
ksw_global2 on one host thread (GLOBAL_BENCH_NO_CPU=1 skips that leg; BSW_GLOBAL_KERNEL=1 / 2w selects the first kernel /
the second kernel's 64-bit slots for A/B runs).  Workload: the pairs of tests/golden/global/global_default.npz tiled T
times; results are checked against the tiled golden.  Usage: python scripts/global_bench.py [tiles=300] [reps=5]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import genomicsbench_b200 as gb
from test_global import load_global_case, _pairs_of
from oracle.pyoracle import KswReference, make_params

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 300
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
c = load_global_case("global_default"); z = c["z"]
pairs, ref, qer = _pairs_of(gb, c)
big, wbig = np.tile(pairs, tiles), np.tile(z["w"], tiles)
const_w = int(os.environ.get("GLOBAL_BENCH_W", "0"))      # one band for all alignments (one launch class; for profiling): results are not compared
if const_w:
    wbig = np.full(len(big), max(const_w, int(np.abs(pairs["len1"] - pairs["len2"]).max())), np.int32)
eng = gb.Engine(**c["P"])
best, st = 1e9, None
for r in range(reps + 1):
    t0 = time.perf_counter()
    score, cigar, off = eng.global_align(big, ref, qer, wbig)
    dt = time.perf_counter() - t0
    if r: best = min(best, dt)
    st = eng.stats()
assert const_w or (np.array_equal(score, np.tile(z["score"], tiles)) and np.array_equal(cigar, np.tile(z["cigar"], tiles)))
# the same call with caller-owned result arrays reused across calls, as a C caller would hold them (fresh numpy arrays
# cost the first touch of every page inside the call and a copy of the operation list behind it)
out = (np.zeros(len(big), np.int32), np.zeros(len(big), np.int32), np.zeros(len(cigar) + 16, np.uint32), np.zeros(len(big) + 1, np.int64))
best_reused, st_reused = 1e9, st
for r in range(reps + 1 if reps else 0):
    t0 = time.perf_counter()
    score2, cigar2, off2 = eng.global_align(big, ref, qer, wbig, out=out)
    dt = time.perf_counter() - t0
    if r and dt < best_reused: best_reused, st_reused = dt, eng.stats()
if reps:
    assert np.array_equal(score2, score) and np.array_equal(cigar2, cigar) and np.array_equal(off2, off)
line = {"metric": "global_alignments_per_sec", "alignments": int(len(big)), "seconds": best, "value": len(big) / best,
        "band_cells": int(st["cells_effective"]), "gcups_band": st["cells_effective"] / best / 1e9,
        "kernel_ms": st["ms_kernel"], "gcups_band_kernel_only": st["cells_effective"] / (st["ms_kernel"] * 1e-3) / 1e9,
        "host_ms": {"prepare_and_enqueue": st["ms_pack"], "wait_for_device": st["ms_d2h"], "results_out": st["ms_scatter"], "call_total": st["ms_total"]},
        "seconds_reused_result_arrays": best_reused if reps else None,
        "value_reused_result_arrays": len(big) / best_reused if reps else None,
        "host_ms_reused_result_arrays": {"prepare_and_enqueue": st_reused["ms_pack"], "wait_for_device": st_reused["ms_d2h"],
                                         "results_out": st_reused["ms_scatter"], "call_total": st_reused["ms_total"], "kernel_ms": st_reused["ms_kernel"]},
        "cigar_ops": int(len(cigar)), "gpu_launches": int(st["kernel_launches"]), "buffers": "pageable numpy arrays",
        "kernel": {"": "second (bsw_global2.cuh), 16-bit slots", "1": "first (bsw_global.cuh)", "2w": "second, 64-bit slots"}.get(
            os.environ.get("BSW_GLOBAL_KERNEL", ""), "?")}
if KswReference.available() and not os.environ.get("GLOBAL_BENCH_NO_CPU"):
    K = KswReference(); P = make_params(**c["P"])
    t0 = time.perf_counter(); done = 0
    while time.perf_counter() - t0 < 3.0:
        for k in range(len(pairs)):
            K.global_align(P, qer[c["qoff"][k]: c["qoff"][k + 1]], ref[c["toff"][k]: c["toff"][k + 1]], int(z["w"][k]))
        done += len(pairs)
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"kind": "reference", "what": "ksw_global2 (tools/bwa/ksw.c, unmodified) called per pair through ctypes",
                            "cores": 1, "value": done / dt, "unit": "alignments/s"}
print(json.dumps(line))
eng.close()
