#!/bin/bash
# compute-sanitizer passes over a small invocation of every kernel (memcheck + racecheck + synccheck).
# Run under gpurun; logs land in gpurun_out/.
set -u
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import genomicsbench_b200 as gb
cfg = gb.gen_named_config("large")
pairs, ref, qer = gb.gen_pairs(cfg, 0, 6000)
ref[::503] = 4                                   # N pairs -> byte kernel
with gb.Engine(long_min_qlen=200) as eng:        # queries >= 200 -> warp-per-pair kernel
    a = pairs.copy(); eng.extend(a, ref, qer, 100)
    pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
    eng.extend(pp, pr, pq, 100)
    assert all(np.array_equal(a[f], pp[f]) for f in gb.RESULT_FIELDS)
    print("stats", eng.stats())
    b = pairs.copy(); eng.extend(b, ref, qer, 16)   # narrow band: the packed kernel's circular rows
    print("w=16", int(b["score"].sum()))
    # banded global alignment: shared-memory rows (small bands) and the global-scratch form (w = 300)
    g = pairs[:1500].copy(); g["len1"] = np.minimum(g["len1"], g["len2"] + 20)
    for w in (25, 300):
        score, cigar, off = eng.global_align(g, ref, qer, w)
    print("global", int(score.sum()), len(cigar))
# a PCIe-bound batch: partitioned streams, speculative copies, small chunks
cfg = gb.gen_named_config("short8")
pairs, ref, qer = gb.gen_pairs(cfg, 0, 140000)
with gb.Engine() as eng:
    a = pairs.copy(); eng.extend(a, ref, qer, 100)
    pp, pr, pq = gb.pinned_copy(pairs), gb.pinned_copy(ref), gb.pinned_copy(qer)
    eng.extend(pp, pr, pq, 100)
    assert all(np.array_equal(a[f], pp[f]) for f in gb.RESULT_FIELDS)
    print("short8 stats", eng.stats())
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitize_$tool.log python /tmp/san_case.py > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
