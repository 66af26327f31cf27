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
    # packed route: 2-bit words read in place, RAW pairs, OutScore + 16-byte results; async queue; two device contexts
    cfg.n_rate = 0.001
    pairs, ref, qer = gb.gen_pairs(cfg, 0, 90000)
    want = pairs.copy(); eng.extend(want, ref, qer, 100)
    batch = gb.PackedBatch.from_pairs(pairs, ref, qer, pinned=True)
    out = eng.extend_packed(batch, 100); out16 = eng.extend_packed(batch, 100, compact=True)
    assert all(np.array_equal(out[f], want[f]) and np.array_equal(out16[f], want[f]) for f in gb.RESULT_FIELDS)
    got = pairs.copy()
    tk = [eng.extend_async(got[k * 512:(k + 1) * 512], ref, qer, 100) for k in range(20)]
    [eng.wait(t) for t in tk]
    assert all(np.array_equal(got[f][:10240], want[f][:10240]) for f in gb.RESULT_FIELDS)
with gb.Engine(devices=[0, 0]) as eng:
    out = eng.extend_packed(batch, 100)
    assert eng.stats()["shards"] == 2 and all(np.array_equal(out[f], want[f]) for f in gb.RESULT_FIELDS)
    print("packed / async / sharded ok")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/sanitize_$tool.log python /tmp/san_case.py > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
