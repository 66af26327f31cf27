#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over the warp-per-pair register kernel (bsw_warp16.cuh): the chunk
# route, a narrow band, and the latency route.  Run under gpurun; logs land in gpurun_out/.
set -u
TOOLS=${1:-"memcheck racecheck synccheck"}
mkdir -p gpurun_out
cat > /tmp/san_warp.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
import genomicsbench_b200 as gb
cfg = gb.gen_named_config("large")
pairs, ref, qer = gb.gen_pairs(cfg, 0, 700)
with gb.Engine(warp_max_pairs=-1) as eng:
    want = pairs.copy(); eng.extend(want, ref, qer, 100)
    want16 = pairs.copy(); eng.extend(want16, ref, qer, 16)
with gb.Engine() as eng:
    a = pairs.copy(); eng.extend(a, ref, qer, 100)
    assert all(np.array_equal(a[f], want[f]) for f in gb.RESULT_FIELDS)
    b = pairs.copy(); eng.extend(b, ref, qer, 16)
    assert all(np.array_equal(b[f], want16[f]) for f in gb.RESULT_FIELDS)
short = pairs[pairs["len2"] <= 255].copy()
with gb.Engine(tiny_batch=1536) as eng:
    a = short.copy(); eng.extend(a, ref, qer, 100)
    assert eng.stats()["n_short"] == len(short)
    sel = pairs["len2"] <= 255
    assert all(np.array_equal(a[f], want[f][sel]) for f in gb.RESULT_FIELDS)
print("warp kernel ok")
PY
for tool in $TOOLS; do
  timeout 300 compute-sanitizer --tool $tool --log-file gpurun_out/sanitize_warp_$tool.log python /tmp/san_warp.py > gpurun_out/sanitize_warp_$tool.out 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_warp_$tool.log | tail -1) $(tail -1 gpurun_out/sanitize_warp_$tool.out)"
done
