#!/bin/bash
# Quick GPU iteration: parity tests + all-config timing.  gpurun --timeout 900 -- 'bash scripts/gpu_quick.sh <tag>'
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -4 $OUT/${TAG}_pytest_gpu.log
timeout 600 python scripts/gpu_check.py 20000 1000000 0 > $OUT/${TAG}_gpu_check.log 2>&1; cp $OUT/gpu_check_v0.json $OUT/${TAG}_gpu_check_all_configs.json
grep -E "parity|int peak" $OUT/${TAG}_gpu_check.log
python - <<PY
import json
g=json.load(open("$OUT/${TAG}_gpu_check_all_configs.json"))
for k,v in g['configs'].items(): print(k, {x: round(v[x],3) for x in ['ms_kernel','gcups_eff','gcups_nom','roofline_frac','e2e_ms','e2e_pinned_ms']})
PY
