#!/bin/bash
# parity of bsw_global on the box + the three bench points (default chunks, one chunk, one band for all)
T=${1:-r04e}
mkdir -p gpurun_out
( timeout 120 python -m pytest tests/test_global.py -m gpu -q 2>&1 | tail -4; echo "pytest rc=${PIPESTATUS[0]}" ) > gpurun_out/${T}_pytest_global.log
GLOBAL_BENCH_NO_CPU=1 timeout 40 python scripts/global_bench.py 300 20 > gpurun_out/${T}_global_bench_k2.json 2>&1
BSW_GLOBAL_CHUNK=262144 GLOBAL_BENCH_NO_CPU=1 timeout 40 python scripts/global_bench.py 300 20 > gpurun_out/${T}_global_bench_k2_chunk262144.json 2>&1
GLOBAL_BENCH_W=20 BSW_GLOBAL_CHUNK=262144 GLOBAL_BENCH_NO_CPU=1 timeout 40 python scripts/global_bench.py 300 10 > gpurun_out/${T}_global_bench_w20.json 2>&1
cat gpurun_out/${T}_pytest_global.log; for f in gpurun_out/${T}_global_bench_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().split('\n')[-1]); print('$f'[-30:], 'fresh %.2f ms reused %.2f ms kernel %.2f ms gcups %.0f'%(d['seconds']*1e3, d['seconds_reused_result_arrays']*1e3, d['kernel_ms'], d['gcups_band_kernel_only']), d['host_ms_reused_result_arrays'])
"; done
