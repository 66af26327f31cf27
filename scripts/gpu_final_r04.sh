#!/bin/bash
# last GPU call of the round: bsw_global bench points of HEAD, then the whole GPU suite
T=${1:-r04g}
mkdir -p gpurun_out
GLOBAL_BENCH_NO_CPU=1 timeout 30 python scripts/global_bench.py 300 20 > gpurun_out/${T}_global_bench_k2.json 2>&1
BSW_GLOBAL_CHUNK=262144 GLOBAL_BENCH_NO_CPU=1 timeout 30 python scripts/global_bench.py 300 20 > gpurun_out/${T}_global_bench_k2_chunk262144.json 2>&1
for f in gpurun_out/${T}_global_bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().split('\n')[-1]); print('$f'[-32:], 'fresh %.2f ms reused %.2f ms kernel %.2f ms'%(d['seconds']*1e3, d['seconds_reused_result_arrays']*1e3, d['host_ms_reused_result_arrays']['kernel_ms']))
"; done
( python -m pytest tests/ -m gpu -x -q 2>&1 | tail -6; echo "pytest rc=${PIPESTATUS[0]}" ) > gpurun_out/${T}_pytest_gpu.log
cat gpurun_out/${T}_pytest_gpu.log
