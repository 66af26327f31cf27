#!/bin/bash
# One GPU call for bsw_global's second kernel: parity (all three forms), A/B throughput, memcheck on the small goldens,
# one ncu capture of the new kernel.  Usage (from the repo root, under gpurun): bash scripts/gpu_global_ab.sh TAG
T=${1:-r04a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${T}_box.txt 2>&1
( timeout 170 python -m pytest tests/test_global.py -m gpu -q 2>&1 | tail -15; echo "pytest rc=${PIPESTATUS[0]}" ) > gpurun_out/${T}_pytest_global.log
for k in 2 1 2w; do
    kk=$k; [ "$k" = 2 ] && kk=""
    BSW_GLOBAL_KERNEL=$kk GLOBAL_BENCH_NO_CPU=1 timeout 60 python scripts/global_bench.py 300 5 > gpurun_out/${T}_global_bench_k$k.json 2> gpurun_out/${T}_global_bench_k$k.err
done
for ch in 65536 262144; do
    BSW_GLOBAL_CHUNK=$ch GLOBAL_BENCH_NO_CPU=1 timeout 60 python scripts/global_bench.py 300 5 > gpurun_out/${T}_global_bench_k2_chunk$ch.json 2> gpurun_out/${T}_global_bench_k2_chunk$ch.err
done
timeout 80 ncu --set full --clock-control none --import-source on -k regex:bsw_global2 -c 3 -f -o gpurun_out/${T}_g2 python scripts/global_bench.py 300 0 > gpurun_out/${T}_ncu.log 2>&1
( timeout 90 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_global.py -m gpu -q -k "golden and (tiny or gaps)" 2>&1 | tail -12; echo "memcheck rc=${PIPESTATUS[0]}" ) > gpurun_out/${T}_sanitize_global2.txt
tail -3 gpurun_out/${T}_pytest_global.log; cat gpurun_out/${T}_global_bench_k*.json | cut -c1-600; tail -3 gpurun_out/${T}_sanitize_global2.txt
