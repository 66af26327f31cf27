#!/bin/bash
# ncu --set full of the second global kernel at full occupancy: one band for all alignments -> few launch classes, large grids
T=${1:-r04d}
mkdir -p gpurun_out
GLOBAL_BENCH_W=20 BSW_GLOBAL_CHUNK=262144 GLOBAL_BENCH_NO_CPU=1 timeout 40 python scripts/global_bench.py 300 10 > gpurun_out/${T}_global_bench_w20.json 2>&1
GLOBAL_BENCH_W=20 BSW_GLOBAL_CHUNK=262144 timeout 80 ncu --set full --clock-control none --import-source on -k regex:bsw_global2 -c 3 -f -o gpurun_out/${T}_g2_w20 python scripts/global_bench.py 300 0 > gpurun_out/${T}_ncu.log 2>&1
cut -c1-330 gpurun_out/${T}_global_bench_w20.json; tail -2 gpurun_out/${T}_ncu.log | cut -c1-200
