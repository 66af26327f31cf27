#!/bin/bash
# the large launch of the second global kernel (one band for all alignments -> 1 800 blocks) under ncu, sections that cost few passes
T=${1:-r04f}
mkdir -p gpurun_out
GLOBAL_BENCH_W=20 BSW_GLOBAL_CHUNK=262144 timeout 60 ncu --section SchedulerStats --section WarpStateStats --section SourceCounters --section LaunchStats --section Occupancy --section SpeedOfLight --section MemoryWorkloadAnalysis --clock-control none --import-source on -k regex:bsw_global2 -s 1 -c 1 -f -o gpurun_out/${T}_g2_w20_big python scripts/global_bench.py 300 0 > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log | cut -c1-160
