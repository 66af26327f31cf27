#!/bin/bash
# ncu evidence for the round: (1) launch list of one bench command, (2) one --set full capture of
# the dominant kernel.  Run under gpurun; outputs land in gpurun_out/ (copy summaries to profiles/).
#   scripts/ncu_profile.sh <workload> <pairs> [skip] [count] [kernel-regex]
set -u
OUT=gpurun_out
mkdir -p $OUT
WL=${1:-short8}
N=${2:-1000000}
SKIP=${3:-30}
CNT=${4:-3}
KRE=${5:-bsw_short_kernel}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $WL --pairs-per-gpu $N"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$WL.csv $CMD > $OUT/launches_$WL.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -f -o $OUT/prof_$WL $CMD > $OUT/prof_$WL.log 2>&1
ls -la $OUT
