#!/bin/bash
# ncu evidence of HEAD: launch list of the bench command + one --set full capture of the DP launches of one step, per
# workload, SUMMARISED ON THE BOX (scripts/ncu_summary.py; the .ncu-rep files are deleted -- gpurun_out/ is capped at 64 MiB).
#   gpurun --timeout 1200 -- 'bash scripts/gpu_ncu.sh <tag> "<workload>:<max launches> ..."'
set -u
TAG=${1:-rXX}
SPEC=${2:-"short8:11 long16:1 large:5 sweep_w100_z100:5 sweep_w500_z100:3"}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-split-legs"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_short8.csv $CMD > /dev/null 2>&1
for S in $SPEC; do
  WL=${S%%:*}; CNT=${S##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:bsw_short16 -c $CNT -f -o /tmp/prof_$WL \
      python scripts/resident_run.py $WL 0 1 > $OUT/${TAG}_prof_$WL.log 2>&1
  python scripts/ncu_summary.py /tmp/prof_$WL.ncu-rep 0 > $OUT/${TAG}_ncu_short16_kernel_$WL.txt 2>&1
  rm -f /tmp/prof_$WL.ncu-rep
  head -c 600 $OUT/${TAG}_ncu_short16_kernel_$WL.txt | head -5
done
du -sh $OUT
