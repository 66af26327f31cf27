#!/bin/bash
# One GPU session's worth of evidence: parity tests, smoke, the bench line, all-config check,
# ncu launch lists + one --set full capture per headline workload.  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
# Outputs land in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/${TAG}_box.txt 2>&1
nproc >> $OUT/${TAG}_box.txt
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench_short8.json 2> $OUT/${TAG}_bench_short8.err; tail -c 600 $OUT/${TAG}_bench_short8.json
timeout 600 python scripts/gpu_check.py 20000 1000000 0 > $OUT/${TAG}_gpu_check.log 2>&1; cp $OUT/gpu_check_v0.json $OUT/${TAG}_gpu_check_all_configs.json
grep parity $OUT/${TAG}_gpu_check.log
for WL in short8 long16; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload $WL"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_$WL.csv $CMD > $OUT/${TAG}_launches_$WL.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsw_short -s 33 -c 11 -f -o $OUT/${TAG}_prof_short8 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload short8 > $OUT/${TAG}_prof_short8.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bsw_short -s 3 -c 1 -f -o $OUT/${TAG}_prof_long16 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --workload long16 > $OUT/${TAG}_prof_long16.log 2>&1
ls -la $OUT
