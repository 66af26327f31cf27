#!/bin/bash
# Launch list of one bsw_global call (every class launch of the second kernel, serialised by ncu) + the bench with more repetitions
T=${1:-r04b}
mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.per_cycle_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,launch__grid_size,launch__shared_mem_per_block_dynamic,launch__occupancy_limit_shared_mem --clock-control none -k regex:bsw_global2 -c 40 --csv --log-file gpurun_out/${T}_launches_global2.csv env BSW_GLOBAL_CHUNK=262144 python scripts/global_bench.py 300 0 > gpurun_out/${T}_launches.log 2>&1
BSW_GLOBAL_CHUNK=262144 GLOBAL_BENCH_NO_CPU=1 timeout 60 python scripts/global_bench.py 300 25 > gpurun_out/${T}_global_bench_k2_chunk262144_reps25.json 2>&1
GLOBAL_BENCH_NO_CPU=1 timeout 60 python scripts/global_bench.py 300 25 > gpurun_out/${T}_global_bench_k2_reps25.json 2>&1
BSW_GLOBAL_KERNEL=1 GLOBAL_BENCH_NO_CPU=1 timeout 60 python scripts/global_bench.py 300 25 > gpurun_out/${T}_global_bench_k1_reps25.json 2>&1
cat gpurun_out/${T}_global_bench_k*_reps25.json | cut -c1-420
