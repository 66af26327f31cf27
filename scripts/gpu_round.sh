#!/bin/bash
# One GPU session's worth of evidence: parity tests, smoke, the bench line (all configs + split legs), chain / global
# benches, ncu launch list + --set full summaries of the DP launches (summarised on the box: gpurun_out/ is capped at
# 64 MiB).  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'; copy what should be judged into profiles/.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > $OUT/${TAG}_box.txt 2>&1
nproc >> $OUT/${TAG}_box.txt
timeout 800 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; tail -c 300 $OUT/${TAG}_bench_ref.json
timeout 800 python bench.py > $OUT/${TAG}_bench_short8.json 2> $OUT/${TAG}_bench_short8.err; echo "bench rc=$?"; tail -c 400 $OUT/${TAG}_bench_short8.json
timeout 300 python scripts/chain_bench.py 400 3 2>/dev/null | tail -1 > $OUT/${TAG}_chain_bench.json; cut -c1-200 $OUT/${TAG}_chain_bench.json
timeout 300 python scripts/global_bench.py 2>/dev/null | tail -1 > $OUT/${TAG}_global_bench.json; cut -c1-200 $OUT/${TAG}_global_bench.json
bash scripts/gpu_ncu.sh $TAG "${NCU_SPEC:-short8:11 long16:1 large:5 sweep_w100_z100:5 sweep_w500_z100:3}" > $OUT/${TAG}_ncu.log 2>&1
du -sh $OUT
