#!/bin/bash
# Multi-GPU evidence: the torchrun bench line at N GPUs (both arms), the multi-device engine tests,
# PCIe probe.  gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_multi.sh <tag> N'
set -u
TAG=${1:-rXX}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_box_n$N.txt 2>&1
nproc >> $OUT/${TAG}_box_n$N.txt
nvidia-smi topo -m >> $OUT/${TAG}_box_n$N.txt 2>&1
timeout 300 python scripts/pcie_probe.py > $OUT/${TAG}_pcie.txt 2>&1; cat $OUT/${TAG}_pcie.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "multi or device or partition" > $OUT/${TAG}_pytest_multi_n$N.log 2>&1; tail -3 $OUT/${TAG}_pytest_multi_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref_n$N.json 2> $OUT/${TAG}_bench_ref_n$N.err
tail -c 400 $OUT/${TAG}_bench_ref_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_short8_n$N.json 2> $OUT/${TAG}_bench_short8_n$N.err
tail -c 1500 $OUT/${TAG}_bench_short8_n$N.json; tail -5 $OUT/${TAG}_bench_short8_n$N.err
