#!/usr/bin/env python
"""Concurrent pinned-copy ceiling of the box: every rank (one per GPU, torchrun) moves page-locked buffers host -> device
and device -> host at the same time; rank 0 prints the aggregate GB/s per direction for 1..N ranks active.  This is the
denominator of the end-to-end scaling figures: N ranks of bench.py push their records / packed words through the same
host memory system and root complexes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/fabric_probe.py
"""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
from genomicsbench_b200 import dist as D  # noqa: E402


def main():
    ctx = D.init_dist()
    torch.cuda.set_device(ctx.local_rank)
    dev = torch.device("cuda", ctx.local_rank)
    bound = D.bind_near_gpu(ctx.local_rank) if ctx.world > 1 else False
    MB = 1 << 20
    h_in = torch.empty(256 * MB, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(256 * MB, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(256 * MB, dtype=torch.uint8, device=dev)
    d_out = torch.empty(256 * MB, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}

    def run(active, h2d, d2h, reps=8):
        """ranks < active copy; returns (seconds, bytes this rank moved per direction)"""
        mine = ctx.rank < active
        D.barrier(ctx)
        t0 = time.perf_counter()
        if mine:
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s_in):
                        d_in.copy_(h_in, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s_out):
                        h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = D.reduce_max(ctx, [dt if mine else 0.0])[0]
        return dt, reps * 256 * MB

    for mode, (a, b) in {"h2d": (True, False), "d2h": (False, True), "both": (True, True)}.items():
        for active in sorted({1, 2, 4, ctx.world} & set(range(1, ctx.world + 1))):
            run(active, a, b, reps=2)
            dt, nbytes = run(active, a, b)
            res[f"{mode}_ranks{active}"] = {"GBs_per_direction": active * nbytes / dt / 1e9,
                                           "GBs_total": active * nbytes * (2 if (a and b) else 1) / dt / 1e9}
    if ctx.is_main:
        out = {"world": ctx.world, "numa_bound": bound, "buffer_MB": 256, "results": res}
        print(json.dumps(out, indent=1))
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"fabric_probe_n{ctx.world}.json").write_text(json.dumps(out, indent=1))
    D.shutdown(ctx)


if __name__ == "__main__":
    main()
