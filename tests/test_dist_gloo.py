"""N>1 host logic on CPU: world_size-2 gloo.  Ranks take disjoint shards of the seeded stream,
'process' them (here with the oracle, as the checker), and the reductions used by bench.py
(max time, summed pairs / cells) agree with a single-process run."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT))
    import genomicsbench_b200 as gb
    from genomicsbench_b200 import dist as D
    from oracle.pyoracle import Oracle, make_params
    ctx = D.init_dist("gloo")
    assert (ctx.rank, ctx.world) == (rank, world)
    per = 700
    first, n = D.weak_shard(per, ctx.rank)
    cfg = gb.gen_named_config("large")
    pairs, ref, qer = gb.gen_pairs(cfg, first, n)
    cells = Oracle().batch(make_params(), pairs, ref, qer, 100, nthreads=1)
    D.barrier(ctx)
    tmax = D.reduce_max(ctx, [float(rank + 1)])[0]
    tot_pairs, tot_cells, checksum = D.reduce_sum(ctx, [float(n), float(cells), float(pairs["score"].sum())])
    sfirst, sn = D.strong_shard(1001, ctx.rank, ctx.world)
    np.save(os.path.join(tmp, f"r{rank}.npy"), np.array([tmax, tot_pairs, tot_cells, checksum, first, n, sfirst, sn]))
    D.shutdown(ctx)


def test_world2_gloo(tmp_path, lib, oracle):
    import torch.multiprocessing as mp
    from oracle.pyoracle import make_params
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0[:4], r1[:4])                       # every rank sees the same reductions
    assert r0[0] == 2.0 and r0[1] == 1400
    assert (r0[4], r0[5], r1[4], r1[5]) == (0, 700, 700, 700)   # disjoint, contiguous weak shards
    assert (r0[6], r0[7], r1[6], r1[7]) == (0, 501, 501, 500)   # strong split covers [0, 1001)
    cfg = lib.gen_named_config("large")
    pairs, ref, qer = lib.gen_pairs(cfg, 0, 1400)
    cells = oracle.batch(make_params(), pairs, ref, qer, 100)
    assert r0[2] == cells and r0[3] == float(pairs["score"].sum())
