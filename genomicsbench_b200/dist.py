"""Multi-GPU plumbing: one process per GPU under torchrun, no data-path collective.

The bsw path is embarrassingly parallel (pairs are independent; the reference's only
parallelism is an OpenMP loop over batches, benchmarks/bsw/main_banded.cpp:279-291), so ranks
only need (a) a rule that hands every rank its shard of the pair stream, (b) a barrier and
(c) max / sum reductions of a few scalars for reporting.  torch.distributed supplies (b) and
(c): NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Sequence


@dataclass
class DistCtx:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    backend: Optional[str] = None     # None => single process, torch.distributed untouched

    @property
    def is_main(self) -> bool:
        return self.rank == 0


def init_dist(backend: Optional[str] = None) -> DistCtx:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* from the environment (torchrun)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return DistCtx()
    import torch
    import torch.distributed as dist
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        # NCCL prints to the process's stdout while the communicator is created (at least its one-line
        # "NCCL version ..." banner; NCCL_DEBUG_FILE does not catch it), where rank 0's single JSON line
        # belongs: create the communicator -- init + a first collective -- with fd 1 pointing at stderr.
        import sys
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
            dist.barrier(device_ids=[local_rank])
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    return DistCtx(rank, world, local_rank, backend)


def shutdown(ctx: DistCtx) -> None:
    if ctx.backend is not None:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def _tensor(ctx: DistCtx, values: Sequence[float]):
    import torch
    dev = torch.device("cuda", ctx.local_rank) if ctx.backend == "nccl" else torch.device("cpu")
    return torch.tensor(list(values), dtype=torch.float64, device=dev)


def barrier(ctx: DistCtx) -> None:
    if ctx.backend is None:
        return
    import torch
    import torch.distributed as dist
    if ctx.backend == "nccl":
        dist.barrier(device_ids=[ctx.local_rank])
        torch.cuda.synchronize()
    else:
        dist.barrier()


def reduce_max(ctx: DistCtx, values: Sequence[float]) -> list:
    if ctx.backend is None:
        return list(values)
    import torch.distributed as dist
    t = _tensor(ctx, values)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().tolist()


def reduce_sum(ctx: DistCtx, values: Sequence[float]) -> list:
    if ctx.backend is None:
        return list(values)
    import torch.distributed as dist
    t = _tensor(ctx, values)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().tolist()


def weak_shard(pairs_per_rank: int, rank: int) -> tuple:
    """Weak scaling: rank r owns pairs [r*P, (r+1)*P) of the config's seeded stream."""
    return rank * pairs_per_rank, pairs_per_rank


def strong_shard(n_total: int, rank: int, world: int) -> tuple:
    """Strong scaling: contiguous near-equal split of [0, n_total)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def bind_near_gpu(physical_index: int) -> bool:
    """Pins the calling process to the CPUs NVML reports as local to the GPU (same NUMA node /
    PCIe root), so that page-locked buffers allocated afterwards are DMA'd without crossing the
    socket interconnect.  Best effort: returns False when NVML or the call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False
