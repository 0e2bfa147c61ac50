"""Multi-GPU plumbing: one process per GPU, samples-per-pixel split across ranks, one reduce of the float4 accumulation
buffers (NCCL over NVLink on GPUs, gloo in the CPU tests).  The reference is single-process (main.rs:730-778 only
fans out threads per pixel); samples are independent, so the path shards by spp with no collective inside the bounce
loop.  Philox streams are keyed by the GLOBAL sample index, so the sample set does not depend on the
number of ranks."""
from __future__ import annotations

from typing import Tuple


def rank_sample_range(total_spp: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first global sample index, number of samples) rendered by `rank`; remainders go to the lowest ranks."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(total_spp, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_accum(accum, dst: int = 0):
    """Sum the per-rank accumulation tensors onto rank `dst` (torch.distributed.reduce; NCCL -> ncclReduce)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
