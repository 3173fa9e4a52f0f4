"""Multi-GPU plumbing: one process per GPU (torch.distributed / NCCL over NVLink).  The hot paths shard by
independent units (seeds, reads): the only exchange is ONE broadcast of the read-only index at start-up
(the reference instead H2D-copies the whole index once per device, nvbio/io/fmindex/fmindex_impl.cu:749-835);
there is no collective in the steady state."""
from typing import Optional, Tuple
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous equal shards; the first n_items % world shards get one extra item"""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_tensors(tensors, src: int = 0):
    """in-place broadcast of a list of tensors already allocated with the same shapes on every rank"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in tensors:
        dist.broadcast(t, src=src)


def broadcast_index(index_or_none, genome_or_none, device, src: int = 0):
    """rank `src` holds (FMIndexDevice, genome words); every other rank passes (None, None) and receives
    replicas.  Metadata travels as a small int64 tensor, the arrays as three broadcasts."""
    from .fmindex import FMIndexDevice
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return index_or_none, genome_or_none
    rank = dist.get_rank()
    meta = torch.zeros(11, dtype=torch.int64, device=device)
    if rank == src:
        f = index_or_none
        meta[:11] = torch.tensor([f.length, f.primary] + list(f.L2) + [f.bwt_occ.numel(), f.ssa.numel(), genome_or_none.numel(),
                                  getattr(f, "sa_interval", 16)], dtype=torch.int64, device=device)
    dist.broadcast(meta, src=src)
    m = [int(v) for v in meta.cpu()]
    if rank == src:
        bwt_occ, ssa, genome = index_or_none.bwt_occ, index_or_none.ssa, genome_or_none
    else:
        bwt_occ = torch.empty(m[7], dtype=torch.int32, device=device)
        ssa = torch.empty(m[8], dtype=torch.int32, device=device)
        genome = torch.empty(m[9], dtype=torch.int32, device=device)
    for t in (bwt_occ, ssa, genome):
        dist.broadcast(t, src=src)
    if rank == src:
        return index_or_none, genome_or_none
    return FMIndexDevice(bwt_occ, ssa, m[2:7], m[0], m[1], sa_interval=m[10]), genome


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
