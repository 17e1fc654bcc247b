"""Block sharding across ranks.

Archive blocks are independent (own header, own model, own coder), so multi-GPU
work is a static partition of the block list: rank r codes a contiguous range,
no collective on the data path.  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only to bring per-block results back to rank 0 and to take
the max of the per-rank device times.
"""
import torch
import torch.distributed as dist


def partition(nblk, world):
    """Contiguous, balanced ranges: [(start, stop)] per rank."""
    base, extra = divmod(nblk, world)
    out, pos = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((pos, pos + n))
        pos += n
    return out


def my_range(nblk, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return partition(nblk, world)[rank]


def gather_blocks(local_blocks, dst=0):
    """Variable-length gather of coded blocks (bytes objects) to `dst`, in block order."""
    world = dist.get_world_size()
    rank = dist.get_rank()
    bufs = [None] * world if rank == dst else None
    dist.gather_object(list(local_blocks), bufs, dst=dst)
    if rank != dst:
        return None
    return [b for part in bufs for b in part]


def max_over_ranks(seconds, device=None):
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
