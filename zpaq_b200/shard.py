"""Block sharding across ranks.

Archive blocks are independent (own header, own model, own coder), so multi-GPU
work is a static partition of the block list: rank r codes a contiguous range,
no collective on the data path.  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only around the path: to scatter raw input blocks from the
rank that holds them, to gather the coded segments back, and to take the max of
the per-rank device times.

scatter_blocks / gather_segments move BYTES as uint8 tensors with point-to-point
sends (NCCL over NVLink when the tensors live in HBM: one message per rank, the
rank's whole contiguous range, so a 1 GiB shard is one transfer, not a thousand);
the block lengths travel first as one small int64 broadcast / all_gather.
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


def _device_of(device):
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def scatter_blocks(blocks, src=0, device=None):
    """Rank `src` holds `blocks` (list of bytes-like); every rank gets its contiguous range.

    Returns (first_index, [uint8 tensor per block]) on every rank.  The tensors are views of ONE receive buffer on
    `device` (default: the current CUDA device under NCCL, CPU under gloo)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_of(device)
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        n[0] = len(blocks)
    dist.broadcast(n, src=src)
    nblk = int(n.item())
    lens = torch.zeros(max(nblk, 1), dtype=torch.int64, device=dev)
    if rank == src:
        lens[:nblk] = torch.tensor([len(b) for b in blocks], dtype=torch.int64)
    dist.broadcast(lens, src=src)
    lens = lens[:nblk].cpu().tolist()
    parts = partition(nblk, world)
    lo, hi = parts[rank]
    total = sum(lens[lo:hi])
    buf = torch.empty(total, dtype=torch.uint8, device=dev)
    if rank == src:
        reqs, keep = [], []
        for r, (a, b) in enumerate(parts):
            flat = torch.frombuffer(bytearray(b"".join(bytes(x) for x in blocks[a:b])), dtype=torch.uint8) if b > a and sum(lens[a:b]) else torch.empty(0, dtype=torch.uint8)
            if r == rank:
                buf.copy_(flat)
            elif flat.numel():
                t = flat.to(dev, non_blocking=False)
                keep.append(t)
                reqs.append(dist.isend(t, dst=r))
        for q in reqs:
            q.wait()
    elif total:
        dist.recv(buf, src=src)
    out, pos = [], 0
    for ln in lens[lo:hi]:
        out.append(buf[pos:pos + ln])
        pos += ln
    return lo, out


def gather_segments(local, dst=0, device=None):
    """Variable-length gather of this rank's coded segments (uint8 tensors or bytes-like, in block order) to `dst`.

    Returns the list of bytes objects of ALL ranks in block order on `dst`, None elsewhere."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_of(device)
    tens = [t if isinstance(t, torch.Tensor) else torch.frombuffer(bytearray(bytes(t)), dtype=torch.uint8) for t in local]
    counts = torch.tensor([len(tens)], dtype=torch.int64, device=dev)
    allc = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, counts)
    maxn = max(1, max(int(c.item()) for c in allc))
    mylens = torch.zeros(maxn, dtype=torch.int64, device=dev)
    if tens:
        mylens[:len(tens)] = torch.tensor([t.numel() for t in tens], dtype=torch.int64)
    alll = [torch.zeros(maxn, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(alll, mylens)
    flat = torch.cat([t.to(dev).reshape(-1) for t in tens]) if tens else torch.empty(0, dtype=torch.uint8, device=dev)
    if rank != dst:
        if flat.numel():
            dist.send(flat, dst=dst)
        return None
    out = []
    for r in range(world):
        lens = alll[r][:int(allc[r].item())].cpu().tolist()
        if r == rank:
            buf = flat
        else:
            buf = torch.empty(sum(lens), dtype=torch.uint8, device=dev)
            if buf.numel():
                dist.recv(buf, src=r)
        host = buf.cpu().numpy().tobytes()
        pos = 0
        for ln in lens:
            out.append(host[pos:pos + ln])
            pos += ln
    return out
