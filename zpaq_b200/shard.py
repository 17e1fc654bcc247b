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


# ----------------------------------------------------------------------------------------------------------------------
# The sharded data path: raw blocks leave the rank that holds them as uint8 tensors (NCCL send/recv over
# NVLink when they live in HBM), are coded where they land by the device-resident entry points (no host copy of any
# payload byte: SHA-1 and the level-5 statistics run on the device too), and the coded blocks travel back the same way.
# ----------------------------------------------------------------------------------------------------------------------
def scatter_flat(flat, lens, src=0, device=None):
    """Rank `src` holds `flat` (ONE uint8 tensor on `device`: all blocks back to back) and `lens`; every rank receives its
    contiguous range.  Returns (first_block_index, my_flat_tensor, my_lens).  One message per rank."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_of(device)
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        n[0] = len(lens)
    dist.broadcast(n, src=src)
    nblk = int(n.item())
    lt = torch.zeros(max(nblk, 1), dtype=torch.int64, device=dev)
    if rank == src:
        lt[:nblk] = torch.tensor(list(lens), dtype=torch.int64)
    dist.broadcast(lt, src=src)
    all_lens = lt[:nblk].cpu().tolist()
    parts = partition(nblk, world)
    starts = [0]
    for x in all_lens:
        starts.append(starts[-1] + x)
    lo, hi = parts[rank]
    mine = torch.empty(starts[hi] - starts[lo], dtype=torch.uint8, device=dev)
    if rank == src:
        reqs = []
        for r, (a, b) in enumerate(parts):
            piece = flat[starts[a]:starts[b]]
            if r == rank:
                mine.copy_(piece)
            elif piece.numel():
                reqs.append(dist.isend(piece.contiguous(), dst=r))
        for q in reqs:
            q.wait()
    elif mine.numel():
        dist.recv(mine, src=src)
    return lo, mine, all_lens[lo:hi]


def gather_flat(flat, lens, dst=0, device=None):
    """Every rank contributes ONE uint8 tensor (its blocks back to back) and their lengths; `dst` receives
    (flat tensor of all ranks in rank order, all lengths), the others (None, None)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_of(device)
    cnt = torch.tensor([len(lens), int(flat.numel())], dtype=torch.int64, device=dev)
    allc = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(allc, cnt)
    counts = [int(c[0].item()) for c in allc]
    sizes = [int(c[1].item()) for c in allc]
    maxn = max(1, max(counts))
    mine = torch.zeros(maxn, dtype=torch.int64, device=dev)
    if lens:
        mine[:len(lens)] = torch.tensor(list(lens), dtype=torch.int64)
    alll = [torch.zeros(maxn, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(alll, mine)
    if rank != dst:
        if flat.numel():
            dist.send(flat.contiguous(), dst=dst)
        return None, None
    out = torch.empty(sum(sizes), dtype=torch.uint8, device=dev)
    pos = 0
    for r in range(world):
        if r == rank:
            out[pos:pos + sizes[r]].copy_(flat)
        elif sizes[r]:
            dist.recv(out[pos:pos + sizes[r]], src=r)
        pos += sizes[r]
    all_lens = []
    for r in range(world):
        all_lens += alll[r][:counts[r]].cpu().tolist()
    return out, all_lens


def _ptrs(flat, lens):
    base, out, pos = flat.data_ptr(), [], 0
    for n in lens:
        out.append(base + pos)
        pos += n
    return out


def compress_sharded(ctx, method, flat=None, lens=None, src=0, device=None, sync=None):
    """Rank `src` holds the raw blocks (`flat` uint8 tensor on `device`, `lens`).  Scatter -> zpq_compress_blocks_dev on every
    rank -> gather.  Returns on `src`: (archive tensor, block lengths, timings dict); elsewhere (None, None, timings).
    `sync()` is called around the phases for timing (e.g. cuda synchronize + barrier)."""
    import time
    sync = sync or (lambda: None)
    lib_bound = lambda n: n + n // 8 + 1024 + 70000
    sync(); t0 = time.perf_counter()
    first, mine, mylens = scatter_flat(flat, lens, src=src, device=device)
    sync(); t1 = time.perf_counter()
    caps = [lib_bound(n) for n in mylens]
    outbuf = torch.empty(sum(caps), dtype=torch.uint8, device=mine.device)
    st, ol = ctx.compress_blocks_dev(method, _ptrs(mine, mylens), mylens, _ptrs(outbuf, caps), caps) if mylens else ([], [])
    tm = ctx.timing() if mylens else {"init_ms": 0.0, "code_ms": 0.0}
    if any(st):
        raise RuntimeError("zpq_compress_blocks_dev: status %s" % sorted(set(st)))
    # pack the blocks back to back (they sit at the start of their capacity slots)
    packed = torch.empty(sum(ol), dtype=torch.uint8, device=mine.device)
    pos = cpos = 0
    for n, c in zip(ol, caps):
        packed[pos:pos + n].copy_(outbuf[cpos:cpos + n])
        pos += n; cpos += c
    sync(); t2 = time.perf_counter()
    arc, alens = gather_flat(packed, ol, dst=src, device=device)
    sync(); t3 = time.perf_counter()
    return arc, alens, {"scatter_s": t1 - t0, "code_wall_s": t2 - t1, "gather_s": t3 - t2,
                        "dev_s": (tm["init_ms"] + tm["code_ms"]) / 1e3, "first_block": first, "blocks": len(mylens)}


def decompress_sharded(ctx, arc=None, src=0, device=None, sync=None, verify=True):
    """Rank `src` holds a complete archive (`arc`: uint8 tensor on `device`).  Its framing is scanned on the host of `src`
    (zpq_scan_archive), the coded blocks are scattered, every rank decodes its blocks with zpq_decode_blocks_dev, checks
    the stored SHA-1s on the device and the payloads are gathered.  Returns on `src` (payload tensor, lengths, timings)."""
    import time
    import zpaq_b200 as Z
    sync = sync or (lambda: None)
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = _device_of(device)
    meta = [None]
    flat = lens = None
    sync(); t0 = time.perf_counter()
    if rank == src:
        host = arc.cpu().numpy().tobytes() if arc.device.type != "cpu" else arc.numpy().tobytes()
        segs = Z.scan_archive(host)
        blocks = {}
        for s in segs:
            b = blocks.setdefault(s.block, {"hdr": host[s.hdr_off:s.hdr_off + s.hlen], "seg_len": [], "hint": 0, "sha1": [], "lo": s.coded_off, "hi": 0})
            b["seg_len"].append(int(s.coded_len)); b["hi"] = int(s.coded_off + s.coded_len)
            b["hint"] = -1 if (s.size_hint < 0 or b["hint"] < 0) else b["hint"] + int(s.size_hint)
            b["sha1"].append(bytes(s.sha1) if s.has_sha1 else None)
        order = sorted(blocks)
        # the coded bytes of a block's segments are NOT contiguous in the archive (segment headers and trailers sit between
        # them): gather them with one concatenation on the device (segments come in archive = block order)
        pieces = [(int(s.coded_off), int(s.coded_len)) for s in segs]
        flat = torch.cat([arc[o:o + n] for o, n in pieces]) if pieces else torch.empty(0, dtype=torch.uint8, device=dev)
        lens = [sum(blocks[k]["seg_len"]) for k in order]
        meta = [[(blocks[k]["hdr"], blocks[k]["seg_len"], blocks[k]["hint"], blocks[k]["sha1"]) for k in order]]
    dist.broadcast_object_list(meta, src=src)           # headers and segment tables: a few hundred bytes per block
    first, mine, mylens = scatter_flat(flat, lens, src=src, device=device)
    sync(); t1 = time.perf_counter()
    mymeta = meta[0][first:first + len(mylens)]
    caps = [(m[2] if m[2] >= 0 else 4 * n + (1 << 16)) + 1 for m, n in zip(mymeta, mylens)]
    outbuf = torch.empty(sum(c + 15 for c in caps), dtype=torch.uint8, device=mine.device)
    optr, pos = [], 0
    for c in caps:
        optr.append(outbuf.data_ptr() + pos); pos += c + 15
    res = ctx.decode_blocks_dev([m[0] for m in mymeta], _ptrs(mine, mylens), mylens, optr, caps, [m[1] for m in mymeta]) if mylens else []
    tm = ctx.timing() if mylens else {"init_ms": 0.0, "code_ms": 0.0}
    bad = [r[0] for r in res if r[0] != 0]
    if bad:
        raise RuntimeError("zpq_decode_blocks_dev: status %s" % sorted(set(bad)))
    # payload = model output minus the post-processor flag byte (PASS); per-segment SHA-1 on the device
    plens = [r[1] - 1 for r in res]
    if verify and res:
        sp, sl, want = [], [], []
        for (stt, olen, sout), p0, m in zip(res, optr, mymeta):
            at = p0 + 1
            for k, n in enumerate(sout):
                n2 = n - (1 if k == 0 else 0)
                if m[3][k] is not None:
                    sp.append(at); sl.append(n2); want.append(m[3][k])
                at += n2
        if sp:
            got, _ = ctx.analyze_dev(sp, sl, sha1=True, gaps=False)
            if got != want:
                raise RuntimeError("SHA-1 mismatch after sharded decode")
    packed = torch.empty(sum(plens), dtype=torch.uint8, device=mine.device)
    pos = cpos = 0
    for n, c in zip(plens, caps):
        packed[pos:pos + n].copy_(outbuf[cpos + 1:cpos + 1 + n])
        pos += n; cpos += c + 15
    sync(); t2 = time.perf_counter()
    out, olens = gather_flat(packed, plens, dst=src, device=device)
    sync(); t3 = time.perf_counter()
    return out, olens, {"scatter_s": t1 - t0, "code_wall_s": t2 - t1, "gather_s": t3 - t2,
                        "dev_s": (tm["init_ms"] + tm["code_ms"]) / 1e3, "first_block": first, "blocks": len(mylens)}
