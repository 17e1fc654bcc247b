#!/usr/bin/env python
"""tools/sharded_roundtrip.py -- the multi-GPU data path around the hot path, end to end:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_roundtrip.py [--blocks 64] [--bytes 1044480] [--method 5] [--check-single]

Rank 0 alone holds the raw blocks (in its HBM).  zpaq_b200.shard.compress_sharded scatters them to the ranks' HBM over
NCCL, every rank codes its contiguous range with zpq_compress_blocks_dev (SHA-1 and level-5 statistics on the device: no
payload byte visits a host), the coded blocks are gathered over NCCL; decompress_sharded sends the archive's blocks the
same way round, verifies the stored SHA-1s on the device and gathers the payloads.  Prints one JSON line with the phase
times (coding: device-side max over ranks).  `--check-single` also compresses everything on rank 0 alone and compares the
archives byte for byte.  (bench.py's default line has every rank code its own data: the coding path has no exchange step;
`bench.py --mode sharded` reports this pipeline.)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gen_blocks_tensor(nblk, nbytes, kind=1, seed=777):
    """Synthetic blocks back to back in ONE pinned host tensor (the generator is part of the library's test data, written
    with the same LCGs as BASELINE.md section 3: reimplemented here with numpy so that no oracle code is on this path)."""
    import numpy as np
    import torch
    out = torch.empty(nblk * nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    a = out.numpy()
    alphabet = np.frombuffer(b"etaoin shrdlu\n", dtype=np.uint8)
    # closed form of the LCG s = s * 1664525 + 1013904223 (mod 2^32): s_k = A^k * s0 + C * (A^k - 1) / (A - 1), by doubling
    for b in range(nblk):
        s0 = np.uint32((seed + b) & 0xFFFFFFFF)
        states = np.empty(nbytes, dtype=np.uint32)
        mul = np.ones(1, dtype=np.uint32); add = np.zeros(1, dtype=np.uint32)
        A, Cc = np.uint32(1664525), np.uint32(1013904223)
        # s_{k+1} = A s_k + C for k = 0..n-1 via vectorised prefix doubling
        m = np.full(nbytes, A, dtype=np.uint32); c = np.full(nbytes, Cc, dtype=np.uint32)
        step = 1
        with np.errstate(over="ignore"):
            while step < nbytes:
                c[step:] = m[step:] * c[:-step] + c[step:]
                m[step:] = m[step:] * m[:-step]
                step *= 2
            states = m * s0 + c
        if kind == 1:
            a[b * nbytes:(b + 1) * nbytes] = alphabet[(states >> np.uint32(24)) % np.uint32(14)]
        else:
            a[b * nbytes:(b + 1) * nbytes] = (states >> np.uint32(24)).astype(np.uint8)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=64)
    ap.add_argument("--bytes", type=int, default=(1 << 20) - 4096)
    ap.add_argument("--method", default="5")
    ap.add_argument("--check-single", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import zpaq_b200 as Z
    from zpaq_b200 import shard

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = Z.Context(local)

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    flat = lens = None
    if rank == 0:
        host = gen_blocks_tensor(args.blocks, args.bytes)
        flat = host.to(dev, non_blocking=False)
        lens = [args.bytes] * args.blocks
    # one small warm-up of both directions (NCCL channels, allocations)
    wf = flat[:args.bytes * min(args.blocks, world)] if rank == 0 else None
    wa, _, _ = shard.compress_sharded(ctx, args.method, wf, lens[:min(args.blocks, world)] if rank == 0 else None, src=0, device=dev, sync=sync)
    shard.decompress_sharded(ctx, wa, src=0, device=dev, sync=sync)

    arc, alens, tc = shard.compress_sharded(ctx, args.method, flat, lens, src=0, device=dev, sync=sync)
    dev_c = shard.max_over_ranks(tc["dev_s"], device=dev)
    out, olens, td = shard.decompress_sharded(ctx, arc, src=0, device=dev, sync=sync)
    dev_d = shard.max_over_ranks(td["dev_s"], device=dev)
    if rank == 0:
        ok = bool(torch.equal(out, flat)) and olens == lens
        same = None
        if args.check_single:
            caps = [n + n // 8 + 80000 for n in lens]
            obuf = torch.empty(sum(caps), dtype=torch.uint8, device=dev)
            st, ol = ctx.compress_blocks_dev(args.method, shard._ptrs(flat, lens), lens, shard._ptrs(obuf, caps), caps)
            pos = 0; pieces = []
            for n, c in zip(ol, caps):
                pieces.append(obuf[pos:pos + n]); pos += c
            same = st == [0] * len(lens) and bool(torch.equal(torch.cat(pieces), arc)) and ol == alens
        mb = args.blocks * args.bytes / 1e6
        print(json.dumps({"tool": "sharded_roundtrip", "n_gpus": world, "blocks": args.blocks, "block_bytes": args.bytes, "method": args.method,
                          "compress": {k: tc[k] for k in ("scatter_s", "code_wall_s", "gather_s")}, "compress_device_s_max": dev_c,
                          "decompress": {k: td[k] for k in ("scatter_s", "code_wall_s", "gather_s")}, "decompress_device_s_max": dev_d,
                          "compress_MBps_e2e": mb / (tc["scatter_s"] + tc["code_wall_s"] + tc["gather_s"]),
                          "decompress_MBps_e2e": mb / (td["scatter_s"] + td["code_wall_s"] + td["gather_s"]),
                          "scatter_GBps": mb / 1e3 / max(tc["scatter_s"], 1e-9), "archive_bytes": int(arc.numel()),
                          "roundtrip_ok": ok, "archive_equals_single_gpu": same}), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
