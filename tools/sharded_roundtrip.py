#!/usr/bin/env python
"""tools/sharded_roundtrip.py -- the multi-GPU plumbing around the hot path, end to end:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/sharded_roundtrip.py [--blocks 64] [--bytes 1044480] [--method 5]

Rank 0 alone holds the raw blocks.  They are scattered to the ranks' HBM over NCCL (zpaq_b200.shard.scatter_blocks),
every rank compresses its contiguous range with the CUDA library, the complete archive blocks are gathered back over NCCL
(gather_segments) and rank 0 decompresses the concatenated archive again and compares.  Prints one JSON line with the
times of the three phases (device-side max over ranks for the coding phase).  bench.py does NOT use this path: its
weak-scaling line has every rank code its own data, because the coding path itself has no exchange step.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=64)
    ap.add_argument("--bytes", type=int, default=(1 << 20) - 4096)
    ap.add_argument("--method", default="5")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import oraclepy as O
    import zpaq_b200 as Z
    from zpaq_b200 import shard

    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    blocks = [O.gen_block(1, 777 + i, args.bytes) for i in range(args.blocks)] if rank == 0 else None

    def sync():
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    sync(); t0 = time.perf_counter()
    first, mine = shard.scatter_blocks(blocks, src=0)
    sync(); t1 = time.perf_counter()
    ctx = Z.Context(local)
    host = [bytes(t.cpu().numpy().tobytes()) for t in mine]
    coded = ctx.compress_blocks(args.method, host) if host else []
    tm = ctx.timing()
    dev_s = shard.max_over_ranks((tm["init_ms"] + tm["code_ms"]) / 1e3 if host else 0.0, device=torch.device("cuda", local))
    sync(); t2 = time.perf_counter()
    allb = shard.gather_segments(coded, dst=0)
    sync(); t3 = time.perf_counter()
    if rank == 0:
        arc = b"".join(allb)
        plain, nb = ctx.decompress_archive(arc, args.blocks * args.bytes)
        ok = nb == args.blocks and plain == b"".join(blocks)
        mb = args.blocks * args.bytes / 1e6
        print(json.dumps({"tool": "sharded_roundtrip", "n_gpus": world, "blocks": args.blocks, "block_bytes": args.bytes, "method": args.method,
                          "scatter_s": t1 - t0, "compress_wall_s": t2 - t1, "compress_device_s_max": dev_s, "gather_s": t3 - t2,
                          "scatter_GBps": mb / 1e3 / max(t1 - t0, 1e-9), "archive_bytes": len(arc), "roundtrip_ok": bool(ok)}), flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
