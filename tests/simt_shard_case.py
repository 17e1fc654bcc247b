"""Run by tests/test_simt_preflight.py as `python tests/simt_shard_case.py <emulator library>`: the sharded data path (zpaq_b200.shard.compress_sharded /
decompress_sharded) over gloo with two ranks, every rank coding with the SIMT emulator build (device memory = host memory).
Checks the gathered archive against the unmodified reference block by block and the round trip against the input."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def worker(rank, world, port, libpath, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import zpaq_b200 as Z
    Z.LIB_PATH = libpath
    from zpaq_b200 import shard
    import oraclepy as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = Z.Context(0)
    pays = [O.gen_block(i % 3, 70 + i, 900 + 211 * i) for i in range(5)]
    flat = lens = None
    if rank == 0:
        flat = torch.frombuffer(bytearray(b"".join(pays)), dtype=torch.uint8)
        lens = [len(p) for p in pays]
    arc, alens, tm = shard.compress_sharded(ctx, "5", flat, lens, src=0, device="cpu")
    out, olens, tm2 = shard.decompress_sharded(ctx, arc, src=0, device="cpu")
    # the same scatter / gather around a method that pre-processes on the device (BWT: suffix sort of each rank's blocks)
    arc_b, alens_b, _ = shard.compress_sharded(ctx, "x0,3ci1", flat, lens, src=0, device="cpu")
    if rank == 0:
        q.put((arc.numpy().tobytes(), alens, out.numpy().tobytes(), olens, arc_b.numpy().tobytes(), alens_b))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import zpaq_b200 as Z
    Z.LIB_PATH = sys.argv[1]           # the SIMT emulator build (test infrastructure; the package has no switch for it)
    import oraclepy as O
    c = mp.get_context("spawn")
    q = c.Queue()
    port = _free_port()
    procs = [c.Process(target=worker, args=(r, 2, port, Z.LIB_PATH, q)) for r in range(2)]
    for p in procs:
        p.start()
    arc, alens, out, olens, arc_b, alens_b = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pays = [O.gen_block(i % 3, 70 + i, 900 + 211 * i) for i in range(5)]
    assert out == b"".join(pays) and olens == [len(p) for p in pays]
    import refpy
    if refpy.available():
        want = [refpy.compress_block(p, "5") for p in pays]
        assert arc == b"".join(want) and alens == [len(w) for w in want]
        want_b = [refpy.compress_block(p, "x0,3ci1") for p in pays]
        assert arc_b == b"".join(want_b) and alens_b == [len(w) for w in want_b]
    print("simt shard ok: %d blocks over 2 ranks, %d archive bytes" % (len(pays), len(arc)))
