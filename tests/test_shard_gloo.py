"""CPU, world_size 2 over gloo: block sharding + gather reproduce the single-process archive.
The per-rank "coder" here is the CPU oracle (test infrastructure) standing in for the GPU."""
import os
import socket

import pytest
import torch.multiprocessing as mp

import oraclepy as O
import zpaq_b200 as Z
from util import assemble_block


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _payloads():
    return [O.gen_block(i % 3, 50 + i, 600 + 37 * i) for i in range(7)]


def _code(p):
    hdr = Z.model_from_method("4", p)
    return assemble_block(hdr, [O.encode(hdr, b"\0" + p)], [p], [b""], [str(len(p)).encode()])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from zpaq_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pays = _payloads()
    lo, hi = shard.my_range(len(pays))
    mine = [_code(p) for p in pays[lo:hi]]
    allb = shard.gather_blocks(mine, dst=0)
    t = shard.max_over_ranks(1.0 + rank)
    # the tensor path: raw blocks scattered from rank 0 (only rank 0 holds them), coded, coded bytes gathered back
    first, got = shard.scatter_blocks(pays if rank == 0 else None, src=0)
    assert first == lo and [bytes(t.numpy().tobytes()) for t in got] == pays[lo:hi]
    coded = shard.gather_segments([_code(bytes(t.numpy().tobytes())) for t in got], dst=0)
    if rank == 0:
        assert coded == allb
        q.put((b"".join(allb), t))
    else:
        assert coded is None
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_contiguous_and_balanced():
    from zpaq_b200.shard import partition
    for n in (0, 1, 7, 1029):
        for w in (1, 2, 4, 8):
            parts = partition(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_two_rank_gloo_gather_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    arc, t = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert arc == b"".join(_code(p) for p in _payloads())
    assert t == 2.0
