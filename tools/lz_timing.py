"""Throughput of the store / LZ77 paths on the GPU box (one GPU thread per block; zpq_lz77.cu). usage: lz_timing.py [blocks]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O
import zpaq_b200 as Z
N = (1 << 20) - 4096
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = Z.Context(0)
pays = [O.gen_block(1, 3000 + k, N) for k in range(nblk)]
ctx.compress_blocks("1", pays[:2])
for method in ("10,5,0", "1", "2", "3,100,0"):
    t0 = time.perf_counter()
    blocks = ctx.compress_blocks(method, pays)
    wall = time.perf_counter() - t0
    p = ctx.prep_timing(); t = ctx.timing()
    mb = nblk * N / 1e6
    print("LZ %-8s %4d x %d B 'mixed': pre-processing %9.1f ms = %8.1f MB/s | coder %8.1f ms | call %8.1f ms = %7.1f MB/s | ratio %.3f"
          % (method, nblk, N, p["transform_ms"], mb / max(p["transform_ms"], 1e-3) * 1e3, t["code_ms"], wall * 1e3, mb / wall, sum(len(b) for b in blocks) / (nblk * N)), flush=True)
ctx.close()
