"""Throughput of the pre-processing kernels (zpq_bwt.cu) on the GPU box, next to libzpaq's own BWT path on the host cores.
usage: python tools/bwt_timing.py [max blocks, default 256]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O
import zpaq_b200 as Z
import refpy as R

N = (1 << 20) - 4096
maxb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = Z.Context(0)
KN = {0: "text", 1: "mixed", 2: "random"}
ctx.compress_blocks("x0,3ci1", [O.gen_block(0, 1, 100000)])       # warm-up: allocations, module load
for nblk, kind, method in [(1, 0, "x0,3ci1"), (16, 0, "x0,3ci1"), (maxb, 0, "x0,3ci1"), (maxb, 1, "x0,3ci1"), (maxb, 2, "x0,3ci1"), (maxb, 1, "x0,7ci1")]:
    pays = [O.gen_block(kind, 3000 + k, N) for k in range(nblk)]
    best = None
    for rep in range(2):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        blocks = ctx.compress_blocks(method, pays)
        wall = time.perf_counter() - t0
        t = ctx.timing(); p = ctx.prep_timing()
        if best is None or p["transform_ms"] < best[0]["transform_ms"]:
            best = (p, t, wall, ctx.launch_count() - l0)
    p, t, wall, nl = best
    mb = nblk * N / 1e6
    print("BWT %-8s %4d x %d B '%s': transform %8.2f ms = %8.1f MB/s | SHA-1+stats %6.2f ms | coder %8.1f ms | call %8.1f ms = %7.1f MB/s e2e | %d launches | ratio %.3f"
          % (method, nblk, N, KN[kind], p["transform_ms"], mb / p["transform_ms"] * 1e3, p["analysis_ms"], t["code_ms"], wall * 1e3, mb / wall, nl,
             sum(len(b) for b in blocks) / (nblk * N)), flush=True)
if R.available():
    import bench
    ht = bench.host_threads()
    for kind in (0, 1):
        s, tot = R.bench("x0,3ci1", ht, N, kind, 3000, ht, 0)
        print("CPU reference compressBlock x0,3ci1 '%s': %d blocks on %d threads: %.2f s = %.1f MB/s (BWT + 2-component model)" % (KN[kind], ht, ht, s, ht * N / 1e6 / s), flush=True)
ctx.close()
