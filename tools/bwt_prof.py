"""One BWT compress + decompress of N 'text' blocks (for ncu captures of the zpq_bwt.cu kernels). usage: bwt_prof.py [N]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oraclepy as O
import zpaq_b200 as Z
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = (1 << 20) - 4096
ctx = Z.Context(0)
pays = [O.gen_block(0, 3000 + k, N // 8) for k in range(n)]      # 1/8 length: the coder's latency is not what is profiled here
blocks = ctx.compress_blocks("x0,3ci1", pays)
out, nb = ctx.decompress_archive(b"".join(blocks), sum(len(p) for p in pays))
assert out == b"".join(pays)
print("ok", ctx.prep_timing(), ctx.launch_count())
ctx.close()
