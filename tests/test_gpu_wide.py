"""GPU: models with more than 32 components (Predictor holds comp[256], libzpaq.h:1131-1137) go through the wide kernel
(one thread per block, zpq_kernels.cu) and must give the reference's bytes like every other tier."""
import pytest

import oraclepy as O
import zpaq_b200 as Z

pytestmark = pytest.mark.gpu

# 34 components: sixteen (ICM, ISSE) pairs with different sparse contexts, a MATCH, a final MIX over all of them;
# 45 components: every component type incl. CM, SSE, MIX2, AVG-free chain of mixers
WIDE_METHODS = [
    "x0,0" + "".join("c0,0,%d,255i1" % (k % 7) for k in range(16)) + "am",
    "x0,0w2c256" + "".join("c0,%d,255i%d" % (2 + k, 1 + k % 2) for k in range(18)) + "ac8,0,255mm16ts19t0",
]


@pytest.mark.parametrize("method", WIDE_METHODS)
def test_wide_models_match_the_oracle_and_the_reference(ctx, method):
    pays = [O.gen_block(1, 21, 2500), O.gen_block(0, 22, 1800), b"", O.gen_block(2, 23, 900)]
    hdr = Z.model_from_method(method, pays[0])
    assert hdr[6] > 32
    datas = [b"\0" + p for p in pays]
    want = [O.encode(hdr, d) for d in datas]
    enc = ctx.encode_blocks([hdr] * len(datas), datas)
    assert ctx.tiers()["wide"] == len(datas)
    assert all(st == 0 and segs[0] == want[i] for i, (st, segs) in enumerate(enc))
    dec = ctx.decode_blocks([hdr] * len(datas), [w + b"\0\0\0\0" for w in want], caps=[len(d) + 8 for d in datas])
    assert all(st == 0 and segs[0] == datas[i] for i, (st, segs) in enumerate(dec))
    import refpy
    if refpy.available():
        blocks = ctx.compress_blocks(method, pays)
        assert blocks == [refpy.compress_block(p, method) for p in pays]
        out, nb = ctx.decompress_archive(b"".join(blocks), sum(len(p) for p in pays))
        assert nb == len(pays) and out == b"".join(pays)
