"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import pytest

import oraclepy as O
import zpaq_b200 as Z

pytestmark = pytest.mark.gpu

CFG_ALL_METHOD = "x0,0ci1,1,1,1,1,1,1,1c256c8,0,255c16,0,255,255aw1c0,0,255,1001,255c0,1010,255m16,24"


def _roundtrip(ctx, hdr, payloads):
    datas = [b"\0" + p for p in payloads]
    want = [O.encode(hdr, d) for d in datas]
    got = ctx.encode_blocks([hdr] * len(datas), datas)
    for i, (st, segs) in enumerate(got):
        assert st == 0, (i, st)
        assert segs[0] == want[i], "coded bytes differ from the oracle in block %d" % i
    dec = ctx.decode_blocks([hdr] * len(datas), [w + b"\0\0\0\0" for w in want], caps=[len(d) + 16 for d in datas])
    for i, (st, segs) in enumerate(dec):
        assert st == 0, (i, st)
        assert segs[0] == datas[i]


@pytest.mark.parametrize("method", ["4", "5", CFG_ALL_METHOD])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_method_parity(kctx, method, kind):
    ctx = kctx
    payloads = [O.gen_block(kind, 100 + i, n) for i, n in enumerate((0, 1, 7, 300, 5000, 40000))]
    hdrs = {}
    for p in payloads:
        hdrs.setdefault(Z.model_from_method(method, p), []).append(p)
    for hdr, ps in hdrs.items():
        _roundtrip(ctx, hdr, ps)
