"""The UNCHANGED zpaq.cpp linked against libzpq_b200.so through zpaq_b200/csrc/libzpaq_shim.cpp
(oracle/Makefile target `cli`; the binaries are built where /root/reference exists and travel to the GPU box)."""
import os
import subprocess

import pytest

import oraclepy as O
from util import ROOT

# ZPQ_CLI_B200 / ZPQ_CLI_STREAM: set by tools/simt (pre-flight without a GPU) to the same programs linked against the emulator build
B200 = os.environ.get("ZPQ_CLI_B200", os.path.join(ROOT, "oracle", "_ref", "zpaq_b200"))
REF = os.path.join(ROOT, "oracle", "_ref", "zpaq_ref")
STREAM = os.environ.get("ZPQ_CLI_STREAM", os.path.join(ROOT, "oracle", "_ref", "shim_stream"))
need_cli = pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(REF)), reason="CLI binaries not built (make -C oracle cli)")


def _files(tmp_path):
    names = []
    for i, kind in enumerate((0, 1, 3)):
        p = tmp_path / ("f%d.bin" % i)
        p.write_bytes(O.gen_block(kind, 900 + i, 150000 + 7000 * i))
        names.append(p.name)
    return names


def _run(exe, args, cwd, log=False):
    env = dict(os.environ)
    if log:
        env["ZPQ_SHIM_LOG"] = "1"       # one stderr line per routed call: proves which path ran
    r = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=int(os.environ.get("ZPQ_TEST_TIMEOUT", "900")), env=env)
    return r.returncode, r.stdout + r.stderr


@need_cli
def test_cli_links_and_lz77_methods_stay_on_the_cpu_code(tmp_path):
    names = _files(tmp_path)
    rc, out = _run(B200, ["add", "a1.zpaq"] + names + ["-method", "1"], tmp_path)
    assert rc == 0, out
    rc, out = _run(REF, ["x", "a1.zpaq", "-to", "o1"], tmp_path)
    assert rc == 0, out
    for n in names:
        assert (tmp_path / "o1" / n).read_bytes() == (tmp_path / n).read_bytes()


@need_cli
def test_zpaq_extract_of_stored_blocks_needs_no_gpu(tmp_path):
    """`zpaq x` through the shim's Decompresser::decompress / readSegmentEnd on an archive without context-model blocks
    (method 1: LZ77, unmodelled): every block is routed to the reference's own decoder, nothing touches the GPU library."""
    names = _files(tmp_path)
    rc, out = _run(REF, ["add", "l.zpaq"] + names + ["-method", "1"], tmp_path)
    assert rc == 0, out
    rc, out = _run(B200, ["x", "l.zpaq", "-to", "ol", "-threads", "2"], tmp_path, log=True)
    assert rc == 0, out
    assert "-> reference CPU code" in out and "-> GPU" not in out, out
    for n in names:
        assert (tmp_path / "ol" / n).read_bytes() == (tmp_path / n).read_bytes()


@need_cli
@pytest.mark.gpu
@pytest.mark.parametrize("method", ["4", "5", "3"])
def test_zpaq_add_on_gpu_extracts_with_stock_zpaq(tmp_path, method):
    """`zpaq add -method 4/5` through the GPU library; the archive is read back by the stock reference CLI.  With `-method 3`
    zpaq asks for BWT on this (redundant) input: suffix sorting, the PCOMP program and the two-component model all run in the
    GPU library (zpq_bwt.cu)."""
    names = _files(tmp_path)
    rc, out = _run(B200, ["add", "g.zpaq"] + names + ["-method", method, "-threads", "8"], tmp_path, log=True)
    assert rc == 0, out
    assert "-> GPU, status 0" in out, out
    rc, out = _run(REF, ["x", "g.zpaq", "-to", "og"], tmp_path)
    assert rc == 0, out
    for n in names:
        assert (tmp_path / "og" / n).read_bytes() == (tmp_path / n).read_bytes()
    # same input through the stock CLI gives the same compressed size (bytes differ only in the dates zpaq stores)
    rc, out = _run(REF, ["add", "r.zpaq"] + names + ["-method", method, "-threads", "8"], tmp_path)
    assert rc == 0, out
    assert abs((tmp_path / "g.zpaq").stat().st_size - (tmp_path / "r.zpaq").stat().st_size) <= 64


@need_cli
@pytest.mark.gpu
@pytest.mark.parametrize("method", ["5", "3"])
def test_zpaq_extract_on_gpu(tmp_path, method):
    """`zpaq x` through the shim's Decompresser::decompress: archives written by the STOCK CLI are decoded by the GPU
    library (method 5: context mixing; method 3: BWT, whose two-component model is decoded on the GPU and whose PCOMP
    inverse transform stays the reference's PostProcessor)."""
    names = _files(tmp_path)
    rc, out = _run(REF, ["add", "s.zpaq"] + names + ["-method", method, "-threads", "4"], tmp_path)
    assert rc == 0, out
    rc, out = _run(B200, ["x", "s.zpaq", "-to", "ox", "-threads", "4"], tmp_path, log=True)
    assert rc == 0, out
    assert "components -> GPU" in out and "status 0" in out, out
    for n in names:
        assert (tmp_path / "ox" / n).read_bytes() == (tmp_path / n).read_bytes()


@pytest.mark.skipif(not os.path.exists(STREAM), reason="stream driver not built (make -C oracle cli)")
def test_stream_api_lz77_stays_on_the_cpu_code(tmp_path):
    """libzpaq::compress / decompress through the shim with a method outside the context-model family: no GPU needed."""
    data = O.gen_block(0, 31, 300000)
    (tmp_path / "in").write_bytes(data)
    rc, out = _run(STREAM, ["c", "1", "in", "a.zpaq"], tmp_path, log=True)
    assert rc == 0, out
    assert "0 of them -> GPU" in out, out
    rc, out = _run(STREAM, ["d", "-", "a.zpaq", "out"], tmp_path, log=True)
    assert rc == 0, out
    assert (tmp_path / "out").read_bytes() == data


@pytest.mark.skipif(not os.path.exists(STREAM), reason="stream driver not built (make -C oracle cli)")
@pytest.mark.gpu
def test_stream_api_through_the_gpu_library(tmp_path):
    """libzpaq::compress(in, out, "50") reads its 1 MiB-class blocks ahead and codes them as one GPU batch; the archive
    equals the reference's block for block, and libzpaq::decompress decodes it in one batch."""
    import refpy
    bs = (1 << 20) - 4096
    data = O.gen_block(1, 5, 2 * bs + 123456)
    (tmp_path / "in").write_bytes(data)
    rc, out = _run(STREAM, ["c", "50", "in", "a.zpaq"], tmp_path, log=True)
    assert rc == 0, out
    assert "3 blocks" in out and "3 of them -> GPU" in out, out
    got = (tmp_path / "a.zpaq").read_bytes()
    if refpy.available():
        want = b"".join(refpy.compress_block(data[i:i + bs], "50", b"name" if i == 0 else None, b"comment" if i == 0 else None)
                        for i in range(0, len(data), bs))
        assert got == want
    rc, out = _run(STREAM, ["d", "-", "a.zpaq", "out"], tmp_path, log=True)
    assert rc == 0, out
    assert "-> GPU batch, status 0" in out, out
    assert (tmp_path / "out").read_bytes() == data


@pytest.mark.skipif(not os.path.exists(STREAM), reason="stream driver not built (make -C oracle cli)")
@pytest.mark.gpu
def test_decompresser_handles_long_zero_runs_before_the_trailer(tmp_path):
    """Decompresser::decompress / readSegmentEnd through the shim on every golden archive of the context-model family,
    in pieces of 1000 bytes.  Eleven of them end their coded data with five to seven zero bytes (the last bytes of the
    range flush are zero): the terminator is the LAST four zeros before the 253 / 254 marker (Decoder::skip,
    libzpaq.cpp:2165), and the stored SHA-1 must still be found behind it."""
    import base64
    from util import load, data_of, archive_bytes
    long_runs = 0
    for e in load("archives.json"):
        arc = archive_bytes(e)
        if arc is None or len(arc) > 200000:
            continue
        blocks = O.parse_blocks(arc)
        if not all(0 < b.header[6] for b in blocks):
            continue
        tail = arc[:-1]
        tail = tail[:-21] if tail[-21] == 253 else tail[:-1]
        long_runs += len(tail) - len(tail.rstrip(b"\0")) >= 5
        (tmp_path / "a.zpaq").write_bytes(arc)
        rc, out = _run(STREAM, ["D", "1000", "a.zpaq", "out"], tmp_path, log=True)
        assert rc == 0, (e["name"], out)
        assert "-> GPU" in out, out
        want = data_of(e["data"]) if isinstance(e["data"], dict) else b"".join(data_of(d) for d in e["data"])
        assert (tmp_path / "out").read_bytes() == want, e["name"]
    assert long_runs >= 5
