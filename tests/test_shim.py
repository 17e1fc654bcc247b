"""The UNCHANGED zpaq.cpp linked against libzpq_b200.so through zpaq_b200/csrc/libzpaq_shim.cpp
(oracle/Makefile target `cli`; the binaries are built where /root/reference exists and travel to the GPU box)."""
import os
import subprocess

import pytest

import oraclepy as O
from util import ROOT

B200 = os.path.join(ROOT, "oracle", "_ref", "zpaq_b200")
REF = os.path.join(ROOT, "oracle", "_ref", "zpaq_ref")
need_cli = pytest.mark.skipif(not (os.path.exists(B200) and os.path.exists(REF)), reason="CLI binaries not built (make -C oracle cli)")


def _files(tmp_path):
    names = []
    for i, kind in enumerate((0, 1, 3)):
        p = tmp_path / ("f%d.bin" % i)
        p.write_bytes(O.gen_block(kind, 900 + i, 150000 + 7000 * i))
        names.append(p.name)
    return names


def _run(exe, args, cwd):
    r = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
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
@pytest.mark.gpu
@pytest.mark.parametrize("method", ["4", "5"])
def test_zpaq_add_on_gpu_extracts_with_stock_zpaq(tmp_path, method):
    """`zpaq add -method 4/5` through the GPU library; the archive is read back by the stock reference CLI."""
    names = _files(tmp_path)
    rc, out = _run(B200, ["add", "g.zpaq"] + names + ["-method", method, "-threads", "8"], tmp_path)
    assert rc == 0, out
    rc, out = _run(REF, ["x", "g.zpaq", "-to", "og"], tmp_path)
    assert rc == 0, out
    for n in names:
        assert (tmp_path / "og" / n).read_bytes() == (tmp_path / n).read_bytes()
    # same input through the stock CLI gives the same compressed size (bytes differ only in the dates zpaq stores)
    rc, out = _run(REF, ["add", "r.zpaq"] + names + ["-method", method, "-threads", "8"], tmp_path)
    assert rc == 0, out
    assert abs((tmp_path / "g.zpaq").stat().st_size - (tmp_path / "r.zpaq").stat().st_size) <= 64
