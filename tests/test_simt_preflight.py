"""CPU tier: the LOGIC of the production CUDA kernels, checked bit for bit against the oracle on the SIMT emulator
(tools/simt: the .cu sources compiled with g++, every CUDA thread a fiber).  This is a pre-flight, not the parity gate:
the `-m gpu` tests on a B200 are.  The emulator build is test infrastructure; nothing in zpaq_b200/ can load it."""
import os
import shutil
import subprocess
import sys

import pytest

from util import ROOT

SIMT = os.path.join(ROOT, "tools", "simt")


@pytest.fixture(scope="module")
def simt_build():
    if shutil.which("g++") is None or os.uname().machine != "x86_64":
        pytest.skip("the emulator needs g++ on x86-64")
    r = subprocess.run([sys.executable, os.path.join(SIMT, "build.py")], capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        pytest.fail("tools/simt/build.py failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return os.path.join(SIMT, "_build", "libzpq_simt.so")


@pytest.mark.timeout(900)
def test_production_kernels_on_the_emulator(simt_build):
    env = dict(os.environ, SIMT_WATCHDOG="600")
    r = subprocess.run([sys.executable, os.path.join(SIMT, "run.py"), "script", os.path.join(ROOT, "tests", "simt_preflight_case.py")],
                       capture_output=True, text=True, timeout=800, env=env)
    assert r.returncode == 0 and "simt preflight ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.timeout(900)
def test_sharded_data_path_two_ranks_gloo(simt_build):
    """zpaq_b200.shard.compress_sharded / decompress_sharded over gloo, world size 2: scatter of raw blocks as tensors, the
    device-resident whole-block entry point (SHA-1 and level-5 statistics on the "device"), gather of coded blocks; the
    archive must equal the unmodified reference's block by block and decode back to the input."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "simt_shard_case.py"), simt_build],
                       capture_output=True, text=True, timeout=800, env=dict(os.environ, SIMT_WATCHDOG="600"))
    assert r.returncode == 0 and "simt shard ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.timeout(600)
def test_emulator_selftest():
    """The emulator's own semantics (shuffles, ballots, reductions, barriers, named barriers, spin hand-shakes between
    warps, exited lanes, atomics) against closed-form answers, independent of the product kernels."""
    if shutil.which("g++") is None or os.uname().machine != "x86_64":
        pytest.skip("the emulator needs g++ on x86-64")
    r = subprocess.run([sys.executable, os.path.join(SIMT, "selftest", "selftest.py")], capture_output=True, text=True, timeout=500)
    assert r.returncode == 0 and "simt selftest ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
