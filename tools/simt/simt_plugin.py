"""pytest plugin used by tools/simt/run.py: points the ctypes binding at the SIMT emulator build in EVERY pytest process
(pytest-xdist workers import zpaq_b200 afresh).  Test infrastructure only; the package itself has no such switch."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
# the CLI tests run the unchanged zpaq.cpp / the stream driver linked against the emulator build (tools/simt/build.py)
os.environ.setdefault("ZPQ_CLI_B200", os.path.join(HERE, "_build", "zpaq_simt"))
os.environ.setdefault("ZPQ_CLI_STREAM", os.path.join(HERE, "_build", "shim_stream_simt"))
os.environ.setdefault("ZPQ_TEST_TIMEOUT", "14400")      # the emulator is ~10^4 times slower than the GPU
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    import zpaq_b200
    zpaq_b200.LIB_PATH = os.path.join(HERE, "_build", "libzpq_simt.so")
    zpaq_b200._lib = None
    zpaq_b200.lib()
