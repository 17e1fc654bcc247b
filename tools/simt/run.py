#!/usr/bin/env python
"""tools/simt/run.py -- run the repo's GPU checks against the SIMT emulator build (tools/simt/build.py) on a machine
without a GPU.  TEST INFRASTRUCTURE ONLY: a pre-flight check of kernel logic; the real gate is `pytest -m gpu` on a B200.

  python tools/simt/run.py smoke                     # __graft_entry__.smoke()
  python tools/simt/run.py pytest [pytest args]      # pytest -m gpu with the emulator library (slow: pick tests with -k)
  python tools/simt/run.py script path.py [args]     # any script that uses zpaq_b200

This is the ONLY place that points the ctypes binding at the emulator library, and it does so by overwriting
zpaq_b200.LIB_PATH in this process: the package itself has no switch, environment variable or fallback for it.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
LIB = os.path.join(HERE, "_build", "libzpq_simt.so")


def main():
    if not os.path.exists(LIB):
        raise SystemExit("build the emulator library first: python tools/simt/build.py")
    import zpaq_b200
    zpaq_b200.LIB_PATH = LIB
    zpaq_b200.lib()          # bind now, before anything else can load the CUDA build
    what = sys.argv[1] if len(sys.argv) > 1 else "smoke"
    if what == "smoke":
        import __graft_entry__ as g
        g.smoke()
    elif what == "pytest":
        import pytest
        sys.path.insert(0, HERE)
        os.environ["PYTHONPATH"] = HERE + os.pathsep + os.environ.get("PYTHONPATH", "")      # xdist workers load the plugin too
        sys.exit(pytest.main(["-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-p", "simt_plugin", os.path.join(ROOT, "tests")] + sys.argv[2:]))
    elif what == "script":
        sys.argv = sys.argv[2:]
        runpy.run_path(sys.argv[0], run_name="__main__")
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
