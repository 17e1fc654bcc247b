#!/usr/bin/env python
"""tools/simt/selftest/selftest.py -- build selftest.cu with the emulator's translator and run it (exit code 0 = ok)."""
import ctypes
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.dirname(HERE)
sys.path.insert(0, SIMT)
import build as B   # the product translator


def main():
    out = os.path.join(SIMT, "_build", "selftest")
    os.makedirs(out, exist_ok=True)
    text, _ = B.translate(open(os.path.join(HERE, "selftest.cu")).read(), "selftest.cu")
    src = os.path.join(out, "selftest.cu.cpp")
    open(src, "w").write('#include "simt.h"\n#line 1 "selftest.cu"\n' + text)
    lib = os.path.join(out, "libsimt_selftest.so")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-w", "-shared", "-I", os.path.join(SIMT, "include"),
                           src, os.path.join(SIMT, "simt.cpp"), "-o", lib, "-lpthread"])
    os.environ.setdefault("SIMT_WATCHDOG", "60")      # a scheduling bug shows up as a hang: dump the scheduler state and exit
    rc = ctypes.CDLL(lib).simt_selftest()
    sys.exit(int(rc != 0))


if __name__ == "__main__":
    main()
