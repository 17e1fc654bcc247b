#!/usr/bin/env python
"""tools/simt/build.py -- build tools/simt/_build/libzpq_simt.so: the repo's CUDA sources (zpaq_b200/csrc) translated for
g++ and linked with the SIMT emulator (tools/simt/simt.cpp).  TEST INFRASTRUCTURE ONLY: a pre-flight check of kernel logic
on a machine without a GPU; the product never loads this library (see tools/simt/include/simt.h).

The translation is textual and touches only what g++ cannot parse:
  kernel<<<grid, block, smem, stream>>>(args);    ->  simt_launch(dim3(grid), dim3(block), smem, [&]() { kernel(args); });
  extern __shared__ ... unsigned char smem[];     ->  unsigned char* smem = simt_dyn_smem();
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(count) : "memory")   ->  simt_named_barrier(id, count)
  asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(count) : "memory") ->  simt_named_arrive(id, count)
  asm volatile("prefetch.global.L1|L2 [%0];" :: "l"(p))                ->  (void)(p)
  while (cond) { }      (spin on a volatile shared word)               ->  while (cond) { simt_spin(); }
  for (...; cond;) { if (++spin > SPIN_CAP) ... }   (capped spin)      ->  ... { simt_spin(); if (++spin > SPIN_CAP) ... }
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "zpaq_b200", "csrc")
OUT = os.path.join(HERE, "_build")

CU = ["zpq_kernels.cu", "zpq_kernels2.cu", "zpq_kernels5.cu", "zpq_kernels6.cu", "zpq_kernels7.cu", "zpq_kernels8.cu", "zpq_prep.cu", "zpq_bwt.cu", "zpq_lz77.cu", "zpq_abi.cu"]
CPP = ["zpq_model.cpp", "zpq_method.cpp", "zpq_pcomp.cpp", "zpq_frame.cpp", "zpq_hcomp.cpp"]


def split_top(s):
    """split on commas that are not inside parentheses / brackets / angle-free template args"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


LAUNCH = re.compile(r"([A-Za-z_][A-Za-z_0-9]*(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);")


def translate(text, name):
    def launch(m):
        cfg = split_top(m.group(2))
        while len(cfg) < 4:
            cfg.append("0")
        return "simt_launch(dim3(%s), dim3(%s), (size_t)(%s), [&]() { %s(%s); });" % (cfg[0], cfg[1], cfg[2], m.group(1), m.group(3))
    text, n_launch = LAUNCH.subn(launch, text)
    text, n_smem = re.subn(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];", r"unsigned char* \1 = simt_dyn_smem();", text)
    text, n_bar = re.subn(r'asm volatile\("bar\.sync %0, %1;"\s*::\s*"r"\((.*?)\),\s*"r"\((\w+)\)\s*:\s*"memory"\)', r"simt_named_barrier(\1, \2)", text)
    text, n_arr = re.subn(r'asm volatile\("bar\.arrive %0, %1;"\s*::\s*"r"\((.*?)\),\s*"r"\((\w+)\)\s*:\s*"memory"\)', r"simt_named_arrive(\1, \2)", text)
    n_bar += n_arr
    text, n_pf = re.subn(r'asm volatile\("prefetch\.global\.L[12] \[%0\];"\s*::\s*"l"\((\w+)\)\)', r"(void)(\1)", text)
    text, n_spin = re.subn(r"(while \(.*\)) \{ \}", r"\1 { simt_spin(); }", text)
    text, n_spin2 = re.subn(r"\{ if \(\+\+(spin_?) > SPIN_CAP\)", r"{ simt_spin(); if (++\1 > SPIN_CAP)", text)     # capped spin loops
    n_spin += n_spin2
    text = text.replace("__noinline__", "__attribute__((noinline))")
    if "asm" in re.sub(r"//.*", "", text):
        raise SystemExit("%s: untranslated inline asm left" % name)
    if "<<<" in text:
        raise SystemExit("%s: untranslated kernel launch left" % name)
    return text, (n_launch, n_smem, n_bar, n_pf, n_spin)


def main():
    os.makedirs(OUT, exist_ok=True)
    opt = os.environ.get("SIMT_OPT", "-O2")
    objs = []
    cmds = []
    for f in sorted(os.listdir(SRC)):
        if f.endswith((".h", ".cuh")):
            t, _ = translate(open(os.path.join(SRC, f)).read(), f)
            open(os.path.join(OUT, f), "w").write(t)
    for f in CU + CPP:
        t, counts = translate(open(os.path.join(SRC, f)).read(), f)
        dst = os.path.join(OUT, f + ".cpp")
        open(dst, "w").write('#include "simt.h"\n#line 1 "%s"\n' % f + t)
        obj = os.path.join(OUT, f + ".o")
        objs.append(obj)
        cmds.append(["g++", "-std=c++17", opt, "-g", "-fPIC", "-fno-strict-aliasing", "-DZPQ_SIMT_EMU=1", "-w",
                     "-I", os.path.join(HERE, "include"), "-I", OUT, "-I", SRC, "-c", dst, "-o", obj])
        print("%-18s launches %d, smem decls %d, named barriers %d, prefetches %d, spin loops %d" % ((f,) + counts))
    obj = os.path.join(OUT, "simt.o")
    objs.append(obj)
    cmds.append(["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-I", os.path.join(HERE, "include"), "-c", os.path.join(HERE, "simt.cpp"), "-o", obj])
    procs = [subprocess.Popen(c) for c in cmds]
    bad = [c for c, p in zip(cmds, procs) if p.wait() != 0]
    if bad:
        raise SystemExit("compile failed: " + " ".join(bad[0]))
    lib = os.path.join(OUT, "libzpq_simt.so")
    # -Bsymbolic: the library's calls to its own cuda* emulation must not be interposed by a real libcudart that torch loaded first
    subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic", "-o", lib + ".new"] + objs + ["-lpthread"])
    os.replace(lib + ".new", lib)          # a process that has the old build loaded keeps its mapping
    print("built", lib)
    # the unchanged zpaq.cpp + the shim, linked against the emulator build instead of the CUDA build (objects from `make -C oracle cli`)
    ref = os.path.join(ROOT, "oracle", "_ref")
    need = [os.path.join(ref, f) for f in ("zpaq.o", "libzpaq_shim.o", "libzpaq_renamed.o")]
    if all(os.path.exists(f) for f in need):
        exe = os.path.join(OUT, "zpaq_simt")
        subprocess.check_call(["g++", "-pthread", "-o", exe] + need + ["-L", OUT, "-lzpq_simt", "-Wl,-rpath," + OUT])
        print("built", exe)
    need = [os.path.join(ref, f) for f in ("shim_stream.o", "libzpaq_shim.o", "libzpaq_renamed.o")]
    if all(os.path.exists(f) for f in need):
        exe = os.path.join(OUT, "shim_stream_simt")
        subprocess.check_call(["g++", "-pthread", "-o", exe] + need + ["-L", OUT, "-lzpq_simt", "-Wl,-rpath," + OUT])
        print("built", exe)


if __name__ == "__main__":
    main()
