#!/usr/bin/env python
"""tools/sass_lines.py obj.o 'kernel-name-substring' [top] -- static SASS instruction count per source line of one kernel
(nvdisasm -g on the sm_100a cubin inside the object).  A CPU-side proxy for the instruction-issue cost of a latency-bound
kernel when no GPU is at hand: the per-bit loops of the coders are straight-line code, so lines x trip count ~ issued
instructions.  (The real numbers come from ncu's source page; see profiles/.)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cur_fn, cur_line, counts, total = None, None, collections.Counter(), collections.Counter()
for ln in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        cur_fn = m.group(1); cur_line = None; continue
    if cur_fn is None or pat not in cur_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        counts[(cur_fn, cur_line)] += 1; total[cur_fn] += 1
for fn, n in total.items():
    print("== %s: %d SASS instructions" % (fn, n))
    rows = sorted(((c, k[1]) for k, c in counts.items() if k[0] == fn), reverse=True)[:top]
    src = {}
    for c, fl in rows:
        text = ""
        if fl:
            p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "..", "csrc", fl[0])
            if fl[0] not in src and os.path.exists(p):
                src[fl[0]] = open(p).read().splitlines()
            if fl[0] in src and fl[1] <= len(src[fl[0]]):
                text = src[fl[0]][fl[1] - 1].strip()[:110]
        print("%5d  %s:%s  %s" % (c, fl[0] if fl else "?", fl[1] if fl else "?", text))
