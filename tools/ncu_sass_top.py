#!/usr/bin/env python
"""tools/ncu_sass_top.py REPORT.ncu-rep [N] [file-substring] -- the N SASS instructions of a capture with the most warp-stall samples,
each with the six instructions before it and its stall reasons (which load / barrier a hot line actually waits on)."""
import csv, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 25; want = sys.argv[3] if len(sys.argv) > 3 else ""
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Line No" and len(r) > 8] + [len(rows)]
def num(x):
    try: return int(x)
    except ValueError: return 0
K = ('stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_no_inst', 'stall_math', 'stall_lg', 'stall_mio',
     'stall_not_selected', 'stall_selected', 'stall_barrier', 'stall_membar', 'stall_dispatch')
for si in range(len(secs) - 1):
    s = secs[si]; hdr = rows[s]; ix = {k: i for i, k in enumerate(hdr)}
    fname = rows[s - 1][1] if s >= 1 and len(rows[s - 1]) > 1 else "?"
    if want not in fname: continue
    cur = None; sass = []
    for r in rows[s + 1:secs[si + 1]]:
        if len(r) != len(hdr): continue
        if r[0] != '': cur = (r[0], r[1])
        if r[2] not in ('-', ''): sass.append((cur, r))
    tot = sum(num(r[ix['# Samples']]) for c, r in sass) or 1
    print("==== %s: %d SASS instructions, %d samples" % (fname, len(sass), tot))
    order = sorted(range(len(sass)), key=lambda i: -num(sass[i][1][ix['# Samples']]))
    for i in order[:N]:
        c, r = sass[i]
        print('--- %5.2f%% line %s: %s' % (100 * num(r[ix['# Samples']]) / tot, c[0], c[1].strip()[:100]))
        for j in range(max(0, i - 6), i + 1):
            cc, rr = sass[j]
            rs = ' '.join('%s=%d' % (k[6:], num(rr[ix[k]])) for k in K if k in ix and num(rr[ix[k]])) if j == i else ''
            print('     %s %-72s smp %5s  %s' % ('>>' if j == i else '  ', rr[3].strip()[:72], rr[ix['# Samples']], rs))
