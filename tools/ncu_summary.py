#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep "title" [block_bits] -- text summary of one `ncu --set full` capture for profiles/:
launch metrics, stall reasons per issued instruction, DRAM traffic, and the source lines with the most stall samples /
executed instructions (needs -lineinfo builds and --import-source on)."""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
bits = float(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); h = rows[0]; v = rows[-1]
print(title)
want = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
vals = {}
for n in want:
    if n in h:
        print(n, '[%s] =' % rows[1][h.index(n)], v[h.index(n)]); vals[n] = (rows[1][h.index(n)], v[h.index(n)])
for i, n in enumerate(h):
    if 'issue_stalled' in n and n.endswith('.ratio'):
        try:
            if float(v[i]) > 0.10:
                print('stall', n.split('issue_stalled_')[-1].replace('_per_warp_active.pct', '').replace('.ratio', ''), '=', v[i], 'cycles per issued instruction')
        except ValueError:
            pass
def to_bytes(u, x):
    x = float(x.replace(",", "")); return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
if bits and 'dram__bytes_read.sum' in vals:
    tr = to_bytes(*vals['dram__bytes_read.sum']) + to_bytes(*vals['dram__bytes_write.sum'])
    print("derived: %.0f block-bits in this launch; DRAM traffic %.1f B per block-bit = %.0f B per input byte; %.0f warp instructions per block-bit"
          % (bits, tr / bits, 8 * tr / bits, float(vals['smsp__inst_executed.sum'][1].replace(",", "")) / bits))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); lines = []; fname = ""; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        ix = {k: i for i, k in enumerate(hdr)}
        try: lines.append((fname, int(r[0]), r[1].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])))
        except ValueError: pass
ts = sum(l[3] for l in lines) or 1; ti = sum(l[4] for l in lines) or 1
print("source page: %d stall samples, %d warp instructions" % (ts, ti))
print("--- top lines by stall samples")
for l in sorted(lines, key=lambda x: -x[3])[:14]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100 * l[3] / ts, 100 * l[4] / ti, l[0], l[1], l[2][:100]))
print("--- top lines by executed instructions")
for l in sorted(lines, key=lambda x: -x[4])[:10]:
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100 * l[4] / ti, 100 * l[3] / ts, l[0], l[1], l[2][:100]))
