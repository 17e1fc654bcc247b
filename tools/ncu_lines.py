"""tools/ncu_lines.py REPORT.ncu-rep [topN] -- per-CUDA-source-line stall samples and executed instructions."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []; fname = ""
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) > 8 and r[0].isdigit():
        ix = {k: i for i, k in enumerate(hdr)}
        try:
            lines.append((fname, int(r[0]), r[1].strip(), int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])))
        except ValueError:
            pass
ts = sum(l[3] for l in lines) or 1; ti = sum(l[4] for l in lines) or 1
print("total samples %d, warp instructions %d" % (ts, ti))
print("--- by stall samples")
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%d  %s" % (100 * l[3] / ts, 100 * l[4] / ti, l[0], l[1], l[2][:100]))
