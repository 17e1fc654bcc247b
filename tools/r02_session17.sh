#!/bin/bash
# r02s: final code -- whole GPU suite, contract bench (short), BWT bench line, ncu of the radix scatter pass + launch list of the
# pre-processing kernels, the CLI through the shim with the 30 s arrival window
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
L=$OUT/r02s_session.log
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,serial --format=csv,noheader
echo "== pytest -m gpu (whole suite)"
( time timeout 1300 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1
echo "== bench, contract line, 3 steps 3 warm-up"
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 > $OUT/r02s_bench_1gpu.json 2> $OUT/r02s_bench_err.log; tail -c 1800 $OUT/r02s_bench_1gpu.json
echo "== bench --config bwt_256"
timeout 400 python bench.py --config bwt_256 --steps 3 --warmup 2 > $OUT/r02s_bench_bwt_256.json 2>> $OUT/r02s_bench_err.log; cat $OUT/r02s_bench_bwt_256.json
tail -5 $OUT/r02s_bench_err.log
echo "== ncu: one radix scatter pass, 64 x 128 KiB text blocks"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:zpq_radix_scatter -s 6 -c 1 -f -o $OUT/r02s_radix_scatter python tools/bwt_prof.py 64 > $OUT/r02s_ncu.log 2>&1
tail -3 $OUT/r02s_ncu.log
python tools/ncu_summary.py $OUT/r02s_radix_scatter.ncu-rep "ncu --set full: zpq_radix_scatter (one 8-bit pass of the suffix sorter, 64 x 130 560 B 'text' blocks = 8.4 M suffixes, round 0), final code" > $OUT/r02s_radix_scatter.summary.txt 2>&1
head -30 $OUT/r02s_radix_scatter.summary.txt
echo "== ncu launch list of the pre-processing kernels (same command)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'zpq_(radix|scan|bwt|ibwt)' -c 600 --csv --log-file $OUT/r02s_bwt_launches.csv python tools/bwt_prof.py 64 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02s_bwt_launches.csv")) if len(r) > 10]
h = rows[0]; ik = h.index("Kernel Name"); iv = h.index("Metric Value"); iu = h.index("Metric Unit")
t = collections.defaultdict(float); c = collections.Counter()
SC = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
for r in rows[1:]:
    try: t[r[ik].split("(")[0]] += float(r[iv].replace(",", "")) * SC.get(r[iu], 1.0); c[r[ik].split("(")[0]] += 1
    except ValueError: pass
tot = sum(t.values()) or 1
for k, v in sorted(t.items(), key=lambda x: -x[1]): print("%-28s %5d launches %10.1f us %5.1f%%" % (k, c[k], v / 1e3, 100 * v / tot))
PY
echo "== CLI through the shim, -threads 512, arrival window 100 ms / 30 s"
TS=512 SKIP_SUBSET=1 SKIP_STOCK=1 timeout 600 bash tools/cli_threads.sh 512 2>&1
echo "== done"
} > $L 2>&1
tail -80 $L
