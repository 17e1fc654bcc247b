#!/bin/bash
# tools/r02_session11.sh -- final 1-GPU pass of round 2: ncu of the production wave kernel, whole -m gpu suite, smoke, the
# reference arm and the contract line exactly as the driver runs them.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02k}
L=$OUT/${R}_session.log
echo "== ncu wave (production default ZPQ_WAVE_VARIANT=57)" | tee $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_wave_kernel -s 60 -c 1 -o $OUT/${R}_wave -f python tools/prof_case.py 5 1029 65536 enc > $OUT/${R}_ncu_wave.log 2>&1
tail -2 $OUT/${R}_ncu_wave.log | tee -a $L
echo "== launch list of a small bench run (never a bench number)" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${R}_launches.csv python bench.py --blocks 64 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY' | tee -a gpurun_out/${ROUND:-r02k}_session.log
import csv, collections, os
p = "gpurun_out/%s_launches.csv" % os.environ.get("ROUND", "r02k")
tot = collections.Counter(); cnt = collections.Counter()
try:
    rows = list(csv.reader(l for l in open(p) if l.startswith('"')))
    hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
    for r in rows[1:]:
        v = float(r[iv].replace(",", "")); u = r[iu]
        ms = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else (v if u == "ms" else v * 1e3))
        k = r[ik].split("(")[0][:60]; tot[k] += ms; cnt[k] += 1
    s = sum(tot.values()) or 1
    for k, v in tot.most_common(): print("%-62s %5d launches %10.2f ms %5.1f%%" % (k, cnt[k], v, 100 * v / s))
except Exception as e:
    print("launch list not parsed:", e)
PY
echo "== smoke" | tee -a $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $L
echo "== pytest -m gpu (whole suite)" | tee -a $L
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee -a $L
echo "== reference arm as the driver runs it" | tee -a $L
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${R}_bench_ref.json ) 2>&1 | tail -4 | tee -a $L
tail -c 700 $OUT/${R}_bench_ref.json | tee -a $L
echo "== contract line as the driver runs it" | tee -a $L
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${R}_bench_1gpu.json ) 2>&1 | tail -4 | tee -a $L
tail -c 3500 $OUT/${R}_bench_1gpu.json | tee -a $L
echo "== done" | tee -a $L
