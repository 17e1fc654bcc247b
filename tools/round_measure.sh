#!/bin/bash
# tools/round_measure.sh -- ONE gpurun call that collects everything the next round needs first (about 25 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round_measure.sh'
# 1. parity gate   2. every kernel switch timed on one batch   3. bench line   4. launch list   5. one ncu --set full
# capture of each production kernel (small batch, per-SM occupancy of the benchmark), with the per-line summaries.
# Everything lands in gpurun_out/ (copy what is to be judged into profiles/).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02}
echo "== parity gate" | tee $OUT/${R}_measure.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee -a $OUT/${R}_measure.log
echo "== encoder switches (ZPQ_WAVE_VARIANT), 1029 blocks x 256 KiB" | tee -a $OUT/${R}_measure.log
timeout 600 python tools/enc_variants.py 1029 262144 1,9,25,29 2>&1 | tee -a $OUT/${R}_measure.log
for g in 8 32; do echo "== ZPQ_GROUP=$g" | tee -a $OUT/${R}_measure.log; ZPQ_GROUP=$g timeout 300 python tools/enc_variants.py 1029 262144 25 2>&1 | grep "rep 1" | tee -a $OUT/${R}_measure.log; done
echo "== decoder switches (ZPQ_DEC_VARIANT), 1029 blocks x 128 KiB" | tee -a $OUT/${R}_measure.log
timeout 900 python tools/dec_variants.py 1029 131072 3,7,15,31 2>&1 | tee -a $OUT/${R}_measure.log
echo "== bench" | tee -a $OUT/${R}_measure.log
timeout 900 python bench.py --steps 1 --warmup 1 > $OUT/${R}_bench_1gpu.json 2>> $OUT/${R}_measure.log; tail -c 1500 $OUT/${R}_bench_1gpu.json | tee -a $OUT/${R}_measure.log
echo "== launch list (never a bench number)" | tee -a $OUT/${R}_measure.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/${R}_launches.csv python bench.py --blocks 64 --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY' | tee -a gpurun_out/${ROUND:-r02}_measure.log
import csv, collections, os, sys
p = "gpurun_out/%s_launches.csv" % os.environ.get("ROUND", "r02")
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
echo "== ncu --set full: one wave launch of the encoder, the two-warp decoder with 7 blocks in one CTA" | tee -a $OUT/${R}_measure.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zpq_wave_kernel -s 20 -c 1 -o $OUT/${R}_wave -f python tools/prof_case.py 5 1029 65536 enc > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_kernel6 -c 1 -o $OUT/${R}_dec6 -f python tools/prof_case.py 5 7 16384 both > /dev/null 2>&1
for k in wave dec6; do
  if [ -f $OUT/${R}_$k.ncu-rep ]; then
    { ncu -i $OUT/${R}_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','sm__warps_active.avg.per_cycle_active','smsp__average_warp_latency_per_inst_issued.ratio','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for n in want:
    if n in h: print(n, '[%s] =' % rows[1][h.index(n)], v[h.index(n)])
for i,n in enumerate(h):
    if 'issue_stalled' in n and n.endswith('.ratio'):
        try:
            if float(v[i])>0.15: print('stall', n.split('issue_stalled_')[-1].replace('_per_warp_active.pct','').replace('.ratio',''), '=', v[i])
        except ValueError: pass
"; python tools/ncu_lines.py $OUT/${R}_$k.ncu-rep 25; } > $OUT/${R}_$k.summary.txt 2>&1
    head -30 $OUT/${R}_$k.summary.txt | tee -a $OUT/${R}_measure.log
  fi
done
echo "== done" | tee -a $OUT/${R}_measure.log
