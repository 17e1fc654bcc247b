#!/bin/bash
# tools/r02_session7.sh -- wave encoder with the compact inlined interpreter; then the full-size benchmark lines.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02g}
L=$OUT/${R}_session.log
echo "== encoder variants, 1029 x 64 KiB, chunk 512 (compact inlined interpreter in the compact variants)" | tee $L
for v in 41 40 9 57 25; do echo "-- ZPQ_WAVE_VARIANT=$v" | tee -a $L; ZPQ_WAVE_VARIANT=$v timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
echo "== ncu wave (default variant)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_wave_kernel -s 60 -c 1 -o $OUT/${R}_wave -f python tools/prof_case.py 5 1029 65536 enc > $OUT/${R}_ncu_wave.log 2>&1
tail -2 $OUT/${R}_ncu_wave.log | tee -a $L
echo "== full-size mixed-model parity test" | tee -a $L
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "full_size" 2>&1 | tail -3 | tee -a $L
echo "== bench configs[1] (m4_256) and configs[3] (chain17_512), 2 steps" | tee -a $L
timeout 900 python bench.py --config m4_256 --steps 2 --warmup 3 > $OUT/${R}_bench_m4_256.json 2>> $L; tail -c 1800 $OUT/${R}_bench_m4_256.json | tee -a $L
timeout 900 python bench.py --config chain17_512 --steps 2 --warmup 3 > $OUT/${R}_bench_chain17_512.json 2>> $L; tail -c 1800 $OUT/${R}_bench_chain17_512.json | tee -a $L
echo "== done" | tee -a $L
