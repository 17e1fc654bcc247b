#!/bin/bash
# tools/r02_session15.sh -- wave encoder, compact variants: four entries of a nibble requested up front + context hash of
# the next byte evaluated behind that request + L2 prefetch of the next byte's hash rows; against variant 25 (previous
# loop) on the same box; ncu of the new loop.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02p}
L=$OUT/${R}_session.log
nvidia-smi --query-gpu=name,serial,clocks.max.sm,clocks.max.memory,power.limit,temperature.gpu --format=csv | tee $L
echo "== variants 57 (new loop), 25 (previous loop) at the benchmark's size" | tee -a $L
timeout 600 python tools/enc_variants.py 1029 1044480 57,25 2>&1 | tee -a $L
echo "== ncu wave (variant 57)" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_wave_kernel -s 60 -c 1 -o $OUT/${R}_wave -f python tools/prof_case.py 5 1029 65536 enc > $OUT/${R}_ncu_wave.log 2>&1
tail -2 $OUT/${R}_ncu_wave.log | tee -a $L
echo "== done" | tee -a $L
