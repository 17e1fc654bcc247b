#!/bin/bash
# tools/r02_session13.sh -- controlled comparison on ONE box: encoder variants through tools/enc_variants.py and the two
# entry points through tools/enc_paths.py (the benchmark reported 84 MB/s where an earlier session's enc_variants run had
# reported 99.7 for the same batch); decoder after the CM / CONST lanes moved to the mixer warp; the bench with the
# shortened warm-up steps.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02n}
L=$OUT/${R}_session.log
nvidia-smi --query-gpu=name,serial,uuid,clocks.max.sm,clocks.max.mem,power.limit,temperature.gpu --format=csv | tee $L
echo "== enc_variants 57,41 at the benchmark's size" | tee -a $L
timeout 600 python tools/enc_variants.py 1029 1044480 57,41 2>&1 | tee -a $L
echo "== enc_paths (default variant, both entry points)" | tee -a $L
timeout 600 python tools/enc_paths.py 2>&1 | tail -4 | tee -a $L
echo "== decoder 1029 x 32 KiB, method 5 and 4: kernel 7 (CM / CONST in the mixer warp) and kernel 6" | tee -a $L
for k in 7 6; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 2>&1 | grep "rep 1" | tee -a $L; done
for k in 7 6; do ZPQ_DEC_KERNEL=$k timeout 300 python tools/dec_variants.py 1029 32768 31 4 2>&1 | grep "rep 1" | tee -a $L; done
echo "== bench, 3 steps, 5 warm-up (shortened warm-up steps)" | tee -a $L
( time timeout 900 python bench.py --gpus 1 --steps 3 --warmup 5 > $OUT/${R}_bench_1gpu.json ) 2>&1 | tail -4 | tee -a $L
tail -c 3500 $OUT/${R}_bench_1gpu.json | tee -a $L
echo "== done" | tee -a $L
