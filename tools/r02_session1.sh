#!/bin/bash
# tools/r02_session1.sh -- first GPU session of round 2: profile the kernels the round-1 benchmark actually ran
# (wave variant 25, decoder variant 31): per-role cycles of the encoder, ncu --set full + source page of one wave launch
# and of the decoder at the benchmark's per-SM occupancy, quick group / variant timings.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=r02a
L=$OUT/${R}_session.log
echo "== nvidia-smi" | tee $L
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee -a $L
nproc | tee -a $L; cat /sys/fs/cgroup/cpu.max 2>/dev/null | tee -a $L
echo "== encoder per-role cycles, 1029 x 64 KiB" | tee -a $L
ZPQ_WAVE_PROF=1 timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L
for g in 8 32; do echo "== ZPQ_GROUP=$g" | tee -a $L; ZPQ_GROUP=$g timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
for c in 512 1024 4096; do echo "== ZPQ_CHUNK=$c" | tee -a $L; ZPQ_CHUNK=$c timeout 300 python tools/prof_case.py 5 1029 65536 enc 2>&1 | tee -a $L; done
echo "== decoder 1029 x 32 KiB, variants" | tee -a $L
timeout 600 python tools/dec_variants.py 1029 32768 15,31 2>&1 | tee -a $L
echo "== ncu wave" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_wave_kernel -s 24 -c 1 -o $OUT/${R}_wave -f python tools/prof_case.py 5 1029 65536 enc > $OUT/${R}_ncu_wave.log 2>&1
echo "== ncu dec6" | tee -a $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zpq_dec_kernel6 -c 1 -o $OUT/${R}_dec6 -f python tools/prof_case.py 5 7 16384 both > $OUT/${R}_ncu_dec.log 2>&1
ls -la $OUT | tee -a $L
echo "== done" | tee -a $L
