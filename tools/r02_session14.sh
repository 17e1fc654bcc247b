#!/bin/bash
# tools/r02_session14.sh -- wave encoder with the four table entries of a nibble requested up front (compact variants),
# against the previous loop (variant 25) and against the library built from commit 7adee48 (the one profiles/r02h measured
# at 99.7 MB/s on another box) on the SAME box.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02o}
L=$OUT/${R}_session.log
nvidia-smi --query-gpu=name,serial,clocks.max.sm,clocks.max.memory,power.limit,temperature.gpu --format=csv | tee $L
echo "== this commit: variants 57 (nibble gather), 25 (previous loop, not compact)" | tee -a $L
timeout 600 python tools/enc_variants.py 1029 1044480 57,25,57 2>&1 | tee -a $L
echo "== library of commit 7adee48: variant 57" | tee -a $L
ZPQ_AB_LIB=_ab/libzpq_b200_7adee48.so timeout 600 python tools/enc_variants.py 1029 1044480 57 2>&1 | tee -a $L
echo "== parity" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4 | tee -a $L
echo "== done" | tee -a $L
