#!/bin/bash
# tools/cli_threads.sh -- SURVEY.md 8(f)-1 measured: the UNCHANGED zpaq CLI linked against libzpq_b200.so
# (oracle/_ref/zpaq_b200) next to the stock CLI (oracle/_ref/zpaq_ref) on the same box and the same files.
# zpaq keeps `-threads` blocks in flight (zpaq.cpp:1946, 2258); the shim's batcher turns whatever is in flight into one
# launch, so the batch size is the caller's `-threads`: the sweep below is the user-visible knob, zpaq.cpp is not touched.
# usage: tools/cli_threads.sh [MiB of input, default 512] > gpurun_out/cli_threads.log
set -u
MB=${1:-512}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d /tmp/zpqcli.XXXXXX)
REF=$ROOT/oracle/_ref/zpaq_ref
B200=$ROOT/oracle/_ref/zpaq_b200
cd "$ROOT"
python - "$W" "$MB" <<'PY'
import sys, os
sys.path.insert(0, "oracle")
import oraclepy as O
w, mb = sys.argv[1], int(sys.argv[2])
os.makedirs(os.path.join(w, "data"))
per = 64
for f in range((mb + per - 1) // per):
    with open(os.path.join(w, "data", "f%02d.txt" % f), "wb") as fh:
        for k in range(min(per, mb - f * per)):
            fh.write(O.gen_block(1, 5000 + f * 100 + k, 1 << 20))
PY
echo "== host: $(nproc) cpus, cgroup $(cat /sys/fs/cgroup/cpu.max 2>/dev/null)"; nvidia-smi --query-gpu=name --format=csv,noheader
cd "$W"
t() { local s=$(date +%s.%N); "$@" > cmd.log 2>&1; local rc=$?; local e=$(date +%s.%N); tail -2 cmd.log | tr '\n' ' '; echo; python -c "import sys; d=float(sys.argv[2])-float(sys.argv[1]); print('rc=%s seconds=%.2f MBps=%.2f ::' % (sys.argv[3], d, float(sys.argv[4])*1.048576/d), ' '.join(sys.argv[5:]))" $s $e $rc ${MBX:-$MB} "$@"; }
if [ -z "${SKIP_STOCK:-}" ]; then
echo "== stock zpaq (CPU), -method 50, -threads 16"
t $REF add ref.zpaq data -method 50 -threads 16
t $REF x ref.zpaq -to out_ref -threads 16
fi
for T in ${TS:-128 512}; do
  echo "== zpaq linked against libzpq_b200.so, -method 50, -threads $T"
  rm -rf b.zpaq out_b
  t $B200 add b.zpaq data -method 50 -threads $T
  t $B200 x b.zpaq -to out_b -threads $T
  diff -r data out_b/data > /dev/null && echo "extracted files identical to the input (T=$T)" || echo "MISMATCH (T=$T)"
done
if [ -z "${SKIP_STOCK:-}" ]; then
echo "== cross check: stock zpaq extracts the GPU-written archive"
t $REF x b.zpaq -to out_x -threads 16
diff -r data out_x/data > /dev/null && echo "stock zpaq reads the GPU-written archive: identical" || echo "MISMATCH (cross)"
fi
if [ -z "${SKIP_SUBSET:-}" ]; then
echo "== small subset at the CLI's default depth (-threads 16, 64 MiB): one launch per 31 blocks"
mkdir sub && cp data/f00.txt sub/
MBX=64; t $B200 add s.zpaq sub -method 50 -threads 16
fi
ls -l ref.zpaq b.zpaq 2>/dev/null | awk '{print $5, $9}'
rm -rf "$W"
