#!/bin/bash
# tools/r02_session12.sh -- 4 GPUs: the contract line (weak scaling, 3 steps) and the sharded data path at the benchmark's
# shard size (1 029 blocks per GPU = BASELINE configs[4]'s 4 GiB archive decompressed across the GPUs of the box).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p $OUT
R=${ROUND:-r02l}
L=$OUT/${R}_session.log
N=${NGPU:-4}
nvidia-smi -L | tee $L
echo "== bench --gpus $N (contract line), 3 steps" | tee -a $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${R}_bench_${N}gpu.json 2>> $L; tail -c 2600 $OUT/${R}_bench_${N}gpu.json | tee -a $L
echo "== bench --mode sharded --gpus $N, 1029 blocks per GPU (4 GiB), 1 step" | tee -a $L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --mode sharded --steps 1 --warmup 1 > $OUT/${R}_bench_sharded_${N}gpu.json 2>> $L; tail -c 2200 $OUT/${R}_bench_sharded_${N}gpu.json | tee -a $L
echo "== done" | tee -a $L
