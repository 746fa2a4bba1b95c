#!/bin/bash
# The 2 / 4 / 8 GPU lines of the headline workload on ONE box (run under `gpurun --gpus 8`):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_scaling.sh r02z'
# One rank per GPU, launched the way the driver launches bench.py.
TAG=${1:-scale}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpus_$TAG.txt 2>&1
PORT=29511
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
    --master-port $PORT bench.py --gpus $n --no-cpu-baseline > $OUT/bench_${TAG}_n$n.json 2> $OUT/bench_${TAG}_n$n.err
  PORT=$((PORT + 1))
done
ls -la $OUT | tail -8
