#!/usr/bin/env bash
# N GPUs (arg 1): bench at that N
set -uo pipefail
N=${1:-8}
OUT=gpurun_out/r02j; mkdir -p $OUT
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 200 --warmup 10 > $OUT/bench_n$N.log 2> $OUT/bench_n$N.err; tail -c 7000 $OUT/bench_n$N.log; tail -5 $OUT/bench_n$N.err
