#!/usr/bin/env bash
# 2 GPUs: new GPU tests (cuda:0 only) + bench at N=2
set -uo pipefail
OUT=gpurun_out/r02g; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 > $OUT/bench_n2.log 2> $OUT/bench_n2.err; tail -c 7000 $OUT/bench_n2.log; tail -5 $OUT/bench_n2.err
