#!/usr/bin/env bash
# 1 GPU: new tests + the N=2 bench code path in one-GPU functional mode (gloo + IPC on cuda:0; timings void)
set -uo pipefail
OUT=gpurun_out/r02i; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_ivfpq.py -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest.log
STB_BENCH_ONE_GPU=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 3 --config4-rows 40000000 --ivfpq-rows-per-gpu 2000000 > $OUT/bench_n2_onegpu.log 2> $OUT/bench_n2_onegpu.err; tail -c 5000 $OUT/bench_n2_onegpu.log; tail -8 $OUT/bench_n2_onegpu.err
timeout 300 python scripts/modes_probe.py 10000000 2>&1 | tail -1 | tee $OUT/modes_q8.log
STB_SCAN_TIER=f32 timeout 300 python scripts/modes_probe.py 10000000 2>&1 | tail -1 | tee $OUT/modes_f32.log
timeout 300 python scripts/ivfpq_probe.py 4000000 4096 2>&1 | grep '"nprobe": 64' | tee $OUT/k5.log
