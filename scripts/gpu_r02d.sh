#!/usr/bin/env bash
# round 2, call d: full GPU suite, K2 after the epilogue pipelining, the new bench.py at N=1, reference arm
set -uo pipefail
OUT=gpurun_out/r02d; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.log
timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -1 | tee $OUT/k2_v2.log
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 10 > $OUT/bench_n1.log 2> $OUT/bench_n1.err; tail -c 6000 $OUT/bench_n1.log; tail -5 $OUT/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_ref.log
