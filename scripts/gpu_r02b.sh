#!/usr/bin/env bash
# round 2, call b: K1 tiers (f32 / h16 / q8), blocked range walk, CTAs-per-SM experiment
set -uo pipefail
OUT=gpurun_out/r02b; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_merge.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.log
for rows in 10000000 1250000; do
  timeout 300 python scripts/tier_probe.py $rows 64 2>&1 | tail -1 | tee $OUT/tier_${rows}.json
  STB_SCAN_CTAS_PER_SM=1 timeout 300 python scripts/tier_probe.py $rows 64 2>&1 | tail -1 | tee $OUT/tier_${rows}_cta1.json
done
