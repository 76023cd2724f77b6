#!/usr/bin/env bash
# round 2, call c: tickets + root-only wide list (K1), K2 pipeline v2 without atomics, fp16 shadow default
set -uo pipefail
OUT=gpurun_out/r02c; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_search.py tests/test_gpu_merge.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest.log
for rows in 10000000 1250000; do
  timeout 300 python scripts/tier_probe.py $rows 64 2>&1 | tail -1 | tee $OUT/tier_${rows}.json
done
STB_SCAN_TICKETS=0 timeout 300 python scripts/tier_probe.py 1250000 64 2>&1 | tail -1 | tee $OUT/tier_1250000_static.json
timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -2 | tee $OUT/k2_v2.log
STB_BATCH_V1=1 timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -2 | tee $OUT/k2_v1.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/k2_v2_launches.csv python scripts/batch_probe.py 10000000 1024 1 > $OUT/k2_v2_ncu.log 2>&1
grep -E "stb_batch|stb_shadow" $OUT/k2_v2_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
