#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/r02e; mkdir -p $OUT
timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -1 | tee $OUT/k2_v2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/k2_v2_launches.csv python scripts/batch_probe.py 10000000 1024 1 > $OUT/k2_v2_ncu.log 2>&1
grep -E "stb_batch|stb_shadow" $OUT/k2_v2_launches.csv | awk -F'","' '{print $5, $NF}' | tail -5
bash scripts/gpu_profiles_r02.sh r02p
