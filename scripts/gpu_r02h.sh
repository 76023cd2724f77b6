#!/usr/bin/env bash
set -uo pipefail
OUT=gpurun_out/r02h; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest.log
for rows in 10000000 1250000; do timeout 300 python scripts/tier_probe.py $rows 64 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print(d['rows'], {t: (round(d[t]['us_per_query_pipelined'], 1), round(d[t]['e2e_ms'] * 1e3, 1), d[t]['proven']) for t in ('f32', 'h16', 'q8')})" | tee $OUT/tier_$rows.log; done
