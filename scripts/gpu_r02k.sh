#!/usr/bin/env bash
# 1 GPU: full suite on the latest tree, overlapped launch mode experiment, K2 + modes + K5 probes
set -uo pipefail
OUT=gpurun_out/r02k; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/pytest.log
STB_SCAN_OVERLAP=1 timeout 900 python -m pytest tests/test_gpu_search.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_overlap.log
fmt='import sys, json; d = json.loads(sys.stdin.read()); print(d["rows"], {t: (round(d[t]["us_per_query_pipelined"], 1), round(d[t]["e2e_ms"] * 1e3, 1), d[t]["proven"]) for t in ("f32", "h16", "q8")})'
for rows in 10000000 1250000 312500; do
  timeout 300 python scripts/tier_probe.py $rows 128 2>&1 | tail -1 | python -c "$fmt" | tee $OUT/tier_$rows.log
  STB_SCAN_OVERLAP=1 timeout 300 python scripts/tier_probe.py $rows 128 2>&1 | tail -1 | python -c "$fmt" | tee $OUT/tier_${rows}_overlap.log
done
timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -1 | tee $OUT/k2.log
