#!/usr/bin/env bash
# First gpurun call of round 2: time and validate the two opt-in paths written at the end of
# round 1 (K2 pipeline v2, fused IVF-PQ search v2) next to the default ones.
#   bash scripts/gpu_round2_first.sh [tag]      (from the repo root on the GPU box)
set -uo pipefail
TAG=${1:-r02_v2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== gated tests (v2 paths)"
STB_TEST_V2=1 timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_ivfpq.py tests/test_host_cpp.py tests/test_gpu_search.py -x -q -m gpu 2>&1 | tail -15 | tee "$OUT/pytest_v2.log"
echo "== K2 v1 vs v2, 10M x 1024"
timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -2 | tee "$OUT/k2_v1.log"
STB_BATCH_V2=1 timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -2 | tee "$OUT/k2_v2.log"
echo "== K2 v2 launch list (per-kernel times)"
STB_BATCH_V2=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$OUT/k2_v2_launches.csv" \
  python scripts/batch_probe.py 10000000 1024 1 > "$OUT/k2_v2_ncu.log" 2>&1
grep -E "stb_batch|stb_shadow" "$OUT/k2_v2_launches.csv" | cut -d, -f5,12- | cut -c1-140 | tail -16
echo "== K5 v1 vs v2 (ivfpq_probe)"
timeout 300 python scripts/ivfpq_probe.py 2>&1 | tail -6 | tee "$OUT/k5_v1.log"
STB_IVFPQ_V2=1 timeout 300 python scripts/ivfpq_probe.py 2>&1 | tail -6 | tee "$OUT/k5_v2.log"
echo "== K1 e2e with and without direct host output"
timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', d['value'], d['e2e']['value'])" | tee "$OUT/e2e_default.log"
STB_DIRECT_OUT=1 timeout 300 python bench.py --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('direct ', d['value'], d['e2e']['value'])" | tee "$OUT/e2e_direct.log"
echo "== K1 half-width shadow scan (value through search_topk_dev needs the shadow: bench builds it for K2)"
timeout 300 python scripts/tail_probe.py 10000000 32 2>&1 | tail -3 | tee "$OUT/k1_f32.log"
STB_SCAN_SHADOW=1 STB_PROBE_PREPARE=1 timeout 300 python scripts/tail_probe.py 10000000 32 2>&1 | tail -3 | tee "$OUT/k1_shadow.log"
echo "== store query over 20k ranges: binary search per row vs range walk"
timeout 300 python scripts/modes_probe.py 10000000 2>&1 | tail -12 | tee "$OUT/modes_default.log"
STB_RANGES_WALK=1 timeout 300 python scripts/modes_probe.py 10000000 2>&1 | tail -12 | tee "$OUT/modes_walk.log"
echo "== fp16 shadow variant: batch tests + timing"
mkdir -p semtools_b200/lib/variants
STB_NVCC_EXTRA=-DSTB_SHADOW_F16=1 STB_LIB_OUT=semtools_b200/lib/variants/libstb_f16.so bash scripts/build_lib.sh
STB_LIB_PATH=$PWD/semtools_b200/lib/variants/libstb_f16.so STB_TEST_V2=1 timeout 600 python -m pytest tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -5 | tee "$OUT/pytest_f16.log"
STB_LIB_PATH=$PWD/semtools_b200/lib/variants/libstb_f16.so STB_BATCH_V2=1 timeout 300 python scripts/batch_probe.py 10000000 1024 5 2>&1 | tail -2 | tee "$OUT/k2_v2_f16.log"
echo "== full bench (side sections batch1024_v2 / ivfpq_v2 included)"
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 10 2>&1 | tail -2 | tee "$OUT/bench.log"
ls -la "$OUT"
