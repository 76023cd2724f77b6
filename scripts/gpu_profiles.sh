#!/usr/bin/env bash
# ncu captures for profiles/: K1 (10M rows), K2 GEMM (10M x 1024), K3 (1M lines).
set -uo pipefail
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stb_batch_gemm -s 1 -c 1 -o "$OUT/k2_gemm" -f python scripts/batch_probe.py 10000000 1024 1 > "$OUT/k2_ncu.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stb_embed_kernel -s 2 -c 1 -o "$OUT/k3_embed" -f python scripts/embed_probe.py > "$OUT/k3_ncu.log" 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stb_scan_topk -s 4 -c 1 -o "$OUT/k1_scan" -f python scripts/tail_probe.py 10000000 4 > "$OUT/k1_ncu.log" 2>&1
ls -la "$OUT"
