#!/usr/bin/env bash
# One gpurun call: tests -> smoke -> bench -> ncu launch list -> ncu full capture.
# Usage (from the repo root on the GPU box): bash scripts/gpu_check.sh [tag]
set -uo pipefail
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > "$OUT/gpu.txt" 2>&1
nproc > "$OUT/nproc.txt"; lscpu | grep -E "Model name|^CPU\(s\)" >> "$OUT/nproc.txt"
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee "$OUT/smoke.log"
echo "== bench"; timeout 900 python bench.py --gpus 1 --steps 200 --warmup 10 2>&1 | tail -3 | tee "$OUT/bench.log"
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file "$OUT/launches.csv" python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/ncu_launches_bench.log" 2>&1
grep -E "stb_|ncclDev" "$OUT/launches.csv" | cut -d, -f5,12- | cut -c1-160 | tail -12
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stb_scan_topk -s 5 -c 2 \
  -o "$OUT/scan_topk" -f python bench.py --steps 10 --warmup 3 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
tail -3 "$OUT/ncu_full.log"
ls -la "$OUT"
