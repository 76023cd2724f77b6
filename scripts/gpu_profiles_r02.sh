#!/usr/bin/env bash
# ncu captures for profiles/ (round 2): K1 per tier, K2 pipeline v2 (sample / thresh / emit GEMM / finish2),
# K3, K5 fused search, and the launch list of the benchmarked command.  One GPU, never multi-rank.
set -uo pipefail
OUT=gpurun_out/${1:-r02p}; mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
for tier in q8 h16 f32; do
  timeout 400 $NCU -k regex:stb_scan_topk -s 3 -c 1 -o $OUT/k1_$tier python scripts/k1_ncu_probe.py $tier 10000000 6 > $OUT/k1_$tier.log 2>&1
done
timeout 600 $NCU -k regex:stb_batch -c 4 -o $OUT/k2_v2 python scripts/batch_probe.py 10000000 1024 1 > $OUT/k2_v2.log 2>&1
timeout 400 $NCU -k regex:stb_embed_kernel -s 2 -c 1 -o $OUT/k3_embed python scripts/embed_probe.py > $OUT/k3.log 2>&1
timeout 600 $NCU -k 'regex:ivf_(coarse_probe|adc_finish)' -s 8 -c 2 -o $OUT/k5_v2 python scripts/ivfpq_probe.py 4000000 4096 > $OUT/k5.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/bench_launches.csv python bench.py --steps 20 --warmup 3 --no-side > $OUT/bench_under_ncu.log 2>&1
ls -la $OUT
