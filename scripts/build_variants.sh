#!/usr/bin/env bash
# Builds scan-kernel tuning variants into semtools_b200/lib/variants/ (git-ignored).
# Each spec: U:LD:THREADS:MINB
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p semtools_b200/lib/variants
SPECS=${*:-"2:0:256:2 2:1:256:2 2:2:256:2 1:0:256:2 1:0:256:3 1:0:256:4 4:0:256:1 2:0:512:1 2:0:128:4 4:0:128:2 1:0:512:2 2:0:256:1"}
for spec in $SPECS; do
  IFS=: read -r U LD T B <<< "$spec"
  out=semtools_b200/lib/variants/libstb_U${U}_LD${LD}_T${T}_B${B}.so
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
      -ccbin /usr/bin/g++ -Xcompiler -fPIC -shared -DSTB_SCAN_U=$U -DSTB_SCAN_LD=$LD \
      -DSTB_SCAN_THREADS=$T -DSTB_SCAN_MINB=$B -Xptxas -v -o $out \
      semtools_b200/csrc/api.cu semtools_b200/csrc/scan_topk.cu semtools_b200/csrc/hits_merge.cu \
      semtools_b200/csrc/embed_pool.cu 2>&1 | grep -A2 "scan_topk_kernelILi1ELi${U}ELb0" | grep -E "spill|Used" | tr '\n' ' ' | sed "s|^|$spec: |"; echo ) &
done
wait
ls semtools_b200/lib/variants | wc -l
