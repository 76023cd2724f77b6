#!/usr/bin/env bash
# Builds libsemtools_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
cd "$(dirname "$0")/.."
mkdir -p semtools_b200/lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -ccbin /usr/bin/g++ -Xcompiler -fPIC -shared ${STB_NVCC_EXTRA:-} \
  -o semtools_b200/lib/libsemtools_b200.so \
  semtools_b200/csrc/api.cu semtools_b200/csrc/scan_topk.cu \
  semtools_b200/csrc/hits_merge.cu semtools_b200/csrc/embed_pool.cu \
  semtools_b200/csrc/batch_scan.cu semtools_b200/csrc/ivfpq.cu "$@"
