#!/usr/bin/env bash
# Builds libcsfm.so (CUDA kernels + C ABI) in-tree for sm_100a. No GPU needed: nvcc cross-compiles.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="${CSFM_OUT:-$HERE/libcsfm.so}"  # CSFM_OUT / CSFM_NVCC_EXTRA: experiment builds beside the product library
SRCS=("$HERE"/csrc/csfm_api.cu "$HERE"/csrc/csfm_build.cu "$HERE"/csrc/csfm_query.cu "$HERE"/csrc/csfm_query2.cu "$HERE"/csrc/csfm_query3.cu "$HERE"/csrc/csfm_sa.cu)
newest=$(ls -t "$HERE"/csrc/* "$HERE"/../include/csfm.h "$HERE"/build.sh | head -1)
if [[ -f "$OUT" && "$OUT" -nt "$newest" && "${1:-}" != "-f" ]]; then
  exit 0
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
  -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Xptxas -v \
  ${CSFM_NVCC_EXTRA:-} --shared -o "$OUT" "${SRCS[@]}" -ccbin /usr/bin/g++ 2> "$HERE/build.log" || { cat "$HERE/build.log" >&2; exit 1; }
grep -E "error|warning: v|registers|spill" "$HERE/build.log" | grep -v "^$" | head -60 || true
