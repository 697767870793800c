#!/usr/bin/env bash
# Builds libintopt_b200.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -Xcompiler -Wall -shared ${INTOPT_NVCC_EXTRA:-} \
    -o "$OUT/libintopt_b200.so" "$HERE/intopt_capi.cu"
echo "built $OUT/libintopt_b200.so"
