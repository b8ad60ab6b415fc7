#!/bin/bash
# Builds traversome_b200/libtvs.so for sm_100a (B200). Run from anywhere; nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libtvs.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared ${TVS_PTXAS_V:+-Xptxas -v} \
    -o "$OUT" "$HERE/tvs_api.cu" "$HERE/tvs_subpaths.cu" "$HERE/tvs_pool.cu" "$HERE/tvs_gaf.cpp" -lcudart -lpthread -lcublas -Xlinker -rpath=/usr/local/cuda/lib64
echo "built $OUT"
