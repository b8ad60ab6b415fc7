#!/bin/bash
# Developer tool: build libtvs.so variants with extra -D flags into build/variants/libtvs_<name>.so (git-ignored; travels with gpurun).
# usage: tools/build_variant.sh name [-DFLAG ...]
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
NAME="$1"; shift
mkdir -p "$ROOT/build/variants"
cd "$ROOT/traversome_b200/csrc"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -shared \
    ${TVS_PTXAS_V:+-Xptxas -v} "$@" -o "$ROOT/build/variants/libtvs_$NAME.so" tvs_api.cu tvs_subpaths.cu tvs_pool.cu tvs_gaf.cpp \
    -lcudart -lpthread -lcublas -Xlinker -rpath=/usr/local/cuda/lib64
echo "built build/variants/libtvs_$NAME.so"
