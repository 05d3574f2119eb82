#!/usr/bin/env bash
# A/B builds of the CUDA library with extra -D flags: scripts/build_variant.sh NAME [-DFLAG=..]...  ->
# little-physics-engine_b200/variants/liblpe_bh_NAME.so (select with LPE_BH_LIB=...; *.so is git-ignored, travels with gpurun)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")/../little-physics-engine_b200" && pwd)"
NAME="$1"; shift
mkdir -p "$HERE/variants"
/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3,-Wall -Xptxas -v --shared "$@" \
  -o "$HERE/variants/liblpe_bh_$NAME.so" "$HERE/csrc/lpe_bh.cu" "$HERE/csrc/workloads.cpp" -lcudart 2> "$HERE/variants/build_$NAME.log"
echo "built variants/liblpe_bh_$NAME.so"
