#!/usr/bin/env bash
# Builds liblpe_bh.so (CUDA kernels + C ABI + host workloads) in-tree for sm_100a.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/liblpe_bh.so"
"$NVCC" -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC,-O3,-Wall -Xptxas -v --shared \
  -o "$OUT" "$HERE/csrc/lpe_bh.cu" "$HERE/csrc/workloads.cpp" -lcudart 2> "$HERE/build.log" || { cat "$HERE/build.log"; exit 1; }
echo "built $OUT"
# the same library with the kernels' own bounds checks compiled in (compute-sanitizer is closed on this pool): used once by
# tests/test_checked_build_gpu.py, never by the product path
if [ "${LPE_BUILD_CHECKED:-1}" = "1" ]; then
  "$NVCC" -std=c++17 -O3 -lineinfo -DLPE_CHECKED -gencode arch=compute_100a,code=sm_100a \
    -Xcompiler -fPIC,-O3,-Wall --shared \
    -o "$HERE/liblpe_bh_checked.so" "$HERE/csrc/lpe_bh.cu" "$HERE/csrc/workloads.cpp" -lcudart 2> "$HERE/build_checked.log" || { cat "$HERE/build_checked.log"; exit 1; }
  echo "built $HERE/liblpe_bh_checked.so"
fi
# the synthetic workload generators alone (no CUDA): what bench.py --impl reference loads instead of the product library
"${CXX:-g++}" -std=c++17 -O3 -fPIC -shared -Wall -o "$HERE/libworkloads.so" "$HERE/csrc/workloads.cpp"
echo "built $HERE/libworkloads.so"
