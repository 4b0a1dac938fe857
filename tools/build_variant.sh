#!/bin/bash
# Build an A/B variant of the engine library next to the product one:
#   tools/build_variant.sh NAME [-DFLAG ...]   ->  dinov2.cpp_b200/lib/libdinov2_b200_NAME.so
# (tools/attn_bench.py, tools/gemm_bench.py and tools/attn_trace.py take the library path as an argument.)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
NAME="$1"; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared "$@" \
     "$ROOT/dinov2.cpp_b200/csrc/engine.cu" -o "$ROOT/dinov2.cpp_b200/lib/libdinov2_b200_${NAME}.so" 2>&1 | grep -E "error|spill|Used" | grep -v " 0 bytes spill" || true
echo "built libdinov2_b200_${NAME}.so"
