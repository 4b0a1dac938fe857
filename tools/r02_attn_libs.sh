#!/bin/bash
# times the stand-alone attention kernel for a list of library variants: tools/r02_attn_libs.sh "" _split _pp1   (suffixes of dinov2.cpp_b200/lib/libdinov2_b200<suffix>.so, see tools/build_variant.sh)
cd "$(dirname "$0")/.."
L=$PWD/dinov2.cpp_b200/lib
for v in "$@"; do timeout 120 python tools/attn_bench.py $L/libdinov2_b200$v.so 2>&1 | tail -1; done
