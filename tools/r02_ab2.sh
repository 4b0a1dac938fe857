#!/bin/bash
# time-only A/B of library variants: tools/r02_ab2.sh name1 name2 ...
cd "$(dirname "$0")/.."
L=dinov2.cpp_b200/lib
for v in "$@"; do
  if [ "$v" = base ]; then f=$L/libdinov2_b200.so; else f=$L/libdinov2_b200_$v.so; fi
  timeout 300 python tools/attn_bench.py $f 2>&1 | grep -E "attention B=|rescale|big image 63" | sed "s/^/[$v] /"
done
