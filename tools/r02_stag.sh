#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/dinov2.cpp_b200/lib
for v in "$@"; do timeout 120 python tools/attn_bench.py $L/libdinov2_b200$v.so 2>&1 | tail -1; done
