#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/dinov2.cpp_b200/lib
timeout 120 python tools/attn_bench.py $L/libdinov2_b200.so 2>&1 | tail -8
for v in _ip0 _pm2 _pm4 _pm0 _pp1; do timeout 120 python tools/attn_bench.py $L/libdinov2_b200$v.so 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k attention 2>&1 | tail -2
