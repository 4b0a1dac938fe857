#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/dinov2.cpp_b200/lib
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_layernorm" 2>&1 | tail -1
for v in "" _ln4 _ln2; do echo "== lib$v"; python tools/gemm_bench.py $L/libdinov2_b200$v.so 2>&1 | grep "gemm " | grep "proj\|fc2\|ln "; done
python tools/ln_prof.py $L/libdinov2_b200_lnprof.so
