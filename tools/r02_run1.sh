#!/bin/bash
# round-2 GPU call 1: correctness of the refactor (graphs, v9 attention), attention A/B, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=dinov2.cpp_b200/lib
{
echo "== attention A/B"
DINO_B200_ATTN=8 timeout 300 python tools/attn_bench.py $L/libdinov2_b200.so 2>&1 | tail -9
timeout 300 python tools/attn_bench.py $L/libdinov2_b200.so 2>&1 | tail -9
timeout 300 python tools/attn_bench.py $L/libdinov2_b200_poly8.so 2>&1 | tail -3
timeout 300 python tools/attn_bench.py $L/libdinov2_b200_poly4.so 2>&1 | tail -3
echo "== trace v9"
timeout 300 python tools/attn_trace.py $L/libdinov2_b200_trace9.so 200 > gpurun_out/r02_attn_v9_cycle_trace.txt 2>&1; tail -3 gpurun_out/r02_attn_v9_cycle_trace.txt
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3
DINO_B200_GRAPH=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/r02_run1.log 2>&1
tail -60 gpurun_out/r02_run1.log
