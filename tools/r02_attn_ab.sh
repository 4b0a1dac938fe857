#!/bin/bash
# attention A/B on one B200: correctness + time of each library variant, a cycle trace, the GPU test-suite, a short bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=dinov2.cpp_b200/lib
{
echo "== attention A/B"
DINO_B200_ATTN=8 timeout 300 python tools/attn_bench.py $L/libdinov2_b200.so 2>&1 | tail -1
timeout 300 python tools/attn_bench.py $L/libdinov2_b200.so 2>&1 | tail -9
for v in "$@"; do timeout 300 python tools/attn_bench.py $L/libdinov2_b200_$v.so 2>&1 | tail -3; done
if [ -f $L/libdinov2_b200_trace10.so ]; then
echo "== trace v10"
timeout 300 python tools/attn_trace.py $L/libdinov2_b200_trace10.so 400 > gpurun_out/r02_attn_v10_cycle_trace.txt 2>&1; tail -2 gpurun_out/r02_attn_v10_cycle_trace.txt
fi
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/r02_attn_ab.log 2>&1
tail -70 gpurun_out/r02_attn_ab.log | cut -c1-600
