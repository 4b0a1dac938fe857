#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=dinov2.cpp_b200/lib
{
for v in poly4 poly3 poly2; do timeout 300 python tools/attn_bench.py $L/libdinov2_b200_$v.so 2>&1 | tail -3; done
timeout 300 python tools/attn_bench.py $L/libdinov2_b200.so 2>&1 | tail -1
} > gpurun_out/r02_run4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -f -o gpurun_out/r02_attn_v10b python tools/attn_once.py $L/libdinov2_b200.so 2 >> gpurun_out/r02_run4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -f -o gpurun_out/r02_attn_v10b_poly4 python tools/attn_once.py $L/libdinov2_b200_poly4.so 2 >> gpurun_out/r02_run4.log 2>&1
grep -v "^==PROF\|^check B" gpurun_out/r02_run4.log | tail
