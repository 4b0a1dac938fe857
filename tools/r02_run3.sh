#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=dinov2.cpp_b200/lib
timeout 300 python tools/attn_trace.py $L/libdinov2_b200_trace10.so 500 > gpurun_out/r02_attn_v10_cycle_trace.txt 2>&1; tail -2 gpurun_out/r02_attn_v10_cycle_trace.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -f -o gpurun_out/r02_attn_v10 python tools/attn_once.py $L/libdinov2_b200.so 2 > gpurun_out/r02_run3.log 2>&1
tail -3 gpurun_out/r02_run3.log
