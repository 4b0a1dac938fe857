#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=dinov2.cpp_b200/lib
DINO_B200_ATTN=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -f -o gpurun_out/r02_attn_v8 python tools/attn_once.py $L/libdinov2_b200.so 2 > gpurun_out/r02_run2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -f -o gpurun_out/r02_attn_v9 python tools/attn_once.py $L/libdinov2_b200.so 2 >> gpurun_out/r02_run2.log 2>&1
tail -5 gpurun_out/r02_run2.log; ls -la gpurun_out/*.ncu-rep
