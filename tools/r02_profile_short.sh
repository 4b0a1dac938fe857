#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export DINO_B200_GRAPH=0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_launches_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 24 -c 1 -f -o gpurun_out/r02_prof_attn \
    python tools/profile_step.py vitl14 64 2 > gpurun_out/r02_prof_attn.log 2>&1
ls -la gpurun_out/r02_prof_attn.ncu-rep gpurun_out/r02_launches_bench.csv
