#!/bin/bash
cd "$(dirname "$0")/.."
run() { env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'])"; }
L=$PWD/dinov2.cpp_b200/lib
for i in 1 2; do
echo -n "pdl off:        "; run DINO_B200_PDL=0
echo -n "pdl early all:  "; run DINO_B200_PDL_ROWS=100000000
echo -n "pdl late all:   "; run DINO_B200_PDL_ROWS=100000000 DINO_B200_LIB=$L/libdinov2_b200_pdllate.so
done
DINO_B200_LIB=$L/libdinov2_b200_pdllate.so python tools/latency.py vitl14 2>&1 | grep "batch 1" | sed 's/$/ [late]/'
python tools/latency.py vitl14 2>&1 | grep "batch 1" | sed 's/$/ [early]/'
