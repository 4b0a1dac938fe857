#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -x -q -k "not vitg and not vitb14_b64" 2>&1 | tail -3
python tools/latency.py vitl14 vits14 2>&1 | grep -v dino_model | grep "batch 1\|realtime"
DINO_B200_PDL=0 python tools/latency.py vitl14 2>&1 | grep "batch 1" | sed 's/$/ [PDL off]/'
for z in 1 0; do DINO_B200_PDL=$z timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('pdl=$z', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'])"; done
