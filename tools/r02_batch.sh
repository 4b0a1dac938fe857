#!/bin/bash
cd "$(dirname "$0")/.."
for b in 64 4 8 16 32 64; do
s=$((1280 / b)); [ $s -lt 20 ] && s=20
timeout 600 python bench.py --batch $b --steps $s --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('batch $b', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'MHz', d['clocks']['reasons'], 'e2e', round(d['e2e']['value'],1))"; done
