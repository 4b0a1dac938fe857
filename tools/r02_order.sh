#!/bin/bash
cd "$(dirname "$0")/.."
for a in "--steps 10 --warmup 3" "--steps 10 --warmup 40" "--steps 40 --warmup 3" "--steps 10 --warmup 3"; do
timeout 600 python bench.py $a --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$a', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1), 'sync', round(d['e2e']['synchronous_value'],1))"; done
