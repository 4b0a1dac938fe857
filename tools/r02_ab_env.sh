#!/bin/bash
# interleaved A/B of one environment switch on the headline bench: tools/r02_ab_env.sh VAR [repeats]
cd "$(dirname "$0")/.."
VAR=$1; N=${2:-3}
for i in $(seq $N); do for z in 1 0; do env $VAR=$z timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$VAR=$z', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1))"; done; done
