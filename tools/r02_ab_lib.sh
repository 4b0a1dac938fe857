#!/bin/bash
# interleaved A/B of library variants on the headline bench: tools/r02_ab_lib.sh repeats name1 name2 ...
cd "$(dirname "$0")/.."
N=$1; shift
for i in $(seq $N); do for v in "$@"; do
  if [ "$v" = base ]; then f=dinov2.cpp_b200/lib/libdinov2_b200.so; else f=dinov2.cpp_b200/lib/libdinov2_b200_$v.so; fi
  DINO_B200_LIB=$PWD/$f timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'MHz  e2e', round(d['e2e']['value'],1), ' profiled attn', round(r['ms_profiled_step']['attn_ms'],2))"
done; done
