#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q -k "not vitg and not vitb14_b64 and not outlier" 2>&1 | tail -3
for z in 1 0 1 0; do
DINO_B200_ZIGZAG=$z timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('zigzag=$z', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms  profiled:', {k:round(v,2) for k,v in r['ms_profiled_step'].items()}, d['clocks']['sm_mhz'])"
done
