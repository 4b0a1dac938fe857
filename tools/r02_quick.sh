#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "not vitg and not vitb14_b64 and not outlier and not vitl14_b64" 2>&1 | tail -2
python tools/attn_bench.py 2>&1 | grep -E "attention B=|rescale|big image"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms', d['clocks']['sm_mhz'], 'e2e', round(d['e2e']['value'],1), {k:round(v,2) for k,v in r['ms_profiled_step'].items()})"
