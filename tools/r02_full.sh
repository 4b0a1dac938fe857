#!/bin/bash
# full check on one B200: GPU test-suite, smoke, default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py 2>gpurun_out/bench_stderr.log | tail -1 > gpurun_out/bench_line.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_line.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks','per_rank_ms_per_step')}))
print('e2e', d['e2e']['value'], 'sync', d['e2e']['synchronous_value'])
r=d['roofline']; print('gemm', r['achieved'], r['frac'], 'attn', r['attention_tflops'], 'step', r['whole_step_tflops'], r['whole_step_frac_of_burst'], r['ms_profiled_step'])
for c in d.get('configs') or []: print(c)
print(d.get('cpu_baseline'))
PY
} > gpurun_out/r02_full.log 2>&1
tail -40 gpurun_out/r02_full.log | cut -c1-700
