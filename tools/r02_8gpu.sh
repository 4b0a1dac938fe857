#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench${N}_stderr.log | tail -1 > gpurun_out/bench${N}_line.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench${N}_line.json'))
print(json.dumps({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','per_rank_ms_per_step','clocks')}))
print('e2e', d['e2e']['value'])
print(json.dumps(d.get('scale_features'), indent=1))
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=1 on the same box:', round(d['value'],1), 'img/s', round(d['ms_per_step'],2), 'ms')"
