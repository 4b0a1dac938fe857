#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
nvidia-smi -L
echo "== two-device tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
echo "== bench 2 GPUs"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench2_stderr.log | tail -1 > gpurun_out/bench2_line.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2_line.json'))
print(json.dumps({k:d[k] for k in ("value","ms_per_step","n_gpus","gpu_launches","per_rank_ms_per_step","scale_features")}, indent=1))
print('e2e', d['e2e']['value'])
PY
} > gpurun_out/r02_2gpu.log 2>&1
tail -40 gpurun_out/r02_2gpu.log; tail -5 gpurun_out/bench2_stderr.log
