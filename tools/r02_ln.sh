#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_layernorm or gemm" 2>&1 | tail -3
timeout 300 python tools/gemm_bench.py 2>&1 | grep "gemm "
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not vitg and not vitb14_b64 and not outlier" 2>&1 | tail -2
bash tools/r02_ab_env.sh DINO_B200_FUSE_LN 2
