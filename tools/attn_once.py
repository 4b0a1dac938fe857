#!/usr/bin/env python
"""One warm-up + N calls of the attention kernel at the ViT-L bench shape (for ncu captures).
usage: python tools/attn_once.py [lib.so] [calls]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200
from dinov2_b200 import engine as E
if len(sys.argv) > 1 and sys.argv[1].endswith(".so"):
    E.LIB_PATH = os.path.abspath(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, N, D = 64, 1370, 1024
torch.manual_seed(0)
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
out = torch.zeros(B * N, D, device="cuda", dtype=torch.half)
for _ in range(n):
    E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D)
torch.cuda.synchronize()
print("done")
