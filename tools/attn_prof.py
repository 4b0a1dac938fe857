#!/usr/bin/env python
"""Barrier-wait cycle counters of the attention kernel's CTA 0 (library built with -DAT10_PROF).
usage: python tools/attn_prof.py dinov2.cpp_b200/lib/libdinov2_b200_prof10.so"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200
from dinov2_b200 import engine as E
E.LIB_PATH = os.path.abspath(sys.argv[1])
os.environ["DINO_B200_TRACE_PTR"] = "/tmp/trace_ptr.txt"
B, N, D = 64, 1370, 1024
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
out = torch.zeros(B * N, D, device="cuda", dtype=torch.half)
for _ in range(3):
    E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D)
torch.cuda.synchronize()
ptr = int(open("/tmp/trace_ptr.txt").read())
buf = (ctypes.c_uint64 * 16)()
ctypes.CDLL("libcudart.so.12").cudaMemcpy(buf, ctypes.c_void_p(ptr), ctypes.sizeof(buf), 2)
names = ["producer kv_empty", "producer q_empty", "MMA0 kv_full", "MMA0 s_free", "MMA0 p_full", "MMA0 q_full", "WG0 s_full",
         "WG0 o_full (epilogue)", "WG0 total", "MMA0 total", "WG0 epilogue total"]
tiles = 457
for i, n in enumerate(names):
    print(f"{n:24s} {buf[i]:10d} cycles  {buf[i] / tiles:8.1f} / tile   {100.0 * buf[i] / max(1, buf[8]):5.1f} % of WG0 total")
