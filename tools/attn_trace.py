#!/usr/bin/env python
"""Cycle trace of the attention kernel (CTA 0).  Needs a trace build of the library next to the normal one:
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared -DAT8_TRACE \\
         dinov2.cpp_b200/csrc/engine.cu -o dinov2.cpp_b200/lib/libdinov2_b200_trace.so
(-DAT3_TRACE / -DAT5_TRACE / -DAT7_TRACE with DINO_B200_ATTN=3/5/7 for the older generations).  Output: profiles/r01_attn_v*_cycle_trace.txt."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200
from dinov2_b200 import engine as E
E.LIB_PATH = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].endswith(".so") else os.path.join(ROOT, "dinov2.cpp_b200", "lib", "libdinov2_b200_trace.so")
N_EVENTS = int(sys.argv[2]) if len(sys.argv) > 2 else 260
os.environ["DINO_B200_TRACE_PTR"] = "/tmp/trace_ptr.txt"
B, N, D = 64, 1370, 1024
qkv = torch.randn(B * N, 3 * D, device="cuda").half()
out = torch.zeros(B * N, D, device="cuda", dtype=torch.half)
for _ in range(2):
    E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D)
torch.cuda.synchronize()
ptr = int(open("/tmp/trace_ptr.txt").read())
buf = (ctypes.c_uint64 * (3 * 512 * 2))()
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpy(buf, ctypes.c_void_p(ptr), ctypes.sizeof(buf), 2)
names = {18: "epilogue start", 19: "epilogue done", 5: "kv_full ok", 6: "s_free ok", 7: "p_full ok", 8: "iter top", 9: "mma issued", 1: "S0 issued", 2: "S1 issued", 3: "PV0 issued", 4: "PV1 issued", 10: "wait S", 11: "got S", 12: "S in regs", 13: "o_full ok", 14: "exp start", 15: "exp done", 16: "P stored", 17: "p_full arrived"}
ev = []
for role in range(3):
    for i in range(512):
        tag, clk = buf[(role * 512 + i) * 2], buf[(role * 512 + i) * 2 + 1]
        if clk == 0:
            break
        ev.append((clk, role, tag >> 32, tag & 0xffffffff))
ev.sort()
t0 = ev[0][0]
rolen = ["MMA", "WG0", "WG1"]
last = {}
for clk, role, eid, idx in [e for e in ev if True][:N_EVENTS]:
    d = clk - last.get(role, clk)
    last[role] = clk
    print(f"{clk - t0:8d}  (+{d:5d})  {rolen[role]:4s} {str(names.get(eid, eid)):16s} #{idx}")
