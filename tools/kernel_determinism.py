#!/usr/bin/env python
"""Run-to-run bit reproducibility of the stand-alone kernels at the ViT-L batch-64 shapes (same inputs, N launches)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dinov2_b200 import engine as E
torch.manual_seed(0)
B, N, D = 64, 1370, 1024
M = B * N
R = int(sys.argv[1]) if len(sys.argv) > 1 else 30
qkv = torch.randn(M, 3 * D, device="cuda").half()
outs = []
ref = None
bad_runs = 0
for i in range(R):
    out = torch.zeros(M, D, device="cuda", dtype=torch.half)
    E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D)
    torch.cuda.synchronize()
    if ref is None:
        ref = out
    else:
        diff = (out != ref).any(dim=-1)
        if bool(diff.any()):
            rows = diff.nonzero().flatten()
            cols = (out[rows[0]] != ref[rows[0]]).nonzero().flatten()
            bad_runs += 1
            print(f"attention run {i}: {int(diff.sum())} rows differ; first row {int(rows[0])} (image {int(rows[0]) // N}, token {int(rows[0]) % N}), last row {int(rows[-1])}; "
                  f"in first row {len(cols)} columns, first {int(cols[0])} (head {int(cols[0]) // 64}); max |diff| {float((out.float()-ref.float()).abs().max()):.3e}", flush=True)
print(f"attention: {bad_runs} of {R - 1} repeat runs differ from the first", flush=True)
for name, epi, Nn, K in (("qkv", E.EPI_BIAS_F16, 3 * D, D), ("proj", E.EPI_RESID_F32, D, D), ("fc1", E.EPI_GELU_F16, 4 * D, D), ("fc2", E.EPI_RESID_F32, D, 4 * D)):
    A = (torch.randn(M, K, device="cuda") * 0.5).half(); Wt = (torch.randn(Nn, K, device="cuda") * 0.05).half()
    bias = torch.randn(Nn, device="cuda") * 0.1; ls = torch.rand(Nn, device="cuda") + 0.3
    f32 = epi == E.EPI_RESID_F32
    X0 = torch.randn(M, Nn, device="cuda") if f32 else None
    ref = None; bad = 0
    for i in range(max(6, R // 3)):
        out = X0.clone() if f32 else torch.zeros(M, Nn, device="cuda", dtype=torch.half)
        E.kernel_gemm(epi, A.data_ptr(), K, Wt.data_ptr(), K, M, Nn, K, bias.data_ptr(), ls.data_ptr() if f32 else 0, out.data_ptr(), Nn)
        torch.cuda.synchronize()
        if ref is None: ref = out
        elif not torch.equal(out, ref):
            bad += 1
            diff = (out != ref).any(dim=-1).nonzero().flatten()
            print(f"gemm {name} run {i}: {len(diff)} rows differ, first {int(diff[0])}, max |diff| {float((out.float()-ref.float()).abs().max()):.3e}", flush=True)
    print(f"gemm {name}: {bad} repeat runs differ", flush=True)
X = torch.randn(M, D, device="cuda"); g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
ref = None; bad = 0
for i in range(6):
    o = torch.empty(M, D, device="cuda", dtype=torch.half)
    E.kernel_layernorm(X.data_ptr(), g.data_ptr(), b.data_ptr(), o.data_ptr(), M, D, 1e-6, True)
    torch.cuda.synchronize()
    if ref is None: ref = o
    elif not torch.equal(o, ref): bad += 1
print(f"layernorm: {bad} repeat runs differ", flush=True)
