#!/usr/bin/env python
"""GEMM kernel alone at the ViT-L bench shapes: time per call (CUDA events) and TFLOP/s.
usage: python tools/gemm_bench.py [lib.so]     env DINO_B200_GEMM_SMS / DINO_B200_GEMM_CG select experiments"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200
from dinov2_b200 import engine as E
if len(sys.argv) > 1:
    E.LIB_PATH = os.path.abspath(sys.argv[1])
torch.manual_seed(0)
# GEMM_D / GEMM_ROWS select another model width / row count (default: ViT-L, batch 64): e.g. GEMM_D=384 GEMM_ROWS=43968 for ViT-S b32
Dm = int(os.environ.get("GEMM_D", "1024"))
M = int(os.environ.get("GEMM_ROWS", str(64 * 1370)))
shapes = [("qkv", E.EPI_BIAS_F16, 3 * Dm, Dm), ("proj", E.EPI_RESID_F32, Dm, Dm), ("fc1", E.EPI_GELU_F16, 4 * Dm, Dm), ("fc2", E.EPI_RESID_F32, Dm, 4 * Dm)]
tag = f"sms={os.environ.get('DINO_B200_GEMM_SMS','all')} cg={os.environ.get('DINO_B200_GEMM_CG','2')} mc={os.environ.get('DINO_B200_GEMM_MC','0')} lib={os.path.basename(E.LIB_PATH)}"
for name, epi, N, K in shapes:
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda") * 0.1
    ls = torch.rand(N, device="cuda") + 0.3
    f32 = epi == E.EPI_RESID_F32
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.half)
    def run():
        E.kernel_gemm(epi, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr() if f32 else 0, out.data_ptr(), N)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if not f32:
        worst = 0.0
        for r0 in (0, 300, M // 2 - 77, M - 512):          # windows at the start, across tile borders, at the ragged end
            ref = (A[r0:r0 + 512].float() @ W.float().t() + bias)
            if epi == E.EPI_GELU_F16:
                v = ref.half().float()
                ref = 0.5 * v * (1 + torch.tanh(0.7978845608028654 * v * (1 + 0.044715 * v * v)))
            worst = max(worst, float((out[r0:r0 + 512].float() - ref).abs().max()))
        print(f"check {name}: max_abs over 4 row windows {worst:.3e}", flush=True)
    else:
        X0 = torch.randn(M, N, device="cuda")
        out.copy_(X0)
        run()
        torch.cuda.synchronize()
        worst = 0.0
        for r0 in (0, 300, M // 2 - 77, M - 512):
            ref = X0[r0:r0 + 512] + ls * (A[r0:r0 + 512].float() @ W.float().t() + bias)
            worst = max(worst, float((out[r0:r0 + 512] - ref).abs().max()))
        print(f"check {name}: max_abs over 4 row windows {worst:.3e}", flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record()
    for _ in range(iters):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / iters
    print(f"gemm {name:5s} M={M} N={N} K={K}: {ms*1000:7.1f} us  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s   [{tag}]", flush=True)
    if f32:
        # the same GEMM with the following LayerNorm fused into its epilogue, and the stand-alone LayerNorm it replaces
        gam = torch.randn(N, device="cuda"); bet = torch.randn(N, device="cuda")
        ln = torch.empty(M, N, device="cuda", dtype=torch.half)
        cnt = torch.zeros(2 * ((M + 127) // 128) + 2, device="cuda", dtype=torch.int32)
        def run2():
            E.kernel_gemm_resid_ln(A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr(), out.data_ptr(),
                                   gam.data_ptr(), bet.data_ptr(), 1e-6, ln.data_ptr(), cnt.data_ptr())
        def run3():
            E.kernel_layernorm(out.data_ptr(), gam.data_ptr(), bet.data_ptr(), ln.data_ptr(), M, N, 1e-6, True)
        for fn, nm in ((run2, name + "+ln"), (run3, "ln")):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            print(f"gemm {nm:7s} M={M} N={N} K={K}: {ms*1000:7.1f} us   [{tag}]", flush=True)
