#!/usr/bin/env python
"""Sustained (seconds-long, power-capped) rate of our GEMM kernels against cuBLAS (torch.matmul) on the ViT-L shapes, with the
SM clock and board power sampled while each loop runs.  Answers: is the in-step GEMM rate limited by the kernel or by the 1 kW cap?
usage: python tools/gemm_power.py [seconds per loop]"""
import os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200
from dinov2_b200 import engine as E

SEC = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
M, D = 64 * 1370, 1024


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.proc = [], None

    def run(self):
        self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0", "-lms", "100"],
                                     stdout=subprocess.PIPE, text=True)
        for line in self.proc.stdout:
            try:
                c, p = [float(x) for x in line.split(",")]
                self.rows.append((c, p))
            except ValueError:
                pass

    def stop(self):
        self.proc.terminate()
        self.join(timeout=2)
        r = self.rows[len(self.rows) // 3:]          # drop the ramp-up third
        if not r:
            return 0.0, 0.0
        return sum(x[0] for x in r) / len(r), sum(x[1] for x in r) / len(r)


def sustained(fn, flops):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler(); s.start()
    time.sleep(0.3)
    n, t0 = 0, time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.perf_counter() - t0 < SEC:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    clk, pw = s.stop()
    return ms, flops / ms / 1e9, clk, pw


torch.manual_seed(0)
for name, epi, N, K in [("qkv", E.EPI_BIAS_F16, 3 * D, D), ("proj", E.EPI_RESID_F32, D, D), ("fc1", E.EPI_GELU_F16, 4 * D, D), ("fc2", E.EPI_RESID_F32, D, 4 * D)]:
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    W = (torch.randn(N, K, device="cuda") * 0.05).half()
    bias = torch.randn(N, device="cuda") * 0.1
    ls = torch.rand(N, device="cuda") + 0.3
    f32 = epi == E.EPI_RESID_F32
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if f32 else torch.half)
    out16 = torch.empty(M, N, device="cuda", dtype=torch.half)
    fl = 2.0 * M * N * K
    ms, tf, clk, pw = sustained(lambda: E.kernel_gemm(epi, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr() if f32 else 0, out.data_ptr(), N), fl)
    print(f"ours   {name:5s} N={N:5d} K={K:5d}: {ms*1000:7.1f} us {tf:7.1f} TFLOP/s  sm {clk:6.0f} MHz  {pw:6.0f} W  -> {tf/clk*1000:6.1f} TFLOP/s per GHz", flush=True)
    ms, tf, clk, pw = sustained(lambda: torch.matmul(A, W.t(), out=out16), fl)
    print(f"cublas {name:5s} N={N:5d} K={K:5d}: {ms*1000:7.1f} us {tf:7.1f} TFLOP/s  sm {clk:6.0f} MHz  {pw:6.0f} W  -> {tf/clk*1000:6.1f} TFLOP/s per GHz (plain fp16 GEMM, no epilogue)", flush=True)
