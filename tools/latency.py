#!/usr/bin/env python
"""Small-batch numbers: batch-1 latency of a device-resident forward (what inference.cpp does per image) with and without
CUDA-graph replay, and frames/s of the pipelined raw-frame path (realtime.cpp's loop: 854x480 BGR frame -> preprocess -> forward
-> PCA colours) — usage: python tools/latency.py [vitl14] [vits14]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import dinov2_b200 as d
from dinov2_b200 import synth

def gguf(name):
    p = f"/tmp/dino_bench/{name}_f16_seed0.gguf"
    os.makedirs("/tmp/dino_bench", exist_ok=True)
    if not os.path.exists(p):
        synth.write_synth_gguf(p, synth.CONFIGS[name], seed=0)
    return p

def batch1(name):
    cfg = synth.CONFIGS[name]
    with d.Engine(gguf(name)) as e:
        x = torch.from_numpy(synth.lcg_batch(0, 1, 518, 518)).cuda()
        cls = torch.empty(1, cfg.hidden_size, device="cuda"); probs = torch.empty(1, cfg.num_classes, device="cuda")
        st = torch.cuda.Stream()
        def run(n):
            for _ in range(n):
                e.forward_device(x.data_ptr(), d.LAYOUT_BGR_HWC, 1, 518, 518, True, cls_ptr=cls.data_ptr(), probs_ptr=probs.data_ptr(), stream=st.cuda_stream)
        with torch.cuda.stream(st):
            run(10); st.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); run(50); b.record(st); st.synchronize()
        l0 = e.kernel_launches
        run(1); torch.cuda.synchronize()
        print(f"{name} batch 1 classify, 518x518, device-resident: {a.elapsed_time(b) / 50:.3f} ms per image "
              f"({e.kernel_launches - l0} kernels, graphs {'off' if os.environ.get('DINO_B200_GRAPH') == '0' else 'on'})", flush=True)

def realtime(name, B=1, n=60):
    with d.Engine(gguf(name)) as e:
        rng = np.random.default_rng(0)
        frames = [torch.from_numpy(rng.integers(0, 256, (B, 480, 854, 3), dtype=np.uint8)).pin_memory().numpy() for _ in range(2)]
        oh, ow = e.preprocess_size(480, 854, False)
        NP = e.n_patches(oh, ow)
        outs = [{"pca_rgb": torch.empty(B, NP, 3, dtype=torch.uint8).pin_memory().numpy()} for _ in range(2)]
        def loop(k):
            e.submit_u8(frames[0], outs[0])
            for i in range(1, k):
                e.submit_u8(frames[i & 1], outs[i & 1]); e.wait()
            e.wait()
        loop(6)
        t0 = time.perf_counter(); loop(n); dt = time.perf_counter() - t0
        # synchronous reference point: forward_u8 + pca_rgb one after the other, patch tokens through the host
        t1 = time.perf_counter()
        for i in range(10):
            r = e.forward_u8(frames[i & 1], want_cls=False); e.pca_rgb(r["patch_tokens"])
        ds = (time.perf_counter() - t1) / 10
        print(f"{name} realtime loop, 854x480 frame -> {ow}x{oh} ({NP} patches), batch {B}: pipelined submit_u8 {n * B / dt:.1f} frames/s "
              f"({dt / n * 1e3:.2f} ms per submission); synchronous forward_u8 + pca_rgb {B / ds:.1f} frames/s", flush=True)

if __name__ == "__main__":
    names = sys.argv[1:] or ["vitl14", "vits14"]
    for nm in names:
        batch1(nm)
    realtime("vits14")
    realtime("vitb14")
