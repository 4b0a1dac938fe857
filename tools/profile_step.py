#!/usr/bin/env python
"""Minimal driver for ncu: N device-resident forward steps of one model/batch, nothing else.
usage: python tools/profile_step.py [model] [batch] [iters] [classify 0|1]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dinov2_b200 as d  # noqa: E402
from dinov2_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "vitl14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
classify = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
cfg = synth.CONFIGS[name]
os.makedirs("/tmp/dino_bench", exist_ok=True)
path = f"/tmp/dino_bench/{name}_f16_seed0.gguf"
if not os.path.exists(path):
    synth.write_synth_gguf(path, cfg, seed=0)
eng = d.Engine(path)
imgs = torch.from_numpy(synth.lcg_batch(0, 2, 518, 518)).cuda().repeat((B + 1) // 2, 1, 1, 1)[:B].contiguous()
cls = torch.empty(B, cfg.hidden_size, device="cuda")
probs = torch.empty(B, cfg.num_classes, device="cuda")
stream = torch.cuda.Stream()
eng.reserve(B, 518, 518)
for _ in range(iters):
    eng.forward_device(imgs.data_ptr(), d.LAYOUT_BGR_HWC, B, 518, 518, classify, cls_ptr=cls.data_ptr(),
                       probs_ptr=probs.data_ptr() if classify else 0, stream=stream.cuda_stream)
torch.cuda.synchronize()
print("launches", eng.kernel_launches, "finite", bool(torch.isfinite(cls).all()))
