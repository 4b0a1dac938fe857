#!/usr/bin/env python
"""1-block ViT-L-wide model: 6 forward passes into fresh output buffers, rows that differ from the first / from the last pass."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200 as d
from dinov2_b200 import synth
B, H, W, first = int(os.environ.get("DET_B", "8")), 518, 518, 64
L = int(os.environ.get("DET_L", "1"))
NP = (H // 14) * (W // 14)
x = torch.from_numpy(synth.lcg_batch(first, B, H, W)).cuda()
cfg = synth.ModelConfig(f"w1024_L{L}", 1024, L, 16)
path = f"/tmp/dino_bench/w1024_L{L}.gguf"
os.makedirs("/tmp/dino_bench", exist_ok=True)
if not os.path.exists(path):
    synth.write_synth_gguf(path, cfg, seed=0)
with d.Engine(path) as e:
    outs = []
    for i in range(6):
        cls = torch.empty(B, 1024, device="cuda"); patch = torch.empty(B, NP, 1024, device="cuda")
        e.forward_device(x.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W, False, cls_ptr=cls.data_ptr(), patch_ptr=patch.data_ptr())
        torch.cuda.synchronize()
        outs.append(patch)
    ref = outs[-1]
    for i in range(5):
        diff = (outs[i] != ref).any(dim=-1)          # [B, NP]
        if bool(diff.any()):
            desc = []
            for im in diff.any(dim=-1).nonzero().flatten().tolist():
                t = diff[im].nonzero().flatten()
                desc.append(f"image {im}: {len(t)} tokens {int(t[0])}..{int(t[-1])}")
            print(f"  run {i} vs run 5: " + "; ".join(desc) + f"; max |diff| {float((outs[i]-ref).abs().max()):.2e}", flush=True)
        else:
            print(f"  run {i} vs run 5: identical", flush=True)
