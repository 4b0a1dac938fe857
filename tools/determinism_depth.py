#!/usr/bin/env python
"""Where does run-to-run nondeterminism of the forward start?  ViT-L-wide models of 1..24 blocks, same images, 4 runs each."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200 as d
from dinov2_b200 import synth
B, H, W, first = int(os.environ.get("DET_B", "8")), 518, 518, 64
NP = (H // 14) * (W // 14)
x = torch.from_numpy(synth.lcg_batch(first, B, H, W)).cuda()
os.makedirs("/tmp/dino_bench", exist_ok=True)
for width, heads, depths in ((1024, 16, (1, 2, 3, 4, 8, 24)), (384, 6, (12,))):
    for L in depths:
        cfg = synth.ModelConfig(f"w{width}_L{L}", width, L, heads)
        path = f"/tmp/dino_bench/w{width}_L{L}.gguf"
        if not os.path.exists(path):
            synth.write_synth_gguf(path, cfg, seed=0)
        with d.Engine(path) as e:
            outs = []
            for i in range(4):
                cls = torch.empty(B, width, device="cuda"); patch = torch.empty(B, NP, width, device="cuda")
                e.forward_device(x.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W, False, cls_ptr=cls.data_ptr(), patch_ptr=patch.data_ptr())
                torch.cuda.synchronize()
                outs.append(patch)
            msg = []
            for i in range(1, 4):
                diff = (outs[i] != outs[0])
                imgs = diff.any(dim=-1).any(dim=-1).nonzero().flatten().tolist()
                if imgs:
                    im = imgs[0]
                    toks = diff[im].any(dim=-1).nonzero().flatten()
                    chans = diff[im].any(dim=0).nonzero().flatten()
                    msg.append(f"run {i}: images {imgs}; image {im}: {len(toks)} tokens (first {int(toks[0])}), {len(chans)} channels (first {int(chans[0])}), max |diff| {float((outs[i]-outs[0]).abs().max()):.2e}")
                else:
                    msg.append(f"run {i}: identical")
            print(f"width {width} blocks {L:2d}: " + " | ".join(msg), flush=True)
