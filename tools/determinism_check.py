#!/usr/bin/env python
"""Is the forward pass bit-reproducible, and do the plain and the fused-gather forward agree, on the inputs where the 8-GPU bench
saw differing rows?  usage: python tools/determinism_check.py [first_image ...]   (ViT-L/14, batch 64)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200 as d
from dinov2_b200 import synth
name, B, H, W = os.environ.get("DET_MODEL", "vitl14"), int(os.environ.get("DET_B", "64")), 518, 518
path = f"/tmp/dino_bench/{name}_f16_seed0.gguf"
os.makedirs("/tmp/dino_bench", exist_ok=True)
if not os.path.exists(path):
    synth.write_synth_gguf(path, synth.CONFIGS[name], seed=0)
D = synth.CONFIGS[name].hidden_size
NP = (H // 14) * (W // 14)
cudart = ctypes.CDLL("libcudart.so.12")
firsts = [int(a) for a in sys.argv[1:]] or [64, 384]
with d.Engine(path) as e:
    for first in firsts:
        x = torch.from_numpy(synth.lcg_batch(first, B, H, W)).cuda()
        cls = [torch.empty(B, D, device="cuda") for _ in range(3)]
        patch = [torch.empty(B, NP, D, device="cuda") for _ in range(3)]
        for i in range(3):
            e.forward_device(x.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W, False, cls_ptr=cls[i].data_ptr(), patch_ptr=patch[i].data_ptr())
        torch.cuda.synchronize()
        for i in (1, 2):
            bad = (cls[i] != cls[0]).any(dim=-1).nonzero().flatten().tolist()
            badp = (patch[i] != patch[0]).any(dim=-1).any(dim=-1).nonzero().flatten().tolist()
            print(f"images {first}..: plain run {i} vs run 0: cls rows differing {bad}, images with differing patch tokens {badp}, "
                  f"max |diff| cls {float((cls[i]-cls[0]).abs().max()):.3e} patch {float((patch[i]-patch[0]).abs().max()):.3e}", flush=True)
        if os.environ.get("DET_GATHER", "1") == "0":
            continue
        gbuf, _ = e.gather_init(0, 1, d.GATHER_CLS, B, H, W)
        for i in range(3):
            e.forward_gather_device(x.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W)
            e.synchronize()
            mine = torch.empty(B, D, device="cuda")
            cudart.cudaMemcpy(ctypes.c_void_p(mine.data_ptr()), ctypes.c_void_p(gbuf), B * D * 4, 3)
            bad = (mine != cls[0]).any(dim=-1).nonzero().flatten().tolist()
            print(f"images {first}..: fused-gather run {i} vs plain run 0: cls rows differing {bad}, max |diff| {float((mine-cls[0]).abs().max()):.3e}", flush=True)
