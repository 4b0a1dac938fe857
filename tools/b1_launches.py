#!/usr/bin/env python
"""One batch-1 ViT-L forward between cudaProfilerStart/Stop, for an ncu launch list (--profile-from-start off):
   DINO_B200_GRAPH=0 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/b1_launches.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dinov2_b200 as d
from dinov2_b200 import synth
name = sys.argv[1] if len(sys.argv) > 1 else "vitl14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
p = f"/tmp/dino_bench/{name}_f16_seed0.gguf"
os.makedirs("/tmp/dino_bench", exist_ok=True)
if not os.path.exists(p):
    synth.write_synth_gguf(p, synth.CONFIGS[name], seed=0)
cfg = synth.CONFIGS[name]
with d.Engine(p) as e:
    x = torch.from_numpy(synth.lcg_batch(0, B, 518, 518)).cuda()
    cls = torch.empty(B, cfg.hidden_size, device="cuda"); probs = torch.empty(B, cfg.num_classes, device="cuda")
    for i in range(3):
        if i == 2: torch.cuda.cudart().cudaProfilerStart()
        e.forward_device(x.data_ptr(), d.LAYOUT_BGR_HWC, B, 518, 518, True, cls_ptr=cls.data_ptr(), probs_ptr=probs.data_ptr())
        torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
