#!/usr/bin/env python
"""Multi-process check of the fused final-LayerNorm + peer-store all-gather (CUDA IPC) against ncclAllGather.
usage: torchrun --nproc-per-node N tools/gather_check.py [model] [batch]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import dinov2_b200 as d
from dinov2_b200 import synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "vits14"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = synth.CONFIGS[name]
path = f"/tmp/dino_bench/{name}_f16_seed0.gguf"
os.makedirs("/tmp/dino_bench", exist_ok=True)
if rank == 0 and not os.path.exists(path):
    synth.write_synth_gguf(path + ".tmp", cfg, seed=0); os.replace(path + ".tmp", path)
dist.barrier()
H = W = 518
eng = d.Engine(path, device=local)
x = torch.from_numpy(synth.lcg_batch(rank * B, B, H, W)).cuda()
D = cfg.hidden_size
cls = torch.empty(B, D, device="cuda")
cudart = ctypes.CDLL("libcudart.so.12")
for what, rows in ((d.GATHER_CLS, 1), (d.GATHER_PATCH, (H // 14) * (W // 14))):
    gbuf, handle = eng.gather_init(rank, world, what, B, H, W)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    for r in range(world):
        if r != rank:
            eng.gather_set_peer(r, ipc_handle=handles[r])
    dist.barrier()
    patch = torch.empty(B, rows, D, device="cuda") if what == d.GATHER_PATCH else None
    for it in range(3):
        eng.forward_gather_device(x.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W, cls_ptr=cls.data_ptr(), patch_ptr=patch.data_ptr() if patch is not None else 0)
        eng.synchronize()
        dist.barrier()
    local_rows = cls.view(B, 1, D) if what == d.GATHER_CLS else patch
    want = torch.empty(world * B, rows, D, device="cuda")
    dist.all_gather_into_tensor(want, local_rows.contiguous())
    torch.cuda.synchronize()
    mine = torch.empty(world * B, rows, D, device="cuda")
    cudart.cudaMemcpy(ctypes.c_void_p(mine.data_ptr()), ctypes.c_void_p(gbuf), mine.numel() * 4, 3)
    bad = [(int((mine[r * B:(r + 1) * B] != want[r * B:(r + 1) * B]).any(dim=-1).sum())) for r in range(world)]
    print(f"rank {rank} what={what}: mismatching rows per source rank {bad} (of {B * rows} each)", flush=True)
    dist.barrier()
eng.close()
dist.destroy_process_group()
