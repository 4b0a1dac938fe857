#!/usr/bin/env python
"""Multi-GPU check (run under torchrun on >= 2 GPUs): every rank runs its shard of a global batch through its own
engine, features are all-gathered over NCCL, and rank 0 compares the gathered result with a single-engine run of
the whole batch.   torchrun --nproc-per-node 2 tools/dp_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import dinov2_b200 as d  # noqa: E402
from dinov2_b200 import dp, synth  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = synth.CONFIGS["mini"]
path = "/tmp/dino_dp_mini.gguf"
if rank == 0:
    synth.write_synth_gguf(path, cfg, seed=3)
dist.barrier()
n_items, H, W = 7, 224, 224                      # ragged: 7 images over `world` ranks
imgs = synth.lcg_batch(100, n_items, H, W)
lo, hi = dp.shard_range(n_items, rank, world)
eng = d.Engine(path, device=local)
out = eng.forward(imgs[lo:hi]) if hi > lo else {"cls": np.zeros((0, cfg.hidden_size), np.float32)}
cls_all = dp.all_gather_features(torch.from_numpy(out["cls"]).cuda(), n_items)
patch_all = dp.all_gather_features(torch.from_numpy(out["patch_tokens"]).cuda(), n_items)
if rank == 0:
    full = eng.forward(imgs)
    ok = np.array_equal(cls_all.cpu().numpy(), full["cls"]) and np.array_equal(patch_all.cpu().numpy(), full["patch_tokens"])
    print(f"dp_check world={world} items={n_items}: gathered features bit-identical to single-engine run: {ok}", flush=True)
    assert ok
dist.destroy_process_group()
