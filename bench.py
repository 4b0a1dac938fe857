#!/usr/bin/env python
"""bench.py — images/s of the DINOv2 forward pass (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--model vitl14] [--batch 64] [--features]

One "step" = one forward pass over one batch of synthetic 518x518 images (SURVEY.md §8d LCG input,
seeded random-init weights in the reference converter's GGUF manifest).  For N > 1 the driver launches
this file under torchrun (one rank per GPU); the batch is per-GPU (weak scaling, pure data parallel,
no data-path collective — images are independent).

Printed JSON (rank 0, one line):
  value         whole-job images/s with inputs resident in HBM (device pointers through the C ABI,
                CUDA-event timed on the launching stream, max over ranks)
  e2e           same metric through the host-buffer C ABI call (dino_b200_forward): H2D of the batch from
                pinned memory + D2H of the result inside the timed region
  roofline      tensor-core roofline of the dominant kernel family (gemm_f16_tcgen05), event-timed per launch
                inside the timed region, against MEASURED_PEAKS.json
  cpu_baseline  the reference's own ggml CPU path (oracle/_ref, built from /root/reference) on this box's cores
--impl reference times that CPU path alone (rank 0; other ranks exit)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 518


def flops_per_image(cfg, n_tok: int) -> dict:
    """L*(24 N D^2 + 4 N^2 D) + 2*NP*588*D (multiply-add = 2)."""
    D, L = cfg.hidden_size, cfg.num_hidden_layers
    npatch = (H // cfg.patch_size) * (W // cfg.patch_size)
    mlp = 2 * n_tok * D * (cfg.mlp_in + cfg.mlp_hidden)
    linear = L * (2 * n_tok * D * 3 * D + 2 * n_tok * D * D + mlp) + 2 * npatch * 588 * D
    attn = L * 4 * n_tok * n_tok * D
    return {"linear": float(linear), "attention": float(attn), "total": float(linear + attn)}


def load_gemm_traffic():
    """DRAM bytes per GEMM launch from the committed ncu --set full capture (profiles/r01_gemm_traffic.json), ViT-L b64 only."""
    p = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return {"bytes_per_launch": d["dram_bytes_per_launch_mean"],
                "algorithmic_bytes_per_launch": sum(k["algorithmic_bytes"] for k in d["kernels"]) / len(d["kernels"]),
                "source": "profiles/r01_gemm_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of the qkv / o-proj / fc1 / fc2 launches of one block)"}
    except Exception:
        return None


def load_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"burst": float(d["bf16_tflops"]), "sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "hbm": float(d["hbm_gbs"]), "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        # samples under load = upper half (the sampler also sees the idle edges of the region)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def ensure_gguf(name: str, rank: int, barrier, quant=None) -> str:
    from dinov2_b200 import synth
    d = os.environ.get("DINO_BENCH_DIR", "/tmp/dino_bench")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, f"{name}_{quant or 'f16'}_seed0.gguf")
    if rank == 0 and not os.path.exists(path):
        tmp = path + f".tmp{os.getpid()}"
        synth.write_synth_gguf(tmp, synth.CONFIGS[name], seed=0, quant=quant)
        os.replace(tmp, path)
    barrier()
    return path


class stdout_to_stderr:
    """The reference library printf()s its hyper-parameter dump (dinov2.cpp:288-299) and NCCL its version banner to fd 1;
    stdout is reserved for the one JSON line, so those writes are sent to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def time_reference(path: str, steps: int, warmup: int, classify: bool) -> dict:
    """The reference's own CPU implementation (oracle/_ref) on one LCG image per step, all host threads."""
    with stdout_to_stderr():
        return _time_reference(path, steps, warmup, classify)


def _time_reference(path: str, steps: int, warmup: int, classify: bool) -> dict:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref as refmod
    from dinov2_b200 import synth
    if not refmod.available():
        raise RuntimeError("oracle/_ref is not built")
    cores = os.cpu_count() or 1
    R = refmod.Reference(path, classify=classify, n_threads=cores, H=H, W=W)
    img = synth.lcg_image(0, H, W)
    for _ in range(warmup):
        R.forward(img)
    t0 = time.perf_counter()
    for _ in range(steps):
        R.forward(img)
    dt = time.perf_counter() - t0
    R.close()
    return {"value": steps / dt, "unit": "images/s", "cores": cores, "kind": "reference",
            "sample": f"{steps} x 1 image of the workload (batch 1: the reference cannot batch), {warmup} warm-up",
            "ms_per_image": dt / steps * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vitl14")
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--features", action="store_true", help="feature-extraction mode (patch tokens) instead of classify")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quant", default=None, choices=[None, "q8_0"], help="checkpoint weight format (q8_0: BASELINE configs[4])")
    ap.add_argument("--gather", default="none", choices=["none", "cls", "patch"],
                    help="all-gather the per-image features across ranks inside the timed step (NCCL over NVLink)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    classify = not args.features

    from dinov2_b200 import synth
    cfg = synth.CONFIGS[args.model]
    n_tok = 1 + cfg.num_register_tokens + (H // cfg.patch_size) * (W // cfg.patch_size)
    workload = (f"{args.model} {args.quant or 'f16'} checkpoint, 518x518 fp16-operand/fp32-accumulate forward, batch {args.batch}/GPU, "
                f"{'classify' if classify else 'features'}")

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        path = ensure_gguf(args.model, 0, lambda: None, args.quant)
        steps = max(1, args.steps)
        cb = time_reference(path, steps, max(0, min(args.warmup, 1)), classify)
        line = {"impl": "reference", "metric": "images/sec " + args.model + " 518px forward", "value": cb["value"], "unit": "images/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_image"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate (ggml CPU)",
                "data": "synthetic (LCG image, seeded random-init weights)",
                "config": {"workload": workload, "note": "reference ggml CPU path, 1 image per step"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    import numpy as np
    import torch
    import torch.distributed as dist
    import dinov2_b200 as d

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: dinov2_b200 has no CPU fallback"}), flush=True)
        return 1
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator comes up; stdout is reserved for the one JSON line
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    path = ensure_gguf(args.model, rank, barrier, args.quant)
    eng = d.Engine(path, device=local_rank)
    B = args.batch
    eng.reserve(B, H, W)

    # distinct images per rank; generated once (LCG), kept pinned on the host and resident on the device
    n_unique = min(B, 8)
    base = synth.lcg_batch(rank * B, n_unique, H, W)
    host_in = torch.from_numpy(np.concatenate([base] * ((B + n_unique - 1) // n_unique))[:B].copy()).pin_memory()
    dev_in = host_in.cuda(non_blocking=True)
    D, C, NP = cfg.hidden_size, cfg.num_classes, (H // cfg.patch_size) * (W // cfg.patch_size)
    dev_cls = torch.empty(B, D, device="cuda")
    dev_probs = torch.empty(B, C, device="cuda") if classify else None
    dev_patch = torch.empty(B, NP, D, device="cuda") if not classify else None
    def make_host_out():
        o = {"cls": torch.empty(B, D).pin_memory().numpy()}
        if classify:
            o["probs"] = torch.empty(B, C).pin_memory().numpy()
            o["logits"] = torch.empty(B, C).pin_memory().numpy()
        else:
            o["patch_tokens"] = torch.empty(B, NP, D).pin_memory().numpy()
        return o
    host_out = make_host_out()
    host_out2 = make_host_out()          # second result set for the pipelined interface (two batches in flight)
    stream = torch.cuda.Stream()
    gathered = None
    if world > 1 and args.gather != "none":
        from dinov2_b200 import dp
        src = dev_cls if args.gather == "cls" or dev_patch is None else dev_patch
        gathered = torch.empty((B * world,) + tuple(src.shape[1:]), device="cuda")

    def step_device():
        eng.forward_device(dev_in.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W, classify, cls_ptr=dev_cls.data_ptr(),
                           patch_ptr=dev_patch.data_ptr() if dev_patch is not None else 0,
                           probs_ptr=dev_probs.data_ptr() if dev_probs is not None else 0, stream=stream.cuda_stream)
        if gathered is not None:      # the only exchange of the path: every rank ends up with the whole batch's features
            dist.all_gather_into_tensor(gathered, dev_cls if args.gather == "cls" or dev_patch is None else dev_patch)

    def step_host():
        eng.forward(host_in.numpy(), classify=classify, layout=d.LAYOUT_BGR_HWC, want_patch=not classify, out=host_out)

    torch.cuda.synchronize()
    # ---- device-resident throughput -------------------------------------------------------------
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_device()
        stream.synchronize()
        barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        eng.set_profiling(True)
        launches0 = eng.kernel_launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        prof_sum = {"gemm_ms": 0.0, "attn_ms": 0.0, "other_ms": 0.0, "total_ms": 0.0}
        e0.record(stream)
        for _ in range(args.steps):
            step_device()
        e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        barrier()
        launches = eng.kernel_launches - launches0
        elapsed_ms = e0.elapsed_time(e1)
        last_prof = eng.get_profile()          # event pairs of the last timed step
        eng.set_profiling(False)
        clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer ABI -------------------------------------------------
    # (a) synchronous call per batch (dino_b200_forward): upload, forward, read-back strictly in sequence
    for _ in range(2):
        step_host()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # (b) the pipelined interface (dino_b200_submit / dino_b200_wait): the upload of batch k+1 runs under the forward of
    # batch k.  Every step still uploads its own inputs from pinned memory and reads its results back; the timed region is
    # first submit -> last wait, pipeline fill and drain included.
    host_np = host_in.numpy()
    outs = (host_out, host_out2)

    def run_pipelined(n):
        eng.submit(host_np, outs[0], classify=classify, layout=d.LAYOUT_BGR_HWC)
        for i in range(1, n):
            eng.submit(host_np, outs[i & 1], classify=classify, layout=d.LAYOUT_BGR_HWC)
            eng.wait()
        eng.wait()

    run_pipelined(2)
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_pipelined(args.steps)
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    t = torch.tensor([elapsed_ms, e2e_ms, e2e_sync_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, e2e_sync_ms = float(t[0]), float(t[1]), float(t[2])
    finite = bool(np.isfinite(host_out["cls"]).all())

    if rank == 0:
        peaks = load_peaks()
        fl = flops_per_image(cfg, n_tok)
        total_images = B * world * args.steps
        value = total_images / (elapsed_ms / 1e3)
        e2e_value = total_images / (e2e_ms / 1e3)
        ms_per_step = elapsed_ms / args.steps
        gemm_launches_per_step = 1 + 4 * cfg.num_hidden_layers
        traffic = load_gemm_traffic() if (args.model == "vitl14" and B == 64) else None
        gemm_tflops = fl["linear"] * B / (last_prof["gemm_ms"] / 1e3) / 1e12 if last_prof["gemm_ms"] > 0 else 0.0
        attn_tflops = fl["attention"] * B / (last_prof["attn_ms"] / 1e3) / 1e12 if last_prof["attn_ms"] > 0 else 0.0
        step_tflops = fl["total"] * B / (ms_per_step / 1e3) / 1e12
        h2d = host_in.numel() * 4
        d2h = sum(v.nbytes for v in host_out.values())
        line = {
            "metric": "images/sec " + args.model + " 518px forward", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate (bf16-rate tensor cores)",
            "data": "synthetic (LCG images, seeded random-init weights in the reference converter's GGUF manifest)",
            "config": {"workload": workload, "tokens_per_image": n_tok, "gflop_per_image": fl["total"] / 1e9,
                       "l2_policy": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world}", "feature_all_gather": args.gather, "outputs_finite": finite},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps,
                    "api": "dino_b200_submit/dino_b200_wait (two batches in flight: upload of batch k+1 under the forward of batch k)",
                    "synchronous_value": total_images / (e2e_sync_ms / 1e3), "synchronous_api": "dino_b200_forward",
                    "synchronous_ms_per_step": e2e_sync_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_f16_tcgen05 (all weight GEMMs, %d launches/step)" % gemm_launches_per_step,
                         "achieved": gemm_tflops, "peak": peaks["sustained"], "unit": "TFLOP/s",
                         "frac": gemm_tflops / peaks["sustained"], "peak_kind": peaks["source"] + " sustained bf16 (kernel timed inside a long step)",
                         "traffic": (traffic["bytes_per_launch"] if traffic else None), "traffic_detail": traffic,
                         "attention_tflops": attn_tflops, "whole_step_tflops": step_tflops,
                         "whole_step_frac_of_burst": step_tflops / peaks["burst"],
                         "ms_last_step": last_prof},
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = time_reference(path, 2, 1, classify)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the oracle is a reported baseline, never a dependency of the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"unavailable: {ex}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
