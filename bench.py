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
  roofline      tensor-core roofline of the dominant kernel family (gemm_f16_tcgen05), event-timed per launch in an extra
                profiled step after the timed loop (the last of five run back to back, so it sees sustained clocks; the timed
                loop itself carries no instrumentation), against MEASURED_PEAKS.json
  per_rank_ms_per_step   min / max / argmax over ranks of the device-timed step (the straggler is visible)
  scale_features (N > 1) the feature-extraction step with the cls all-gather (NCCL over NVLink) inside the timed region
  configs (N = 1)        BASELINE.json configs[1], [2], [4] run for a few steps after the headline
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
    """DRAM bytes per GEMM launch from the committed ncu --set full capture (profiles/r02_gemm_traffic.json), ViT-L b64 only."""
    p = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")
    try:
        with open(p) as f:
            d = json.load(f)
        return {"bytes_per_launch": d["dram_bytes_per_launch_mean"],
                "algorithmic_bytes_per_launch": sum(k["algorithmic_bytes"] for k in d["kernels"]) / len(d["kernels"]),
                "source": "profiles/r02_gemm_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of the qkv / o-proj / fc1 / fc2 launches of one block)"}
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


class Workload:
    """One (model, batch, mode) case: engine, resident inputs / outputs, and the three ways of stepping it."""

    def __init__(self, d, torch, synth, model, quant, batch, classify, rank, local_rank, barrier):
        import numpy as np
        self.d, self.torch = d, torch
        self.cfg = synth.CONFIGS[model]
        self.model, self.quant, self.B, self.classify = model, quant, batch, classify
        self.path = ensure_gguf(model, rank, barrier, quant)
        self.eng = d.Engine(self.path, device=local_rank)
        self.eng.reserve(batch, H, W)
        cfg = self.cfg
        n_unique = min(batch, 8)     # distinct images per rank; generated once (LCG), kept pinned on the host and resident on the device
        base = synth.lcg_batch(rank * batch, n_unique, H, W)
        self.host_in = torch.from_numpy(np.concatenate([base] * ((batch + n_unique - 1) // n_unique))[:batch].copy()).pin_memory()
        self.dev_in = self.host_in.cuda(non_blocking=True)
        self.D, self.C, self.NP = cfg.hidden_size, cfg.num_classes, (H // cfg.patch_size) * (W // cfg.patch_size)
        self.n_tok = 1 + cfg.num_register_tokens + self.NP
        self.dev_cls = torch.empty(batch, self.D, device="cuda")
        self.dev_probs = torch.empty(batch, self.C, device="cuda") if classify else None
        self.dev_patch = torch.empty(batch, self.NP, self.D, device="cuda") if not classify else None
        self.host_out = [self._host_out(), self._host_out()]   # two result sets: the pipelined interface keeps two batches in flight
        self.stream = torch.cuda.Stream()

    def _host_out(self):
        t = self.torch
        o = {"cls": t.empty(self.B, self.D).pin_memory().numpy()}
        if self.classify:
            o["probs"] = t.empty(self.B, self.C).pin_memory().numpy()
            o["logits"] = t.empty(self.B, self.C).pin_memory().numpy()
        else:
            o["patch_tokens"] = t.empty(self.B, self.NP, self.D).pin_memory().numpy()
        return o

    def step_device(self):
        self.eng.forward_device(self.dev_in.data_ptr(), self.d.LAYOUT_BGR_HWC, self.B, H, W, self.classify, cls_ptr=self.dev_cls.data_ptr(),
                                patch_ptr=self.dev_patch.data_ptr() if self.dev_patch is not None else 0,
                                probs_ptr=self.dev_probs.data_ptr() if self.dev_probs is not None else 0, stream=self.stream.cuda_stream)

    def run_pipelined(self, n):
        e, x, o = self.eng, self.host_in.numpy(), self.host_out
        e.submit(x, o[0], classify=self.classify, layout=self.d.LAYOUT_BGR_HWC)
        for i in range(1, n):
            e.submit(x, o[i & 1], classify=self.classify, layout=self.d.LAYOUT_BGR_HWC)
            e.wait()
        e.wait()

    def time_device(self, steps, warmup, barrier, after_step=None, sampler=None):
        """K steps with inputs resident in HBM, CUDA events on the launching stream, NO per-kernel instrumentation inside the
        timed region (the forward is one CUDA-graph replay per step).  Returns (ms total, kernel launches)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            for _ in range(warmup):
                self.step_device()
                if after_step:
                    after_step()
            self.stream.synchronize()
            barrier()
            torch.cuda.synchronize()
            if sampler:
                sampler.start()
            l0 = self.eng.kernel_launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            for _ in range(steps):
                self.step_device()
                if after_step:
                    after_step()
            e1.record(self.stream)
            self.stream.synchronize()
            torch.cuda.synchronize()
            barrier()
            return e0.elapsed_time(e1), self.eng.kernel_launches - l0

    def profile_one_step(self, lead_in=4):
        """Per-kernel event pairs (dino_b200_set_profiling) of ONE step outside the timed region.  `lead_in` profiled steps run
        back to back right before it, so the measured step sees the same sustained, power-capped clocks as the timed loop (a lone
        step after an idle gap runs ~12 % faster at burst clocks and would flatter every kernel)."""
        torch = self.torch
        self.eng.set_profiling(True)
        with torch.cuda.stream(self.stream):
            for _ in range(lead_in + 1):
                self.step_device()
            self.stream.synchronize()
        prof = self.eng.get_profile()
        self.eng.set_profiling(False)
        return prof

    def time_pipelined(self, steps, barrier):
        self.run_pipelined(2)
        barrier()
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        self.run_pipelined(steps)
        return (time.perf_counter() - t0) * 1e3

    def time_synchronous(self, steps, barrier):
        f = lambda: self.eng.forward(self.host_in.numpy(), classify=self.classify, layout=self.d.LAYOUT_BGR_HWC,
                                     want_patch=not self.classify, out=self.host_out[0])
        for _ in range(2):
            f()
        barrier()
        self.torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            f()
        self.torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    def io_bytes(self):
        return self.host_in.numel() * 4, sum(v.nbytes for v in self.host_out[0].values())

    def close(self):
        self.eng.close()


def workload_name(model, quant, batch, classify):
    return (f"{model} {quant or 'f16'} checkpoint, 518x518 fp16-operand/fp32-accumulate forward, batch {batch}/GPU, "
            f"{'classify' if classify else 'features'}")


# BASELINE.json configs[1], [2], [4] (configs[3] is the headline, configs[0] the reference's own CPU case = cpu_baseline)
EXTRA_CONFIGS = [
    {"baseline_config": 1, "model": "vits14_reg4", "quant": None, "batch": 32, "classify": True},
    {"baseline_config": 2, "model": "vitb14", "quant": None, "batch": 64, "classify": False},
    {"baseline_config": 4, "model": "vitg14", "quant": "q8_0", "batch": 16, "classify": True},
]


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
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configs (N = 1 only) after the headline")
    ap.add_argument("--quant", default=None, choices=[None, "q8_0"], help="checkpoint weight format (q8_0: BASELINE configs[4])")
    ap.add_argument("--gather", default="none", choices=["none", "cls", "patch"],
                    help="all-gather the per-image features across ranks inside the HEADLINE step too (default: only in scale_features)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    classify = not args.features

    from dinov2_b200 import synth
    cfg = synth.CONFIGS[args.model]
    workload = workload_name(args.model, args.quant, args.batch, classify)

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

    def max_over_ranks(vals):
        t = torch.tensor(vals, device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def per_rank(val):
        t = torch.tensor([val], device="cuda", dtype=torch.float64)
        if world == 1:
            return [float(val)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(x) for x in out]

    wl = Workload(d, torch, synth, args.model, args.quant, args.batch, classify, rank, local_rank, barrier)
    B = args.batch

    # optional exchange inside the headline step (off by default: the classify path has nothing to exchange)
    gathered, gather_src = None, None
    if world > 1 and args.gather != "none":
        gather_src = wl.dev_cls if args.gather == "cls" or wl.dev_patch is None else wl.dev_patch
        gathered = torch.empty((B * world,) + tuple(gather_src.shape[1:]), device="cuda")
    after = (lambda: dist.all_gather_into_tensor(gathered, gather_src)) if gathered is not None else None

    torch.cuda.synchronize()
    # ---- device-resident throughput: uninstrumented timed loop, then ONE extra profiled step ------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    elapsed_ms, launches = wl.time_device(args.steps, args.warmup, barrier, after_step=after, sampler=sampler)
    clocks = sampler.stop() if sampler else None
    last_prof = wl.profile_one_step()
    rank_ms = per_rank(elapsed_ms / args.steps)

    # ---- end to end through the host-buffer ABI --------------------------------------------------------------------------
    # (a) synchronous call per batch (dino_b200_forward): upload, forward, read-back strictly in sequence
    e2e_sync_ms = wl.time_synchronous(args.steps, barrier)
    barrier()
    # (b) the pipelined interface (dino_b200_submit / dino_b200_wait): the upload of batch k+1 runs under the forward of
    # batch k.  Every step still uploads its own inputs from pinned memory and reads its results back; the timed region is
    # first submit -> last wait, pipeline fill and drain included.
    e2e_ms = wl.time_pipelined(args.steps, barrier)
    barrier()
    elapsed_ms, e2e_ms, e2e_sync_ms = max_over_ranks([elapsed_ms, e2e_ms, e2e_sync_ms])
    finite = bool(np.isfinite(wl.host_out[0]["cls"]).all())

    # ---- N > 1: the feature-extraction path with its one exchange (all-gather of the cls embeddings, NCCL over NVLink) inside
    # the timed step, so that the scaling record contains a real collective and its cost -------------------------------------
    scale_features = None
    if world > 1:
        fw = Workload(d, torch, synth, args.model, args.quant, B, False, rank, local_rank, barrier) if classify else wl
        g_out = torch.empty(B * world, fw.D, device="cuda")
        f_ms, _ = fw.time_device(max(3, args.steps // 2), 3, barrier, after_step=lambda: dist.all_gather_into_tensor(g_out, fw.dev_cls))
        n_f = max(3, args.steps // 2)
        f_ms = max_over_ranks([f_ms])[0]
        # the collective alone, same stream, same buffers
        with torch.cuda.stream(fw.stream):
            for _ in range(3):
                dist.all_gather_into_tensor(g_out, fw.dev_cls)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(fw.stream)
            for _ in range(20):
                dist.all_gather_into_tensor(g_out, fw.dev_cls)
            b.record(fw.stream)
            fw.stream.synchronize()
        coll_us = max_over_ranks([a.elapsed_time(b) / 20 * 1e3])[0]
        scale_features = {"workload": workload_name(args.model, args.quant, B, False) + ", cls all-gather inside the step",
                          "value": B * world * n_f / (f_ms / 1e3), "unit": "images/s", "ms_per_step": f_ms / n_f, "steps": n_f,
                          "collective": "ncclAllGather of [B, D] fp32 cls embeddings (torch.distributed, NVLink)",
                          "gathered_bytes_per_rank": int(B * world * fw.D * 4), "collective_us_alone": coll_us}
        # the engine's own exchange: the final LayerNorm of every rank stores its cls rows straight into every rank's gather
        # buffer (peer memory opened through CUDA IPC handles) — no separate collective kernel at all
        try:
            import ctypes
            gbuf, handle = fw.eng.gather_init(rank, world, d.GATHER_CLS, B, H, W)
            handles = [None] * world
            dist.all_gather_object(handles, handle)
            for r in range(world):
                if r != rank:
                    fw.eng.gather_set_peer(r, ipc_handle=handles[r])
            barrier()

            def step_fused():
                fw.eng.forward_gather_device(fw.dev_in.data_ptr(), d.LAYOUT_BGR_HWC, B, H, W, stream=fw.stream.cuda_stream)

            saved, fw.step_device = fw.step_device, step_fused
            g_ms, _ = fw.time_device(n_f, 3, barrier)
            fw.step_device = saved
            g_ms = max_over_ranks([g_ms])[0]
            # correctness of the exchange: every rank's gather buffer == NCCL all-gather of every rank's cls output
            # (same stream for the forward and the collective: the collective must see the finished dev_cls)
            with torch.cuda.stream(fw.stream):
                fw.step_device()
                dist.all_gather_into_tensor(g_out, fw.dev_cls)
                fw.stream.synchronize()
            torch.cuda.synchronize()
            barrier()
            mine = torch.empty(B * world, fw.D, device="cuda")
            ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(mine.data_ptr()), ctypes.c_void_p(gbuf), B * world * fw.D * 4, 3)
            same = bool(torch.equal(mine, g_out))
            if not same:   # which source rank's rows differ on this rank, and by how much (stderr: stdout is the JSON line)
                bad = [int((mine[r * B:(r + 1) * B] != g_out[r * B:(r + 1) * B]).any(dim=-1).sum()) for r in range(world)]
                print(f"[gather check] rank {rank}: rows differing per source rank {bad}, max |diff| "
                      f"{float((mine - g_out).abs().max()):.3e}, finite {bool(torch.isfinite(mine).all())}/{bool(torch.isfinite(g_out).all())}",
                      file=sys.stderr, flush=True)
            flags = torch.tensor([1.0 if same else 0.0], device="cuda")
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            scale_features["fused_peer_store"] = {"value": B * world * n_f / (g_ms / 1e3), "unit": "images/s", "ms_per_step": g_ms / n_f,
                                                  "api": "dino_b200_gather_init / _set_peer (CUDA IPC) / _forward_gather_device",
                                                  "kernel": "layernorm_gather_kernel: final LayerNorm + stores to every rank's buffer",
                                                  "equals_nccl_all_gather_on_every_rank": bool(flags.item() == 1.0)}
        except Exception as ex:
            scale_features["fused_peer_store"] = {"error": str(ex)[:300]}
        if fw is not wl:
            fw.close()

    configs = None
    if world == 1 and not args.no_configs and args.model == "vitl14" and classify:
        configs = []
        for c in EXTRA_CONFIGS:
            try:
                w2 = Workload(d, torch, synth, c["model"], c["quant"], c["batch"], c["classify"], rank, local_rank, barrier)
                k = 5
                ms, nl = w2.time_device(k, 3, barrier)
                p2 = w2.profile_one_step()
                e2 = w2.time_pipelined(k, barrier)
                fl2 = flops_per_image(w2.cfg, w2.n_tok)
                h2d2, d2h2 = w2.io_bytes()
                configs.append({"baseline_config": c["baseline_config"], "workload": workload_name(c["model"], c["quant"], c["batch"], c["classify"]),
                                "value": c["batch"] * k / (ms / 1e3), "unit": "images/s", "ms_per_step": ms / k, "steps": k, "warmup": 3,
                                "e2e": {"value": c["batch"] * k / (e2 / 1e3), "unit": "images/s", "h2d_bytes_per_step": h2d2, "d2h_bytes_per_step": d2h2},
                                "whole_step_tflops": fl2["total"] * c["batch"] / (ms / k / 1e3) / 1e12,
                                "gemm_tflops": fl2["linear"] * c["batch"] / (p2["gemm_ms"] / 1e3) / 1e12 if p2["gemm_ms"] > 0 else None,
                                "attention_tflops": fl2["attention"] * c["batch"] / (p2["attn_ms"] / 1e3) / 1e12 if p2["attn_ms"] > 0 else None,
                                "gpu_launches": int(nl)})
                w2.close()
            except Exception as ex:      # a side measurement must never cost the headline line
                configs.append({"baseline_config": c["baseline_config"], "error": str(ex)[:200]})

    if rank == 0:
        peaks = load_peaks()
        fl = flops_per_image(cfg, wl.n_tok)
        total_images = B * world * args.steps
        value = total_images / (elapsed_ms / 1e3)
        e2e_value = total_images / (e2e_ms / 1e3)
        ms_per_step = elapsed_ms / args.steps
        gemm_launches_per_step = 1 + 4 * cfg.num_hidden_layers
        traffic = load_gemm_traffic() if (args.model == "vitl14" and B == 64) else None
        gemm_tflops = fl["linear"] * B / (last_prof["gemm_ms"] / 1e3) / 1e12 if last_prof["gemm_ms"] > 0 else 0.0
        attn_tflops = fl["attention"] * B / (last_prof["attn_ms"] / 1e3) / 1e12 if last_prof["attn_ms"] > 0 else 0.0
        step_tflops = fl["total"] * B / (ms_per_step / 1e3) / 1e12
        h2d, d2h = wl.io_bytes()
        line = {
            "metric": "images/sec " + args.model + " 518px forward", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate (bf16-rate tensor cores)",
            "data": "synthetic (LCG images, seeded random-init weights in the reference converter's GGUF manifest)",
            "config": {"workload": workload, "tokens_per_image": wl.n_tok, "gflop_per_image": fl["total"] / 1e9,
                       "l2_policy": "inputs+activations per step (>2 GB) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world}", "feature_all_gather": args.gather, "outputs_finite": finite,
                       "timed_region": "one CUDA-graph replay per step, no per-kernel events (those come from extra profiled steps after the loop)"},
            "clocks": clocks,
            "per_rank_ms_per_step": {"min": min(rank_ms), "max": max(rank_ms), "argmax_rank": int(np.argmax(rank_ms)),
                                     "all": [round(x, 3) for x in rank_ms]},
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
                         "measured_in": "the last of five profiled steps run back to back after the timed loop (direct launches, event pair per kernel, sustained clocks)",
                         "ms_profiled_step": last_prof},
        }
        if scale_features is not None:
            line["scale_features"] = scale_features
        if configs is not None:
            line["configs"] = configs
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb = time_reference(wl.path, 2, 1, classify)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as ex:  # the oracle is a reported baseline, never a dependency of the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": f"unavailable: {ex}"}
        print(json.dumps(line), flush=True)
    wl.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
