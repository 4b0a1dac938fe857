#!/usr/bin/env python
"""Bring-up / diagnosis harness for the GPU box: runs each kernel check in its own subprocess (a trap or
sticky CUDA error in one must not poison the rest), prints error statistics instead of asserting, and writes
everything to gpurun_out/gpu_check.log.   usage: python tools/gpu_check.py [check ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def stats(name, got, ref):
    import torch
    got = got.float().cpu()
    ref = ref.float().cpu()
    d = (got - ref).abs()
    nmse = float(((got - ref) ** 2).sum() / (ref ** 2).sum().clamp_min(1e-30))
    bad = int((~torch.isfinite(got)).sum())
    print(f"  {name}: max_abs {float(d.max()):.4e} mean_abs {float(d.mean()):.4e} nmse {nmse:.3e} ref_abs {float(ref.abs().mean()):.3e} nonfinite {bad}",
          flush=True)
    return nmse


def check_gemm(args):
    import torch
    from dinov2_b200 import engine as E
    torch.manual_seed(0)
    dev = "cuda"
    shapes = [(128, 128, 64), (128, 256, 128), (200, 128, 128), (300, 384, 384), (2740, 1152, 384), (4096, 1024, 1024), (2740, 3072, 1024), (1000, 512, 640)]
    for (M, N, K) in shapes:
        A = (torch.randn(M, K, device=dev) * 0.5).half()
        W = (torch.randn(N, K, device=dev) * 0.05).half()
        bias = torch.randn(N, device=dev) * 0.1
        ref = A.float() @ W.float().t() + bias
        print(f"gemm M={M} N={N} K={K}", flush=True)
        # BIAS_F16
        out = torch.zeros(M, N, device=dev, dtype=torch.half)
        E.kernel_gemm(E.EPI_BIAS_F16, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), 0, out.data_ptr(), N)
        torch.cuda.synchronize()
        stats("bias_f16", out, ref)
        # GELU
        out = torch.zeros(M, N, device=dev, dtype=torch.half)
        E.kernel_gemm(E.EPI_GELU_F16, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), 0, out.data_ptr(), N)
        torch.cuda.synchronize()
        v = ref.half().float()
        g = (0.5 * v * (1 + torch.tanh(0.7978845608028654 * v * (1 + 0.044715 * v * v)))).half().float()
        stats("gelu_f16", out, g)
        # RESID
        ls = torch.rand(N, device=dev) + 0.3
        X = torch.randn(M, N, device=dev)
        X0 = X.clone()
        E.kernel_gemm(E.EPI_RESID_F32, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr(), X.data_ptr(), N)
        torch.cuda.synchronize()
        stats("resid_f32", X, X0 + ls * ref)
        # PATCH (np=100 patches per image, 3 prefix tokens)
        np_, toff = 100, 3
        if M % np_ == 0:
            nimg = M // np_
            ntok = toff + np_
            pos = torch.randn(1 + np_, N, device=dev)
            Xp = torch.zeros(nimg * ntok, N, device=dev)
            E.kernel_gemm(E.EPI_PATCH_F32, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), 0, Xp.data_ptr(), N, pos.data_ptr(), np_, ntok, toff)
            torch.cuda.synchronize()
            want = torch.zeros_like(Xp).view(nimg, ntok, N)
            want[:, toff:] = ref.view(nimg, np_, N) + pos[1:]
            stats("patch_f32", Xp, want.view(-1, N))
        # SWIGLU (N = gate+up rows, interleaved per 256)
        if N % 256 == 0:
            hid = N // 2
            Wg = (torch.randn(hid, K, device=dev) * 0.05).half()
            Wu = (torch.randn(hid, K, device=dev) * 0.05).half()
            bg = torch.randn(hid, device=dev) * 0.1
            bu = torch.randn(hid, device=dev) * 0.1
            Wi = torch.empty(N, K, device=dev, dtype=torch.half)
            bi = torch.empty(N, device=dev)
            j = torch.arange(hid, device=dev)
            gi = (j // 128) * 256 + (j % 128)
            Wi[gi] = Wg; Wi[gi + 128] = Wu; bi[gi] = bg; bi[gi + 128] = bu
            out = torch.zeros(M, hid, device=dev, dtype=torch.half)
            E.kernel_gemm(E.EPI_SWIGLU_F16, A.data_ptr(), K, Wi.data_ptr(), K, M, N, K, bi.data_ptr(), 0, out.data_ptr(), hid)
            torch.cuda.synchronize()
            gt = A.float() @ Wg.float().t() + bg
            ut = A.float() @ Wu.float().t() + bu
            stats("swiglu_f16", out, torch.nn.functional.silu(gt) * ut)


def check_attention(args):
    import torch
    from dinov2_b200 import engine as E
    torch.manual_seed(0)
    dev = "cuda"
    for (B, N, D) in [(1, 28, 128), (2, 128, 64), (2, 200, 128), (2, 1370, 384), (3, 1374, 128), (1, 2171, 64)]:
        H = D // 64
        qkv = (torch.randn(B * N, 3 * D, device=dev)).half()
        out = torch.zeros(B * N, D, device=dev, dtype=torch.half)
        E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D)
        torch.cuda.synchronize()
        q, k, v = [t.view(B, N, H, 64).permute(0, 2, 1, 3) for t in qkv.float().split(D, dim=1)]
        p = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
        ref = (p @ v).permute(0, 2, 1, 3).reshape(B * N, D)
        print(f"attention B={B} N={N} D={D}", flush=True)
        stats("out", out, ref)


def check_layernorm(args):
    import torch
    from dinov2_b200 import engine as E
    torch.manual_seed(0)
    dev = "cuda"
    for (rows, D) in [(28, 128), (100, 192), (2740, 384), (1000, 768), (4096, 1024), (999, 1536)]:
        X = torch.randn(rows, D, device=dev) * 2 + 0.5
        g = torch.randn(D, device=dev)
        b = torch.randn(D, device=dev)
        ref = torch.nn.functional.layer_norm(X, (D,), g, b, 1e-6)
        o16 = torch.zeros(rows, D, device=dev, dtype=torch.half)
        o32 = torch.zeros(rows, D, device=dev)
        E.kernel_layernorm(X.data_ptr(), g.data_ptr(), b.data_ptr(), o16.data_ptr(), rows, D, 1e-6, True)
        E.kernel_layernorm(X.data_ptr(), g.data_ptr(), b.data_ptr(), o32.data_ptr(), rows, D, 1e-6, False)
        torch.cuda.synchronize()
        print(f"layernorm rows={rows} D={D}", flush=True)
        stats("f16", o16, ref)
        stats("f32", o32, ref)


def _model_case(name, quant, H, W, B, classify, seed=1, use_ref=True):
    import numpy as np
    import dinov2_b200 as d
    from dinov2_b200 import synth
    import restate
    import ref as refmod
    cfg = synth.CONFIGS[name]
    os.makedirs("/tmp/dino_w", exist_ok=True)
    path = f"/tmp/dino_w/{name}_{quant or 'f16'}_{seed}.gguf"
    if not os.path.exists(path):
        synth.write_synth_gguf(path, cfg, seed=seed, quant=quant)
    imgs = synth.lcg_batch(0, B, H, W)
    t0 = time.time()
    eng = d.Engine(path)
    t1 = time.time()
    out = eng.forward(imgs, classify=classify)
    t2 = time.time()
    print(f"model {name} quant={quant} {H}x{W} B={B} classify={classify}: load {t1 - t0:.2f}s forward {t2 - t1:.3f}s launches {eng.kernel_launches}", flush=True)
    import torch
    m = restate.RefModel(path)
    for i in sorted(set([0, B - 1])):
        r = restate.forward(m, imgs[i], classify=classify)
        keys = ["cls", "patch_tokens"] + (["logits", "probs"] if classify else [])
        for k in keys:
            stats(f"img{i} {k} vs restate", torch.from_numpy(out[k][i]), torch.from_numpy(r[k]))
        if classify:
            print(f"    top1 engine {int(out['probs'][i].argmax())} restate {int(r['probs'].argmax())}", flush=True)
    if use_ref and refmod.available():
        R = refmod.Reference(path, classify=classify, H=H, W=W)
        o = R.forward(imgs[0])
        for k in (["logits", "probs"] if classify else ["cls", "patch_tokens"]):
            stats(f"img0 {k} vs REFERENCE", torch.from_numpy(out[k][0]), torch.from_numpy(o[k]))
        print(f"    reference CPU time {R.last_ms:.1f} ms ({R.n_threads} threads)", flush=True)
        R.close()
    eng.close()


def check_model_tiny(args):
    _model_case("tiny", None, 70, 70, 3, False)
    _model_case("tiny", None, 70, 70, 2, True)
    _model_case("tiny_noreg", None, 98, 84, 2, False)
    _model_case("tiny", "q8_0", 70, 70, 2, True)
    _model_case("tiny_swiglu", None, 70, 70, 2, True)


def check_model_vits(args):
    _model_case("vits14", None, 518, 518, 2, False, seed=0)
    _model_case("vits14_reg4", None, 518, 518, 2, True, seed=0)


def check_bench_quick(args):
    """device-resident forward timing for any config: bench_quick:<model>:<batch>[:quant]"""
    import numpy as np
    import torch
    import dinov2_b200 as d
    from dinov2_b200 import synth
    name = args[0] if args else "vitl14"
    B = int(args[1]) if len(args) > 1 else 64
    quant = args[2] if len(args) > 2 and args[2] != "f16" else None
    cfg = synth.CONFIGS[name]
    path = f"/tmp/dino_w/{name}_{quant or 'f16'}_0.gguf"
    os.makedirs("/tmp/dino_w", exist_ok=True)
    t0 = time.time()
    if not os.path.exists(path):
        synth.write_synth_gguf(path, cfg, seed=0, quant=quant)
    t1 = time.time()
    eng = d.Engine(path)
    t2 = time.time()
    imgs = torch.from_numpy(synth.lcg_batch(0, 2, 518, 518)).cuda()
    imgs = imgs.repeat((B + 1) // 2, 1, 1, 1)[:B].contiguous()
    cls = torch.empty(B, cfg.hidden_size, device="cuda")
    probs = torch.empty(B, cfg.num_classes, device="cuda")
    stream = torch.cuda.Stream()
    st = stream.cuda_stream
    eng.reserve(B, 518, 518)
    n_tok = 1 + cfg.num_register_tokens + 37 * 37
    D, L = cfg.hidden_size, cfg.num_hidden_layers
    gflop = (L * (2 * n_tok * D * 3 * D + 2 * n_tok * D * D + 2 * n_tok * D * (cfg.mlp_in + cfg.mlp_hidden) + 4 * n_tok * n_tok * D)
             + 2 * 1369 * 588 * D) / 1e9
    with torch.cuda.stream(stream):
        for _ in range(3):
            eng.forward_device(imgs.data_ptr(), 1, B, 518, 518, True, cls_ptr=cls.data_ptr(), probs_ptr=probs.data_ptr(), stream=st)
        stream.synchronize()
        eng.set_profiling(True)
        eng.forward_device(imgs.data_ptr(), 1, B, 518, 518, True, cls_ptr=cls.data_ptr(), probs_ptr=probs.data_ptr(), stream=st)
        stream.synchronize()
        prof = eng.get_profile()
        eng.set_profiling(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n = 5
        for _ in range(n):
            eng.forward_device(imgs.data_ptr(), 1, B, 518, 518, True, cls_ptr=cls.data_ptr(), probs_ptr=probs.data_ptr(), stream=st)
        e1.record(stream)
        stream.synchronize()
    ms = e0.elapsed_time(e1) / n
    ok = bool(torch.isfinite(cls).all()) and bool(torch.isfinite(probs).all()) and bool(((probs.sum(1) - 1).abs() < 1e-4).all())
    print(f"bench {name} quant={quant} B={B}: gguf {t1 - t0:.1f}s load {t2 - t1:.1f}s | {ms:.2f} ms/step  {B / ms * 1000:.1f} img/s  "
          f"{gflop * B / ms:.0f} TFLOP/s ({gflop * B / ms / 1676.0 * 100:.1f}% of burst roofline)  outputs ok={ok}", flush=True)
    print("  profile(ms)", {k: round(v, 2) for k, v in prof.items()}, flush=True)


def check_attn_bench(args):
    """time the attention kernel alone at the ViT-L bench shape; variant via env DINO_B200_ATTN / _PINGPONG"""
    import torch
    from dinov2_b200 import engine as E
    B, N, D = (int(a) for a in args[:3]) if len(args) >= 3 else (64, 1370, 1024)
    qkv = torch.randn(B * N, 3 * D, device="cuda").half()
    out = torch.zeros(B * N, D, device="cuda", dtype=torch.half)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D, st.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        n = 10
        for _ in range(n):
            E.kernel_attention(qkv.data_ptr(), out.data_ptr(), B, N, D, st.cuda_stream)
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * N * N * D * B
    print(f"attn_bench B={B} N={N} D={D} variant={os.environ.get('DINO_B200_ATTN','3')} pp={os.environ.get('DINO_B200_ATTN_PINGPONG','1')}: "
          f"{ms*1000:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def check_gemm_bench(args):
    """time each GEMM shape of a ViT-L layer alone"""
    import torch
    from dinov2_b200 import engine as E
    M = int(args[0]) if args else 87680
    st = torch.cuda.Stream()
    for name, epi, N, K in [("qkv", E.EPI_BIAS_F16, 3072, 1024), ("proj", E.EPI_RESID_F32, 1024, 1024),
                            ("fc1", E.EPI_GELU_F16, 4096, 1024), ("fc2", E.EPI_RESID_F32, 1024, 4096)]:
        A = (torch.randn(M, K, device="cuda") * 0.5).half()
        W = (torch.randn(N, K, device="cuda") * 0.05).half()
        bias = torch.randn(N, device="cuda") * 0.1
        ls = torch.rand(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if epi == E.EPI_RESID_F32 else torch.half)
        with torch.cuda.stream(st):
            for _ in range(3):
                E.kernel_gemm(epi, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr(), out.data_ptr(), N, stream=st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            n = 10
            for _ in range(n):
                E.kernel_gemm(epi, A.data_ptr(), K, W.data_ptr(), K, M, N, K, bias.data_ptr(), ls.data_ptr(), out.data_ptr(), N, stream=st.cuda_stream)
            e1.record(st)
            st.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"gemm_bench {name} M={M} N={N} K={K}: {ms*1000:.1f} us  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


CHECKS = {"gemm": check_gemm, "attention": check_attention, "layernorm": check_layernorm, "model_tiny": check_model_tiny,
          "model_vits": check_model_vits, "bench_quick": check_bench_quick, "attn_bench": check_attn_bench, "gemm_bench": check_gemm_bench}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        CHECKS[sys.argv[2]](sys.argv[3:])
        sys.exit(0)
    todo = sys.argv[1:] or ["layernorm", "gemm", "attention", "model_tiny", "model_vits"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "gpu_check.log"), "a")
    for item in todo:
        name, *a = item.split(":")
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name] + a, capture_output=True, text=True, timeout=600)
            txt = f"=== {item} rc={r.returncode} ({time.time() - t0:.1f}s)\n{r.stdout}\n{r.stderr[-4000:]}\n"
        except subprocess.TimeoutExpired as ex:
            txt = f"=== {item} TIMEOUT\n{(ex.stdout or b'').decode(errors='replace') if isinstance(ex.stdout, bytes) else ex.stdout}\n"
        print(txt, flush=True)
        log.write(txt)
        log.flush()
