"""CPU restatement of the reference DINOv2 forward pass (numpy).

*** TEST INFRASTRUCTURE ONLY. ***  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (dinov2.cpp_b200/) never does and fails loudly without its CUDA
library.

Parity status: PINNED against the reference itself — the unmodified reference
dinov2.cpp + ggml CPU backend compiled from /root/reference by oracle/Makefile
(oracle/_ref/libdino_ref_*.so).  tests/test_oracle.py runs both on the same
gguf + input and the committed fixtures under tests/golden/ were produced by
that reference build (tests/golden/make_golden.py).  The reference repo has no
golden vectors or tests of its own for this path.

Every step cites the reference code it restates.  Numerics contract
(SURVEY.md appendix A): weight GEMMs round the f32 activation to fp16 and
accumulate in f32 (ggml-cpu.c:1278-1363, vec.cpp:128-168); attention QK^T and
PV, LayerNorm, softmax, bias/LayerScale/residual are f32; GELU is the tanh
form through an fp16 table (vec.h:428-457); pixels are rounded to fp16 inside
im2col (ggml.c:4005, ops.cpp:5846).  q8_0 weights: the activation row is
quantised to Q8_0 as well (ggml-cpu.c:256-258, ggml-cpu-quants.c:738); since
d_w*q_w and d_x*q_x are exact in f32 the block-scaled integer dot equals the
f32 dot of the two dequantised operands up to accumulation order.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from typing import Dict, Optional

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gguf_io():
    name = "_oracle_gguf_io"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(_ROOT, "dinov2.cpp_b200", "gguf_io.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


G = _gguf_io()
f32 = np.float32


def r16(x: np.ndarray) -> np.ndarray:
    """round-to-nearest-even to IEEE fp16, widened back (GGML_FP32_TO_FP16 / F16C)."""
    return x.astype(np.float16).astype(np.float32)


def q8_roundtrip(x: np.ndarray) -> np.ndarray:
    """quantize_row_q8_0 then dequantise (ggml-cpu-quants.c:738; block = 32 along K)."""
    m, k = x.shape
    blk = x.reshape(m, k // 32, 32)
    amax = np.abs(blk).max(axis=2)
    d = (amax / f32(127.0)).astype(f32)
    inv = np.where(d != 0, f32(1.0) / np.where(d != 0, d, 1), f32(0.0)).astype(f32)
    x0 = blk * inv[..., None]
    q = np.sign(x0) * np.floor(np.abs(x0) + f32(0.5))
    return (q * r16(d)[..., None]).astype(f32).reshape(m, k)


def q8_1_quantise(x: np.ndarray):
    """quantize_row_q8_1 (ggml-quants.c:220-253): as q8_0 plus s = fp16(d * sum(q)) per block, d being the unrounded scale.
    Returns (dequantised activations q * fp16(d), s) with s of shape [m, k/32]."""
    m, k = x.shape
    blk = x.reshape(m, k // 32, 32)
    amax = np.abs(blk).max(axis=2)
    d = (amax / f32(127.0)).astype(f32)
    inv = np.where(d != 0, f32(1.0) / np.where(d != 0, d, 1), f32(0.0)).astype(f32)
    x0 = blk * inv[..., None]
    q = np.sign(x0) * np.floor(np.abs(x0) + f32(0.5))
    s = r16((q.sum(axis=2) * d).astype(f32))
    return (q * r16(d)[..., None]).astype(f32).reshape(m, k), s


class RefModel:
    """Weights decoded from a gguf written by the reference converter / quantiser."""

    def __init__(self, path: str):
        gg = G.read_gguf(path)
        kv = gg.kv
        self.hidden_size = int(kv["hidden_size"])
        self.num_hidden_layers = int(kv["num_hidden_layers"])
        self.num_attention_heads = int(kv["num_attention_heads"])
        self.num_classes = int(kv.get("num_classes", 0))
        self.patch_size = int(kv["patch_size"])
        self.img_size = int(kv["img_size"])
        self.ftype = int(kv["ftype"]) % 1000            # GGML_QNT_VERSION_FACTOR, dinov2.cpp:307
        self.num_register_tokens = int(kv["num_register_tokens"])
        self.eps = f32(1e-6)                            # dinov2.h:33
        self.types = {n: t.ggml_type for n, t in gg.tensors.items()}
        self.w: Dict[str, np.ndarray] = {n: np.asarray(G.to_numpy(t)) for n, t in gg.tensors.items()}
        # q4_1 / q5_1: the dot product against q8_1 activations needs q*d and the per-block minimum apart
        self.split: Dict[str, tuple] = {n: G.dequantize_legacy(t.data, t.ne, t.ggml_type, split=True) for n, t in gg.tensors.items()
                                        if t.ggml_type in (G.GGML_TYPE_Q4_1, G.GGML_TYPE_Q5_1)}

    def is_q8(self, name: str) -> bool:
        return self.types[name] == G.GGML_TYPE_Q8_0

    def act_type(self, name: str) -> str:
        """vec_dot_type of the weight's type (ggml-cpu.c:214-267): what the activation rows are converted to."""
        t = self.types[name]
        if t in (G.GGML_TYPE_Q8_0, G.GGML_TYPE_Q4_0, G.GGML_TYPE_Q5_0):
            return "q8_0"
        if t in (G.GGML_TYPE_Q4_1, G.GGML_TYPE_Q5_1):
            return "q8_1"
        return "f16"


def mul_mat(model: RefModel, wname: str, x: np.ndarray) -> np.ndarray:
    """ggml_mul_mat(W, x): y[m, n] = sum_k W[n, k] * conv(x)[m, k]   (ggml-cpu.c:1266-1458)."""
    w = model.w[wname]
    w = w.reshape(w.shape[0], -1).astype(f32)
    kind = model.act_type(wname)
    if kind == "q8_0":
        a = q8_roundtrip(x)                             # q8_0, q4_0, q5_0 weights
    elif kind == "q8_1":
        # q4_1 / q5_1 (ggml-cpu-quants.c vec_dot_q4_1_q8_1): sum over blocks of d_w d_a sum(q_w q_a) + m_w * s_a
        a, s_a = q8_1_quantise(x)
        qd, m_w = model.split[wname]
        return a @ qd.reshape(qd.shape[0], -1).astype(f32).T + s_a @ m_w.astype(f32).T
    else:
        a = r16(x)                                      # from_float to vec_dot_type F16
    return a @ w.T


def layer_norm(x: np.ndarray, g: np.ndarray, b: np.ndarray, eps) -> np.ndarray:
    """ggml_norm (ops.cpp:3109-3158): f64 sums, centred variance; then *w +b (dinov2.cpp:694-700)."""
    mean = (x.astype(np.float64).sum(axis=1) / x.shape[1]).astype(f32)
    v = x - mean[:, None]
    var = ((v * v).astype(np.float64).sum(axis=1) / x.shape[1]).astype(f32)
    scale = f32(1.0) / np.sqrt(var + eps, dtype=f32)
    return (v * scale[:, None]) * g + b


def gelu_lut(u: np.ndarray) -> np.ndarray:
    """ggml_vec_gelu_f32 with GGML_GELU_FP16 (vec.h:443-457) + table init (ggml-cpu.c:3395-3403)."""
    v = r16(u)
    t = f32(0.5) * v * (f32(1.0) + np.tanh(f32(0.79788456080286535587989211986876) * v * (f32(1.0) + f32(0.044715) * v * v), dtype=f32))
    out = r16(t)
    out = np.where(u <= f32(-10.0), f32(0.0), out)
    out = np.where(u >= f32(10.0), u, out)
    return out.astype(f32)


def silu(x: np.ndarray) -> np.ndarray:
    """ggml_silu_f32: x / (1 + exp(-x)) (vec.h / vec.cpp:170)."""
    return (x / (f32(1.0) + np.exp(-x, dtype=f32))).astype(f32)


def softmax_rows(s: np.ndarray, scale) -> np.ndarray:
    """ggml_soft_max_ext (ops.cpp:4641-4737): scale first, max, exp, f64 sum, multiply by 1/sum."""
    wp = s * f32(scale)
    mx = wp.max(axis=-1, keepdims=True)
    e = np.exp(wp - mx, dtype=f32)
    inv = (1.0 / e.astype(np.float64).sum(axis=-1, keepdims=True)).astype(f32)
    return e * inv


def _cubic_weights(x):
    A = f32(-0.75)
    x = f32(x)
    c0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A
    c1 = ((A + 2) * x - (A + 3)) * x * x + 1
    c2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1
    return np.array([c0, c1, c2, f32(1.0) - c0 - c1 - c2], dtype=f32)


def resize_cubic(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """cv::resize(INTER_CUBIC) on a float32 [H, W] plane (OpenCV convention: half-pixel centres,
    a = -0.75, replicated border) — what interpolate_pos_embed calls per channel (dinov2.cpp:210)."""
    sh, sw = src.shape
    if (sh, sw) == (dh, dw):
        return src.copy()

    def taps(dn, sn):
        idx = np.zeros((dn, 4), dtype=np.int64)
        wts = np.zeros((dn, 4), dtype=f32)
        scale = sn / dn
        for d in range(dn):
            fx = f32((d + 0.5) * scale - 0.5)
            s = int(np.floor(fx))
            wts[d] = _cubic_weights(fx - f32(s))
            idx[d] = np.clip(np.arange(s - 1, s + 3), 0, sn - 1)
        return idx, wts

    xi, xw = taps(dw, sw)
    yi, yw = taps(dh, sh)
    tmp = None
    for k in range(4):
        term = src[:, xi[:, k]] * xw[:, k][None, :]
        tmp = term if tmp is None else tmp + term
    out = None
    for k in range(4):
        term = tmp[yi[:, k], :] * yw[:, k][:, None]
        out = term if out is None else out + term
    return out.astype(f32)


def interpolate_pos_embed(model: RefModel, H: int, W: int) -> np.ndarray:
    """dinov2.cpp:159-225. Returns [1 + gh*gw, D]."""
    D, ps = model.hidden_size, model.patch_size
    gh, gw = H // ps, W // ps
    M = model.img_size // ps
    pos = model.w["embeddings.position_embeddings"].reshape(-1, D).astype(f32)
    if gh * gw == M * M:                                   # early return keys on the *count* (dinov2.cpp:176)
        return pos[: M * M + 1].copy()
    out = np.empty((1 + gh * gw, D), dtype=f32)
    out[0] = pos[0]
    grid = pos[1:].reshape(M, M, D)
    for c in range(D):
        out[1:, c] = resize_cubic(np.ascontiguousarray(grid[:, :, c]), gh, gw).reshape(-1)
    return out


def forward(model: RefModel, img_bgr_hwc: np.ndarray, classify: bool = False,
            keep: Optional[dict] = None) -> Dict[str, np.ndarray]:
    """One image through the reference graph (dinov2.cpp:616-838). `keep`, if given, receives
    named intermediates for stage-level debugging."""
    D, L, nh = model.hidden_size, model.num_hidden_layers, model.num_attention_heads
    R, ps = model.num_register_tokens, model.patch_size
    hd = D // nh
    H, W, _ = img_bgr_hwc.shape
    gh, gw = H // ps, W // ps
    npatch = gh * gw
    w = model.w

    # dino_predict: BGR interleaved -> RGB planar (dinov2.cpp:914-931)
    rgb = np.ascontiguousarray(img_bgr_hwc[:, :, ::-1].transpose(2, 0, 1)).astype(f32)
    # ggml_conv_2d_sk_p0 = im2col(fp16) + mul_mat (ggml.c:3995-4017, ops.cpp:5784-5855);
    # column index = c*ps*ps + ky*ps + kx, patch p = y*gw + x; pixels past gh*ps / gw*ps are ignored
    cols = rgb[:, : gh * ps, : gw * ps].reshape(3, gh, ps, gw, ps).transpose(1, 3, 0, 2, 4).reshape(npatch, 3 * ps * ps)
    wpe = w["embeddings.patch_embeddings.projection.weight"].reshape(D, -1).astype(f32)
    x = r16(cols) @ wpe.T + w["embeddings.patch_embeddings.projection.bias"].reshape(1, D)
    # [cls ; patches] + pos  (dinov2.cpp:660-671); registers get no pos-embed (:673-685)
    x = np.concatenate([w["embeddings.cls_token"].reshape(1, D), x], axis=0) + interpolate_pos_embed(model, H, W)
    if R > 0:
        x = np.concatenate([x[:1], w["embeddings.register_tokens"].reshape(R, D), x[1:]], axis=0)
    x = x.astype(f32)
    if keep is not None:
        keep["tokens"] = x.copy()
    scale = f32(1.0) / np.sqrt(f32(hd))
    swiglu = L == 40                                       # dinov2.cpp:740
    N = x.shape[0]
    for l in range(L):
        b = f"encoder.layer.{l}."
        h = layer_norm(x, w[b + "norm1.weight"], w[b + "norm1.bias"], model.eps)
        qkv = mul_mat(model, b + "attention.attention.qkv.weight", h) + w[b + "attention.attention.qkv.bias"]
        q = qkv[:, :D].reshape(N, nh, hd).transpose(1, 0, 2)
        k = qkv[:, D:2 * D].reshape(N, nh, hd).transpose(1, 0, 2)
        v = qkv[:, 2 * D:].reshape(N, nh, hd).transpose(1, 0, 2)
        s = q @ k.transpose(0, 2, 1)                       # f32 x f32 (dinov2.cpp:531)
        p = softmax_rows(s, scale)
        o = (p @ v).transpose(1, 0, 2).reshape(N, D)       # f32 x f32 (dinov2.cpp:536-543)
        y = mul_mat(model, b + "attention.output.dense.weight", o) + w[b + "attention.output.dense.bias"]
        x = (y * w[b + "layer_scale1.lambda1"] + x).astype(f32)
        if keep is not None and l == 0:
            keep["qkv0"], keep["attn0"], keep["x_attn0"] = qkv.copy(), o.copy(), x.copy()
        h = layer_norm(x, w[b + "norm2.weight"], w[b + "norm2.bias"], model.eps)
        if swiglu:
            u = mul_mat(model, b + "mlp.weights_in.weight", h) + w[b + "mlp.weights_in.bias"]
            half = u.shape[1] // 2
            m = silu(u[:, :half]) * u[:, half:]
            y = mul_mat(model, b + "mlp.weights_out.weight", m) + w[b + "mlp.weights_out.bias"]
        else:
            u = mul_mat(model, b + "mlp.fc1.weight", h) + w[b + "mlp.fc1.bias"]
            m = gelu_lut(u)
            y = mul_mat(model, b + "mlp.fc2.weight", m) + w[b + "mlp.fc2.bias"]
        x = (y * w[b + "layer_scale2.lambda1"] + x).astype(f32)
        if keep is not None and l == 0:
            keep["x0"] = x.copy()
    x = layer_norm(x, w["layernorm.weight"], w["layernorm.bias"], model.eps)
    out = {"cls": x[0].copy(), "patch_tokens": x[1 + R:].copy(), "tokens": x}
    if classify:
        # pooling includes registers and divides by the constant (img_size/patch)^2 (dinov2.cpp:770-776, 800-803)
        n_embd = model.img_size // ps
        pooled = (x[1:].astype(np.float64).sum(axis=0)).astype(f32) * (f32(1.0) / f32(n_embd * n_embd))
        z = mul_mat(model, "classifier.weight", np.concatenate([x[0], pooled])[None, :])[0] + w["classifier.bias"]
        e = np.exp(z - z.max(), dtype=f32)
        out["logits"] = z.astype(f32)
        out["probs"] = (e * f32(1.0 / e.astype(np.float64).sum())).astype(f32)
    return out
