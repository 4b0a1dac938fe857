"""Deterministic synthetic DINOv2 checkpoints and inputs.

Real DINOv2 weights cannot be downloaded in this environment, so tests and
bench.py run on random-init checkpoints written in exactly the tensor manifest
the reference converter produces (reference scripts/dinov2-to-gguf.py:49-166,
dumped in SURVEY.md appendix B) — same names, ggml shapes, dtypes and KV keys,
so the reference's own `dino_model_load` (dinov2.cpp:239-352) accepts them
(tests/test_oracle.py loads every file written here through oracle/_ref).

Unlike HF's init (zero biases, unit LayerNorm gains) every parameter family
gets non-trivial values, so that a dropped bias / LayerScale / register token
shows up as a parity failure.

The synthetic *input* follows SURVEY.md §8(d): a 32-bit LCG per image, values
in [-2, 2], laid out H x W x BGR-interleaved float32 — the layout
`dino_preprocess` hands to `dino_predict` (dinov2.cpp:135-156, 914-931).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import gguf_io as G


@dataclass(frozen=True)
class ModelConfig:
    name: str
    hidden_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_register_tokens: int = 0
    num_classes: int = 1000
    patch_size: int = 14
    img_size: int = 518
    swiglu: bool = False           # reference selects SwiGLU iff num_hidden_layers == 40 (dinov2.cpp:740)

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def mlp_in(self) -> int:       # rows of fc1 / weights_in
        return 8 * ((int(self.hidden_size * 4 * 2 / 3) + 7) // 8) * 2 if self.swiglu else 4 * self.hidden_size

    @property
    def mlp_hidden(self) -> int:   # columns of fc2 / weights_out
        return self.mlp_in // 2 if self.swiglu else self.mlp_in


CONFIGS = {
    # BASELINE.json configs[0..4]
    "vits14": ModelConfig("vits14", 384, 12, 6),
    "vits14_reg4": ModelConfig("vits14_reg4", 384, 12, 6, num_register_tokens=4),
    "vitb14": ModelConfig("vitb14", 768, 12, 12),
    "vitl14": ModelConfig("vitl14", 1024, 24, 16),
    "vitg14": ModelConfig("vitg14", 1536, 40, 24, swiglu=True),
    # small parity-test cases (same architecture family, head_dim 64)
    "tiny": ModelConfig("tiny", 128, 2, 2, num_register_tokens=2, num_classes=10, img_size=70),
    "tiny_noreg": ModelConfig("tiny_noreg", 128, 2, 2, num_classes=10, img_size=70),
    "mini": ModelConfig("mini", 192, 3, 3, num_register_tokens=1, num_classes=24, img_size=224),
    # 40 layers triggers the reference's SwiGLU branch; kept narrow so the CPU oracle stays fast
    "tiny_swiglu": ModelConfig("tiny_swiglu", 192, 40, 3, num_classes=10, img_size=70, swiglu=True),
}


def _weight(rng, shape, std=0.02):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)


def make_tensors(cfg: ModelConfig, seed: int = 0, quant: Optional[str] = None, outliers: bool = False) -> List[G.GGUFTensor]:
    """Tensor list in the converter's file order (state_dict order, fused qkv appended last).

    outliers=True mimics the "massive activation" statistics of trained DINOv2 checkpoints (SURVEY.md appendix D caveat): in
    every block the LayerNorm gains of four fixed channels and the LayerScale factors of four OTHER fixed channels are 8-12x
    larger, so the LN outputs feeding the fp16 GEMM operands and the residual stream carry values an order of magnitude
    above the typical ones.  (Both in the SAME channels, or 30x, drives the network into a regime where the reference no
    longer agrees with itself: its two oracle forms differ by NMSE 4e-2.  At this setting its self-noise is ~3e-7.)
    The draw order of the base generator is unchanged (same seed -> same regular weights)."""
    rng = np.random.default_rng(seed)
    orng = np.random.default_rng(seed + 7919)
    hot = orng.choice(cfg.hidden_size, size=8, replace=False) if outliers else None

    def spike(v, which):
        if hot is None:
            return v
        v = v.copy()
        idx = hot[:4] if which == "ln" else hot[4:]
        v[idx] *= orng.uniform(8.0, 12.0, size=idx.size).astype(np.float32)
        return v
    D, L = cfg.hidden_size, cfg.num_hidden_layers
    P, g = cfg.patch_size, cfg.grid

    def lin(name, w):
        # the reference quantiser touches 2-D tensors whose name matches ".*weight" (dinov2.cpp:227-236)
        w16 = w.astype(np.float16)
        if quant == "q8_0":
            return G.q8_0_tensor(name, w16.astype(np.float32))
        return G.f16_tensor(name, w16)

    out: List[G.GGUFTensor] = []
    out.append(G.f32_tensor("embeddings.cls_token", _weight(rng, (1, 1, D), 0.5)))
    out.append(G.f32_tensor("embeddings.position_embeddings", _weight(rng, (1, 1 + g * g, D), 0.2)))
    if cfg.num_register_tokens:
        out.append(G.f32_tensor("embeddings.register_tokens", _weight(rng, (1, cfg.num_register_tokens, D), 0.5)))
    out.append(G.f16_tensor("embeddings.patch_embeddings.projection.weight", _weight(rng, (D, 3, P, P), 0.03)))
    out.append(G.f32_tensor("embeddings.patch_embeddings.projection.bias", _weight(rng, (1, D, 1, 1), 0.1)))
    qkv = []
    for l in range(L):
        b = f"encoder.layer.{l}."
        wq = _weight(rng, (3 * D, D), 0.04)
        bq = _weight(rng, (3 * D,), 0.1)
        qkv.append((b + "attention.attention.qkv.weight", wq, b + "attention.attention.qkv.bias", bq))
        out.append(lin(b + "attention.output.dense.weight", _weight(rng, (D, D), 0.03)))
        out.append(G.f32_tensor(b + "attention.output.dense.bias", _weight(rng, (D,), 0.05)))
        out.append(G.f32_tensor(b + "layer_scale1.lambda1", spike(0.3 + 0.7 * rng.random(D, dtype=np.float32), "ls")))
        out.append(G.f32_tensor(b + "norm1.weight", spike(1 + _weight(rng, (D,), 0.1), "ln")))
        out.append(G.f32_tensor(b + "norm1.bias", _weight(rng, (D,), 0.05)))
        if cfg.swiglu:
            out.append(lin(b + "mlp.weights_in.weight", _weight(rng, (cfg.mlp_in, D), 0.03)))
            out.append(G.f32_tensor(b + "mlp.weights_in.bias", _weight(rng, (cfg.mlp_in,), 0.05)))
            out.append(lin(b + "mlp.weights_out.weight", _weight(rng, (D, cfg.mlp_hidden), 0.03)))
            out.append(G.f32_tensor(b + "mlp.weights_out.bias", _weight(rng, (D,), 0.05)))
        else:
            out.append(lin(b + "mlp.fc1.weight", _weight(rng, (cfg.mlp_in, D), 0.03)))
            out.append(G.f32_tensor(b + "mlp.fc1.bias", _weight(rng, (cfg.mlp_in,), 0.05)))
            out.append(lin(b + "mlp.fc2.weight", _weight(rng, (D, cfg.mlp_hidden), 0.03)))
            out.append(G.f32_tensor(b + "mlp.fc2.bias", _weight(rng, (D,), 0.05)))
        out.append(G.f32_tensor(b + "layer_scale2.lambda1", spike(0.3 + 0.7 * rng.random(D, dtype=np.float32), "ls")))
        out.append(G.f32_tensor(b + "norm2.weight", spike(1 + _weight(rng, (D,), 0.1), "ln")))
        out.append(G.f32_tensor(b + "norm2.bias", _weight(rng, (D,), 0.05)))
    out.append(G.f32_tensor("layernorm.weight", 1 + _weight(rng, (D,), 0.1)))
    out.append(G.f32_tensor("layernorm.bias", _weight(rng, (D,), 0.05)))
    out.append(lin("classifier.weight", _weight(rng, (cfg.num_classes, 2 * D), 0.05)))
    out.append(G.f32_tensor("classifier.bias", _weight(rng, (cfg.num_classes,), 0.1)))
    for wn, w, bn, b in qkv:
        out.append(lin(wn, w))
        out.append(G.f32_tensor(bn, b))
    return out


def write_synth_gguf(path: str, cfg: ModelConfig, seed: int = 0, quant: Optional[str] = None, outliers: bool = False) -> None:
    ftype = {None: 1, "q8_0": 8}[quant]
    kv = [("general.architecture", G.T_STR, "dinov2")]
    kv += [(str(i), G.T_STR, f"class_{i:04d}") for i in range(cfg.num_classes)]
    kv += [
        ("hidden_size", G.T_U32, cfg.hidden_size),
        ("num_hidden_layers", G.T_U32, cfg.num_hidden_layers),
        ("num_attention_heads", G.T_U32, cfg.num_attention_heads),
        ("num_classes", G.T_U32, cfg.num_classes),
        ("patch_size", G.T_U32, cfg.patch_size),
        ("img_size", G.T_U32, cfg.img_size),
        ("ftype", G.T_U32, ftype),
        ("num_register_tokens", G.T_U32, cfg.num_register_tokens),
    ]
    G.write_gguf(path, kv, make_tensors(cfg, seed, quant, outliers))


def lcg_image(index: int, H: int, W: int) -> np.ndarray:
    """SURVEY.md §8(d) synthetic image: float32 [H, W, 3] BGR-interleaved in [-2, 2]."""
    n = H * W * 3
    # closed form of s_{k+1} = a s_k + c (mod 2^32), vectorised by doubling
    a, c = np.uint64(1664525), np.uint64(1013904223)
    mask = np.uint64(0xFFFFFFFF)
    s = np.empty(n, dtype=np.uint64)
    s0 = np.uint64((12345 + index) & 0xFFFFFFFF)
    s[0] = (s0 * a + c) & mask
    filled = 1
    ak, ck = a, c                      # coefficients of the `filled`-step jump
    while filled < n:
        m = min(filled, n - filled)
        s[filled:filled + m] = (s[:m] * ak + ck) & mask
        ck = (ck * ak + ck) & mask
        ak = (ak * ak) & mask
        filled += m
    v = ((s >> np.uint64(8)) & np.uint64(0xFFFF)).astype(np.float32)
    return (v / np.float32(65535.0) * np.float32(4.0) - np.float32(2.0)).reshape(H, W, 3)


def lcg_batch(start: int, B: int, H: int, W: int) -> np.ndarray:
    return np.stack([lcg_image(start + i, H, W) for i in range(B)])
