"""The oracle itself (CPU): numpy restatement vs the committed golden vectors (produced by the unmodified
reference build, tests/golden/make_golden.py) and, where oracle/_ref is present, vs the reference live."""
import os

import numpy as np
import pytest

import dinov2_b200  # noqa: F401
from dinov2_b200 import synth
import ref as refmod
import restate
from conftest import nmse

GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = np.load(os.path.join(GOLD, "golden.npz"))
F16 = os.path.join(GOLD, "tiny_f16.gguf")
Q8 = os.path.join(GOLD, "tiny_q8_0.gguf")

# restatement-vs-reference bounds: f32 accumulation order only (observed ~1e-9 on the tiny model)
TOL_F16 = 1e-7
# q8_0: the reference's activation quantiser flips roundings between builds (~7e-5 on ViT-S)
TOL_Q8 = 2e-5


def test_restatement_features_golden():
    m = restate.RefModel(F16)
    r = restate.forward(m, synth.lcg_image(0, 70, 70))
    assert nmse(r["cls"], G["f16_feat_cls"]) < TOL_F16
    assert nmse(r["patch_tokens"], G["f16_feat_patch"]) < TOL_F16
    assert np.abs(r["patch_tokens"] - G["f16_feat_patch"]).max() < 1e-3


def test_restatement_classify_golden():
    m = restate.RefModel(F16)
    r = restate.forward(m, synth.lcg_image(0, 70, 70), classify=True)
    assert nmse(r["logits"], G["f16_cls_logits"]) < TOL_F16
    assert nmse(r["probs"], G["f16_cls_probs"]) < TOL_F16
    assert int(r["probs"].argmax()) == int(G["f16_cls_probs"].argmax())
    assert abs(float(r["probs"].sum()) - 1.0) < 1e-5


def test_restatement_non_native_grid_golden():
    """98x84 input -> 7x6 patch grid: exercises the bicubic pos-embed resampling (dinov2.cpp:159-225)."""
    m = restate.RefModel(F16)
    pos = restate.interpolate_pos_embed(m, 98, 84)
    assert pos.shape == G["f16_nn_pos"].shape
    assert np.abs(pos - G["f16_nn_pos"]).max() < 2e-6
    r = restate.forward(m, synth.lcg_image(3, 98, 84))
    assert nmse(r["patch_tokens"], G["f16_nn_patch"]) < TOL_F16
    assert nmse(r["cls"], G["f16_nn_cls"]) < TOL_F16


def test_restatement_q8_0_golden():
    m = restate.RefModel(Q8)
    assert m.is_q8("classifier.weight") and not m.is_q8("embeddings.patch_embeddings.projection.weight")
    r = restate.forward(m, synth.lcg_image(0, 70, 70), classify=True)
    assert nmse(r["patch_tokens"], G["q8_feat_patch"]) < TOL_Q8
    assert nmse(r["logits"], G["q8_cls_logits"]) < 20 * TOL_Q8
    assert int(r["probs"].argmax()) == int(G["q8_cls_probs"].argmax())


def test_identity_pos_embed_keyed_on_patch_count():
    m = restate.RefModel(F16)
    p = restate.interpolate_pos_embed(m, 70, 70)
    assert np.array_equal(p, m.w["embeddings.position_embeddings"].reshape(-1, 128))


def test_gelu_table_semantics():
    u = np.array([-11.0, -10.0, -3.0, -0.1, 0.0, 0.7, 3.3, 10.0, 12.5], dtype=np.float32)
    g = restate.gelu_lut(u)
    assert g[0] == 0 and g[1] == 0 and g[-1] == u[-1] and g[-2] == u[-2]
    mid = g[2:-2]
    assert np.array_equal(mid, mid.astype(np.float16).astype(np.float32))     # table entries are fp16 values


def test_classify_pooling_quirk():
    """Pooling divides by the constant (img_size/patch)^2 and includes register tokens (dinov2.cpp:770-803)."""
    m = restate.RefModel(F16)
    img = synth.lcg_image(0, 70, 70)
    r = restate.forward(m, img, classify=True)
    x = r["tokens"]
    pooled = x[1:].sum(axis=0) / 25.0                   # 2 registers + 25 patches summed, divided by 25
    z = restate.mul_mat(m, "classifier.weight", np.concatenate([x[0], pooled])[None, :].astype(np.float32))[0]
    z = z + m.w["classifier.bias"]
    assert np.abs(z - r["logits"]).max() < 1e-4


needs_ref = pytest.mark.skipif(not refmod.available(), reason="oracle/_ref not built on this box")


@needs_ref
def test_reference_live_matches_golden():
    R = refmod.Reference(F16, classify=False, n_threads=2, H=70, W=70)
    o = R.forward(synth.lcg_image(0, 70, 70))
    R.close()
    # same sources, possibly another ISA variant (v3/v4): within the reference's own build noise
    assert nmse(o["patch_tokens"], G["f16_feat_patch"]) < 1e-8


@needs_ref
@pytest.mark.parametrize("name,H,W,classify", [("tiny_swiglu", 70, 70, True), ("mini", 224, 224, False),
                                                ("mini", 210, 238, True)])
def test_restatement_matches_reference_live(name, H, W, classify, workdir):
    cfg = synth.CONFIGS[name]
    p = os.path.join(workdir, name + ".gguf")
    synth.write_synth_gguf(p, cfg, seed=2)
    img = synth.lcg_image(1, H, W)
    R = refmod.Reference(p, classify=classify, n_threads=4, H=H, W=W)
    o = R.forward(img)
    R.close()
    r = restate.forward(restate.RefModel(p), img, classify=classify)
    if classify:
        assert nmse(r["logits"], o["logits"]) < 1e-6
        assert int(r["probs"].argmax()) == int(o["probs"].argmax())
    else:
        assert nmse(r["patch_tokens"], o["patch_tokens"]) < 1e-6
        assert nmse(r["cls"], o["cls"]) < 1e-6


@pytest.mark.parametrize("tag", ["q4_0", "q4_1", "q5_0", "q5_1"])
def test_restatement_legacy_quant_goldens(tag):
    """The other four types of the reference's quantize tool (files written by that tool): weights dequantised as
    dequantize_row_q4_0..q5_1, activations converted to q8_0 (q4_0, q5_0) or q8_1 with its fp16 block sum (q4_1, q5_1)."""
    m = restate.RefModel(os.path.join(GOLD, f"tiny_{tag}.gguf"))
    assert m.act_type("classifier.weight") == ("q8_1" if tag.endswith("_1") else "q8_0")
    assert m.act_type("embeddings.patch_embeddings.projection.weight") == "f16"      # 4-D tensors are never quantised
    img = synth.lcg_image(0, 70, 70)
    r = restate.forward(m, img, classify=False)
    assert nmse(r["patch_tokens"], G[f"{tag}_feat_patch"]) < TOL_Q8
    r = restate.forward(m, img, classify=True)
    assert nmse(r["logits"], G[f"{tag}_cls_logits"]) < 20 * TOL_Q8
    assert int(r["probs"].argmax()) == int(G[f"{tag}_cls_probs"].argmax())
