"""Checkpoints written by the reference's OWN toolchain: random-init HuggingFace Dinov2(WithRegisters)ForImageClassification ->
the unmodified /root/reference/scripts/dinov2-to-gguf.py (HF state_dict -> fused qkv -> gguf, :69-166).  The files and the
reference build's outputs on them are committed fixtures (tests/golden/make_converted.py, run where the reference tree exists);
every other checkpoint in this suite comes from the repo's own writer (synth.py / gguf_io.py).

CPU: the engine's GGUF reader and the oracle restatement accept the converter's container (tensor order: state_dict first,
fused qkv appended last; gguf-py's KV / alignment conventions).  GPU: the engine loads them and matches the reference build."""
import ctypes
import os

import numpy as np
import pytest

import dinov2_b200 as d
from dinov2_b200 import gguf_io, synth
import ref as refmod
import restate
from conftest import nmse

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CG = np.load(os.path.join(GOLD, "converted.npz"))
TAGS = ("noreg", "reg2")


@pytest.mark.parametrize("tag", TAGS)
def test_converter_manifest_is_what_survey_appendix_b_says(tag):
    gg = gguf_io.read_gguf(os.path.join(GOLD, f"hf_conv_{tag}.gguf"))
    assert gg.kv["general.architecture"] == "dinov2"
    assert (gg.kv["hidden_size"], gg.kv["num_hidden_layers"], gg.kv["num_attention_heads"], gg.kv["num_classes"]) == (128, 2, 2, 10)
    assert gg.kv["num_register_tokens"] == (2 if tag == "reg2" else 0) and gg.kv["ftype"] == 1
    names = list(gg.tensors)
    assert names[-1].endswith("attention.attention.qkv.bias") and names[-2].endswith("attention.attention.qkv.weight")   # fused qkv appended last
    assert ("embeddings.register_tokens" in gg.tensors) == (tag == "reg2")
    assert gg.tensors["embeddings.patch_embeddings.projection.weight"].ne == (14, 14, 3, 128)
    assert gg.tensors["embeddings.patch_embeddings.projection.bias"].ne == (1, 1, 128, 1)


@pytest.mark.parametrize("tag", TAGS)
def test_restatement_matches_reference_on_converter_files(tag):
    m = restate.RefModel(os.path.join(GOLD, f"hf_conv_{tag}.gguf"))
    img = synth.lcg_image(0, 70, 70)
    f = restate.forward(m, img, classify=False)
    c = restate.forward(m, img, classify=True)
    assert nmse(f["patch_tokens"], CG[f"{tag}_feat_patch"]) < 5e-7      # the oracle's own two forms agree to ~1e-7 (tests/test_oracle.py)
    assert nmse(f["cls"], CG[f"{tag}_feat_cls"]) < 5e-7      # the oracle's own two forms agree to ~1e-7 (tests/test_oracle.py)
    assert nmse(c["logits"], CG[f"{tag}_cls_logits"]) < 5e-7      # the oracle's own two forms agree to ~1e-7 (tests/test_oracle.py)
    assert int(c["probs"].argmax()) == int(CG[f"{tag}_cls_probs"].argmax())


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_engine_on_converter_files_matches_reference(tag):
    path = os.path.join(GOLD, f"hf_conv_{tag}.gguf")
    imgs = synth.lcg_batch(0, 3, 70, 70)
    with d.Engine(path) as e:
        assert (e.hidden_size, e.num_hidden_layers, e.num_register_tokens, e.num_classes) == (128, 2, 2 if tag == "reg2" else 0, 10)
        assert e.label(0) == "LABEL_0"                              # HF's default id2label, written by the converter (:43, :118-120)
        out = e.forward(imgs, classify=True)
    assert nmse(out["patch_tokens"][0], CG[f"{tag}_feat_patch"]) < 1e-6
    assert np.abs(out["patch_tokens"][0] - CG[f"{tag}_feat_patch"]).max() < 5e-3
    assert nmse(out["cls"][0], CG[f"{tag}_feat_cls"]) < 1e-6
    assert nmse(out["logits"][0], CG[f"{tag}_cls_logits"]) < 1e-6
    assert nmse(out["probs"][0], CG[f"{tag}_cls_probs"]) < 1e-6
    assert int(out["probs"][0].argmax()) == int(CG[f"{tag}_cls_probs"].argmax())
    if refmod.available():                                          # and live, on another image of the batch
        R = refmod.Reference(path, classify=True, n_threads=2, H=70, W=70)
        o = R.forward(imgs[2])
        R.close()
        assert nmse(out["logits"][2], o["logits"]) < 1e-6
        assert int(out["probs"][2].argmax()) == int(o["probs"].argmax())
