"""GGUF container + synthetic-checkpoint writer (CPU)."""
import os

import numpy as np
import pytest

import dinov2_b200  # noqa: F401
from dinov2_b200 import gguf_io as G
from dinov2_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_roundtrip(workdir):
    cfg = synth.CONFIGS["tiny"]
    p = os.path.join(workdir, "rt.gguf")
    synth.write_synth_gguf(p, cfg, seed=3)
    g = G.read_gguf(p)
    assert g.kv["general.architecture"] == "dinov2"
    assert g.kv["hidden_size"] == 128 and g.kv["num_register_tokens"] == 2 and g.kv["ftype"] == 1
    assert g.kv["0"] == "class_0000" and g.kv["9"] == "class_0009"
    ts = synth.make_tensors(cfg, seed=3)
    assert list(g.tensors) == [t.name for t in ts]
    for t in ts:
        assert g.tensors[t.name].ne == t.ne and g.tensors[t.name].ggml_type == t.ggml_type
        assert np.array_equal(g.tensors[t.name].data, t.data)


def test_manifest_matches_reference_converter():
    """Names / ggml shapes / dtypes of SURVEY.md appendix B (dumped from a file the reference converter wrote)."""
    cfg = synth.CONFIGS["vits14_reg4"]
    ts = {t.name: t for t in synth.make_tensors(synth.ModelConfig("x", 384, 1, 6, num_register_tokens=4), seed=0)}
    D = 384
    assert ts["embeddings.cls_token"].ne == (D, 1, 1) and ts["embeddings.cls_token"].ggml_type == G.GGML_TYPE_F32
    assert ts["embeddings.position_embeddings"].ne == (D, 1370, 1)
    assert ts["embeddings.register_tokens"].ne == (D, 4, 1)
    assert ts["embeddings.patch_embeddings.projection.weight"].ne == (14, 14, 3, D)
    assert ts["embeddings.patch_embeddings.projection.weight"].ggml_type == G.GGML_TYPE_F16
    assert ts["embeddings.patch_embeddings.projection.bias"].ne == (1, 1, D, 1)
    assert ts["encoder.layer.0.attention.attention.qkv.weight"].ne == (D, 3 * D)
    assert ts["encoder.layer.0.mlp.fc1.weight"].ne == (D, 4 * D) and ts["encoder.layer.0.mlp.fc2.weight"].ne == (4 * D, D)
    assert ts["classifier.weight"].ne == (2 * D, 1000)
    assert list(ts)[-2:] == ["encoder.layer.0.attention.attention.qkv.weight", "encoder.layer.0.attention.attention.qkv.bias"]
    assert len(synth.make_tensors(cfg, 0)) == 177 and len(synth.make_tensors(synth.CONFIGS["vits14"], 0)) == 176
    g = synth.CONFIGS["vitg14"]
    assert (g.mlp_in, g.mlp_hidden) == (8192, 4096)


def test_q8_0_against_reference_quantize_tool():
    """tiny_q8_0.gguf was written by the reference's `quantize` binary from tiny_f16.gguf."""
    f16 = G.read_gguf(os.path.join(GOLD, "tiny_f16.gguf"))
    q8 = G.read_gguf(os.path.join(GOLD, "tiny_q8_0.gguf"))
    assert q8.kv["ftype"] == 8
    n_q = 0
    for name, t in q8.tensors.items():
        src = f16.tensors[name]
        if t.ggml_type == G.GGML_TYPE_Q8_0:
            n_q += 1
            assert name.endswith("weight") and len(t.ne) == 2          # dinov2.cpp:227-236
            w = G.to_numpy(src).astype(np.float32)
            mine = G.quantize_q8_0(w).reshape(-1, 34)
            theirs = t.data.reshape(-1, 34)
            assert np.array_equal(mine[:, :2], theirs[:, :2])            # block scales: bit-exact
            dq = np.abs(mine[:, 2:].view(np.int8).astype(int) - theirs[:, 2:].view(np.int8).astype(int))
            # the reference binary is built with -ffast-math (x*id vs x/d): rare off-by-one at .5 ties
            assert dq.max() <= 1 and (dq != 0).mean() < 1e-3
            deq = G.dequantize_q8_0(t.data, t.ne)
            assert np.abs(deq - w).max() <= np.abs(w).max() / 127 * 0.51 + 1e-7
        else:
            assert t.ggml_type == src.ggml_type and np.array_equal(t.data, src.data)
    assert n_q == 2 * 4 + 1                                              # qkv, dense, fc1, fc2 per layer + classifier


def test_lcg_image_matches_scalar_definition():
    img = synth.lcg_image(5, 4, 5)
    s = (12345 + 5) & 0xFFFFFFFF
    want = []
    for _ in range(4 * 5 * 3):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        want.append(np.float32(((s >> 8) & 0xFFFF)) / np.float32(65535.0) * np.float32(4.0) - np.float32(2.0))
    assert np.array_equal(img.reshape(-1), np.array(want, dtype=np.float32))
    assert img.min() >= -2 and img.max() <= 2


def test_reader_rejects_garbage(workdir):
    p = os.path.join(workdir, "bad.gguf")
    with open(p, "wb") as f:
        f.write(b"NOTGGUF" + b"\0" * 64)
    with pytest.raises(ValueError):
        G.read_gguf(p)


@pytest.mark.parametrize("tag,gtype,levels", [("q4_0", G.GGML_TYPE_Q4_0, 16), ("q4_1", G.GGML_TYPE_Q4_1, 16), ("q5_0", G.GGML_TYPE_Q5_0, 32),
                                               ("q5_1", G.GGML_TYPE_Q5_1, 32)])
def test_legacy_quant_files_from_reference_quantize_tool(tag, gtype, levels):
    """tiny_q4_0 .. tiny_q5_1.gguf were written by the reference's `quantize` binary from tiny_f16.gguf: our reader sizes
    the blocks correctly, and the dequantised weights sit within one quantisation step of the originals."""
    f16 = G.read_gguf(os.path.join(GOLD, "tiny_f16.gguf"))
    q = G.read_gguf(os.path.join(GOLD, f"tiny_{tag}.gguf"))
    n_quant = 0
    for name, t in q.tensors.items():
        w = np.asarray(G.to_numpy(f16.tensors[name]), dtype=np.float32)
        if t.ggml_type != gtype:
            assert t.ggml_type == f16.tensors[name].ggml_type
            assert np.array_equal(np.asarray(G.to_numpy(t), np.float32).ravel(), w.ravel())     # (the tool may drop unit dims)
            continue
        n_quant += 1
        deq = G.dequantize_legacy(t.data, t.ne, gtype)
        assert deq.shape == w.shape
        blk = w.reshape(-1, 32)
        step = (blk.max(axis=1) - blk.min(axis=1) if tag.endswith("_1") else 2 * np.abs(blk).max(axis=1)) / (levels - 1)
        err = np.abs(deq.reshape(-1, 32) - blk).max(axis=1)
        assert (err <= 1.01 * step + 1e-6).all()
        qd, m = G.dequantize_legacy(t.data, t.ne, gtype, split=True)
        assert np.allclose(qd.reshape(-1, 32) + m.reshape(-1, 1), deq.reshape(-1, 32), atol=1e-6)
    assert n_quant == 9          # qkv / output.dense / fc1 / fc2 of both layers + the classifier (dinov2.cpp do_quantize)


@pytest.mark.parametrize("tag,itype", [("q4_0", 2), ("q4_1", 3), ("q5_0", 6), ("q5_1", 7), ("q8_0", 8)])
def test_quantize_gguf_reproduces_the_reference_tool(tag, itype, tmp_path):
    """dino_b200_quantize_gguf (host-only replacement of dino_model_quantize, dinov2.cpp:354-452) against the files the
    reference's own `quantize` binary wrote from the same input: identical container (size, KVs with the re-set ftype
    last, trimmed tensor dims, offsets, alignment) and identical payload up to the handful of blocks where the reference
    build's -ffast-math moves a value across a rounding boundary (one quantisation step)."""
    out = str(tmp_path / f"mine_{tag}.gguf")
    dinov2_b200.quantize_gguf(os.path.join(GOLD, "tiny_f16.gguf"), out, itype)
    mine = np.fromfile(out, dtype=np.uint8)
    ref = np.fromfile(os.path.join(GOLD, f"tiny_{tag}.gguf"), dtype=np.uint8)
    assert mine.size == ref.size
    differing = np.nonzero(mine != ref)[0]
    assert differing.size <= 2e-4 * ref.size, differing.size
    a, b = G.read_gguf(out), G.read_gguf(os.path.join(GOLD, f"tiny_{tag}.gguf"))
    assert list(a.kv.keys()) == list(b.kv.keys()) and a.kv["ftype"] == itype
    assert list(a.tensors.keys()) == list(b.tensors.keys())
    for name in a.tensors:
        ta, tb = a.tensors[name], b.tensors[name]
        assert ta.ggml_type == tb.ggml_type and tuple(ta.ne) == tuple(tb.ne)
        wa, wb = np.asarray(G.to_numpy(ta), np.float32), np.asarray(G.to_numpy(tb), np.float32)
        if ta.ggml_type == itype:
            step = np.abs(wb).max() / (7 if itype in (2, 3) else 15 if itype in (6, 7) else 127)
            assert np.abs(wa - wb).max() <= 1.01 * step
        else:
            assert np.array_equal(wa, wb)


def test_quantize_gguf_errors(tmp_path):
    with pytest.raises(dinov2_b200.DinoB200Error):
        dinov2_b200.quantize_gguf(os.path.join(GOLD, "tiny_f16.gguf"), str(tmp_path / "x.gguf"), 12)     # a K-quant: not offered
    with pytest.raises(dinov2_b200.DinoB200Error):
        dinov2_b200.quantize_gguf(str(tmp_path / "missing.gguf"), str(tmp_path / "x.gguf"), 8)
    with pytest.raises(dinov2_b200.DinoB200Error):
        dinov2_b200.quantize_gguf(os.path.join(GOLD, "tiny_q8_0.gguf"), str(tmp_path / "x.gguf"), 2)      # already quantised
