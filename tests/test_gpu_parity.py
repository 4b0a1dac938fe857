"""End-to-end parity of the CUDA engine (through the C ABI) against the oracle: the committed golden vectors
(unmodified reference build), the numpy restatement on seeded inputs, and — where oracle/_ref travelled to this
box — the reference live.  Tolerances: f16 checkpoints NMSE <= 1e-6, max_abs <= 5e-3, top-1
equal (the reference differs from itself by NMSE ~1e-7 / max_abs ~2e-3 across build flags); q8_0 checkpoints
NMSE <= 1.5e-4 and top-1 equal (the reference quantises activations to int8; its self-noise is ~7e-5)."""
import os

import numpy as np
import pytest

import dinov2_b200 as d
from dinov2_b200 import synth
import ref as refmod
import restate
from conftest import nmse

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
G = np.load(os.path.join(GOLD, "golden.npz"))
F16 = os.path.join(GOLD, "tiny_f16.gguf")
Q8 = os.path.join(GOLD, "tiny_q8_0.gguf")
NMSE_F16, MAXABS_F16, NMSE_Q8 = 1e-6, 5e-3, 1.5e-4


@pytest.fixture(scope="module")
def tiny():
    with d.Engine(F16) as e:
        yield e


def test_loader_reads_hparams_and_labels(tiny):
    assert (tiny.hidden_size, tiny.num_hidden_layers, tiny.num_attention_heads) == (128, 2, 2)
    assert (tiny.num_register_tokens, tiny.num_classes, tiny.img_size, tiny.patch_size, tiny.ftype) == (2, 10, 70, 14, 1)
    assert tiny.label(3) == "class_0003" and tiny.label(10) is None


def test_golden_features(tiny):
    out = tiny.forward(synth.lcg_batch(0, 1, 70, 70))
    assert nmse(out["patch_tokens"][0], G["f16_feat_patch"]) < NMSE_F16
    assert nmse(out["cls"][0], G["f16_feat_cls"]) < NMSE_F16
    assert np.abs(out["patch_tokens"][0] - G["f16_feat_patch"]).max() < MAXABS_F16


def test_golden_classify(tiny):
    out = tiny.forward(synth.lcg_batch(0, 1, 70, 70), classify=True)
    assert nmse(out["logits"][0], G["f16_cls_logits"]) < NMSE_F16
    assert nmse(out["probs"][0], G["f16_cls_probs"]) < NMSE_F16
    assert int(out["probs"][0].argmax()) == int(G["f16_cls_probs"].argmax())
    assert abs(float(out["probs"][0].sum()) - 1) < 1e-5


def test_golden_non_native_grid_and_device_pos_embed(tiny):
    """98x84 -> 7x6 grid: the engine resamples the pos-embed on the device (bicubic, OpenCV convention)."""
    pos = tiny.get_pos_embed(98, 84)
    assert np.abs(pos - G["f16_nn_pos"]).max() < 2e-6
    out = tiny.forward(synth.lcg_batch(3, 1, 98, 84))
    assert nmse(out["patch_tokens"][0], G["f16_nn_patch"]) < NMSE_F16
    assert nmse(out["cls"][0], G["f16_nn_cls"]) < NMSE_F16


def test_host_pos_embed_override_matches_device_path(tiny):
    base = tiny.forward(synth.lcg_batch(3, 1, 98, 84))["patch_tokens"].copy()
    tiny.set_pos_embed(7, 6, G["f16_nn_pos"])          # what the reference's host interpolate_pos_embed returns
    over = tiny.forward(synth.lcg_batch(3, 1, 98, 84))["patch_tokens"]
    # the two embeddings differ in the last fp32 bits (device vs host bicubic); fp16 roundings of the attention
    # probabilities downstream turn that into ~1e-10 NMSE at most
    assert nmse(over, base) < 1e-8


def test_golden_q8_0():
    with d.Engine(Q8) as e:
        assert e.ftype == 8
        out = e.forward(synth.lcg_batch(0, 1, 70, 70), classify=True)
    assert nmse(out["patch_tokens"][0], G["q8_feat_patch"]) < NMSE_Q8
    assert nmse(out["logits"][0], G["q8_cls_logits"]) < 10 * NMSE_Q8       # 10 logits: a noisy statistic
    assert int(out["probs"][0].argmax()) == int(G["q8_cls_probs"].argmax())


def test_batch_is_independent_and_layouts_agree(tiny):
    imgs = synth.lcg_batch(0, 5, 70, 70)
    full = tiny.forward(imgs, classify=True)
    for i in (0, 4):
        one = tiny.forward(imgs[i:i + 1], classify=True)
        assert np.array_equal(one["patch_tokens"][0], full["patch_tokens"][i])      # bit-exact: no cross-image leakage
        assert np.array_equal(one["probs"][0], full["probs"][i])
    planar = np.ascontiguousarray(imgs[:, :, :, ::-1].transpose(0, 3, 1, 2))         # BGR-HWC -> RGB planar
    alt = tiny.forward(planar, classify=True, layout=d.LAYOUT_RGB_PLANAR)
    assert np.array_equal(alt["patch_tokens"], full["patch_tokens"])
    again = tiny.forward(imgs, classify=True)
    assert np.array_equal(again["logits"], full["logits"])                           # deterministic


def test_errors(tiny):
    with pytest.raises(d.DinoB200Error):
        tiny.forward(np.zeros((1, 71, 70, 3), np.float32))          # not a patch multiple
    with pytest.raises(d.DinoB200Error):
        d.Engine("/nonexistent/model.gguf")
    bad = os.path.join(GOLD, "make_golden.py")
    with pytest.raises(d.DinoB200Error):
        d.Engine(bad)                                               # not a GGUF


CASES = [
    ("tiny_noreg", None, 70, 70, 3, True),
    ("tiny_swiglu", None, 70, 70, 2, True),          # 40 layers -> the reference's SwiGLU branch
    ("tiny_swiglu", "q8_0", 70, 70, 2, False),
    ("mini", None, 224, 224, 2, False),
    ("mini", None, 210, 238, 2, True),               # ragged, non-square, non-native grid
    ("mini", "q8_0", 224, 224, 2, True),
]


@pytest.mark.parametrize("name,quant,H,W,B,classify", CASES)
def test_seeded_models_vs_restatement(name, quant, H, W, B, classify, workdir):
    cfg = synth.CONFIGS[name]
    p = os.path.join(workdir, f"{name}_{quant}.gguf")
    synth.write_synth_gguf(p, cfg, seed=5, quant=quant)
    imgs = synth.lcg_batch(10, B, H, W)
    with d.Engine(p) as e:
        out = e.forward(imgs, classify=classify)
    m = restate.RefModel(p)
    tol = NMSE_Q8 if quant else NMSE_F16
    for i in range(B):
        r = restate.forward(m, imgs[i], classify=classify)
        assert nmse(out["patch_tokens"][i], r["patch_tokens"]) < tol
        assert nmse(out["cls"][i], r["cls"]) < tol
        if classify:
            assert int(out["probs"][i].argmax()) == int(r["probs"].argmax())
            assert nmse(out["logits"][i], r["logits"]) < (20 * tol if quant else tol)


@pytest.mark.parametrize("name,classify", [("vits14", False), ("vits14_reg4", True)])
def test_vits14_518_vs_oracle(name, classify, workdir):
    """BASELINE.json configs[0]/[1] shape: ViT-S/14 (+4 registers), 518x518, N = 1370 / 1374 tokens."""
    cfg = synth.CONFIGS[name]
    p = os.path.join(workdir, name + ".gguf")
    synth.write_synth_gguf(p, cfg, seed=0)
    imgs = synth.lcg_batch(0, 3, 518, 518)
    with d.Engine(p) as e:
        out = e.forward(imgs, classify=classify)
    m = restate.RefModel(p)
    r = restate.forward(m, imgs[2], classify=classify)
    assert nmse(out["patch_tokens"][2], r["patch_tokens"]) < NMSE_F16
    assert np.abs(out["patch_tokens"][2] - r["patch_tokens"]).max() < MAXABS_F16
    inside = np.abs(out["patch_tokens"][2] - r["patch_tokens"]) <= 1e-3 + 1e-3 * np.abs(r["patch_tokens"])
    assert inside.mean() > 0.999
    if classify:
        assert int(out["probs"][2].argmax()) == int(r["probs"].argmax())
        assert nmse(out["logits"][2], r["logits"]) < NMSE_F16
    if refmod.available():
        R = refmod.Reference(p, classify=classify, H=518, W=518)
        o = R.forward(imgs[0])
        R.close()
        if classify:
            assert int(out["probs"][0].argmax()) == int(o["probs"].argmax())
            assert nmse(out["logits"][0], o["logits"]) < NMSE_F16
        else:
            assert nmse(out["patch_tokens"][0], o["patch_tokens"]) < NMSE_F16
            assert np.abs(out["patch_tokens"][0] - o["patch_tokens"]).max() < MAXABS_F16


def test_full_size_properties_vitl14(workdir):
    """BASELINE.json configs[3] at full width (ViT-L/14, 518^2) where the CPU oracle is too slow to run in a test:
    size-independent properties — batch permutation equivariance (bit-exact), softmax normalisation, LayerNorm
    statistics of the output tokens, determinism."""
    cfg = synth.CONFIGS["vitl14"]
    p = os.path.join(workdir, "vitl14.gguf")
    synth.write_synth_gguf(p, cfg, seed=0)
    imgs = synth.lcg_batch(0, 4, 518, 518)
    with d.Engine(p) as e:
        a = e.forward(imgs, classify=True)
        b = e.forward(np.ascontiguousarray(imgs[::-1]), classify=True)
        g = np.frombuffer(open(p, "rb").read()[:0], dtype=np.uint8)   # noqa: F841 (keep file alive)
    assert np.isfinite(a["patch_tokens"]).all()
    assert np.array_equal(a["patch_tokens"], b["patch_tokens"][::-1])
    assert np.array_equal(a["probs"], b["probs"][::-1])
    assert np.abs(a["probs"].sum(axis=1) - 1).max() < 1e-5
    assert np.array_equal(a["probs"].argmax(axis=1), a["logits"].argmax(axis=1))
    # final LayerNorm: (y - beta) / gamma has zero mean / unit variance per token
    from dinov2_b200 import gguf_io
    gg = gguf_io.read_gguf(p)
    gam = gguf_io.to_numpy(gg.tensors["layernorm.weight"])
    bet = gguf_io.to_numpy(gg.tensors["layernorm.bias"])
    z = (a["patch_tokens"][0] - bet) / gam
    assert np.abs(z.mean(axis=1)).max() < 1e-3 and np.abs(z.var(axis=1) - 1).max() < 1e-2
    # and, where the reference build travelled to this box, the reference itself on one image of the headline model
    # (24 layers of accumulated rounding: same tolerances as every other f16 case)
    if refmod.available():
        R = refmod.Reference(p, classify=True, H=518, W=518)
        o = R.forward(imgs[0])
        R.close()
        assert int(a["probs"][0].argmax()) == int(o["probs"].argmax())
        assert nmse(a["logits"][0], o["logits"]) < NMSE_F16
        assert nmse(a["probs"][0], o["probs"]) < NMSE_F16


def test_forward_is_bit_reproducible_vitl14(workdir):
    """Three forward passes over the same 16 images give the same bits.  Images 64.. of the LCG stream are the ones on which the
    round-2 race in the attention kernel's growth path showed (image 67 differed by ~1e-3 in cls / 2e-2 in patch tokens from
    run to run; tests/test_gpu_kernels.py::test_attention_sparse_growth_is_exact_and_reproducible is the kernel-level case)."""
    cfg = synth.CONFIGS["vitl14"]
    p = os.path.join(workdir, "vitl14.gguf")
    if not os.path.exists(p):
        synth.write_synth_gguf(p, cfg, seed=0)
    imgs = synth.lcg_batch(64, 16, 518, 518)
    with d.Engine(p) as e:
        runs = [e.forward(imgs, classify=True) for _ in range(3)]
    for r in runs[1:]:
        for k in runs[0]:
            assert np.array_equal(r[k], runs[0][k]), k


@pytest.mark.parametrize("tag", ["q4_0", "q4_1", "q5_0", "q5_1"])
def test_golden_legacy_quant_types(tag):
    """Checkpoints written by the reference's quantize tool in its other four formats: weights are dequantised once at load
    (same values the reference's dot products see), activations are not quantised to int8 -> same tolerance as q8_0."""
    with d.Engine(os.path.join(GOLD, f"tiny_{tag}.gguf")) as e:
        out = e.forward(synth.lcg_batch(0, 1, 70, 70), classify=True)
    assert nmse(out["patch_tokens"][0], G[f"{tag}_feat_patch"]) < NMSE_Q8
    assert nmse(out["cls"][0], G[f"{tag}_feat_cls"]) < NMSE_Q8
    assert nmse(out["logits"][0], G[f"{tag}_cls_logits"]) < 10 * NMSE_Q8
    assert int(out["probs"][0].argmax()) == int(G[f"{tag}_cls_probs"].argmax())


def test_pipelined_submit_wait_matches_synchronous_forward(workdir):
    """dino_b200_submit / dino_b200_wait (two batches in flight, uploads under the previous forward) return exactly what
    the synchronous dino_b200_forward returns for each batch, in order."""
    import torch
    path = os.path.join(workdir, "tiny_pipe.gguf")
    cfg = synth.CONFIGS["tiny"]
    synth.write_synth_gguf(path, cfg, seed=3)
    batches = [torch.from_numpy(synth.lcg_batch(10 * k, 3, 70, 70)).pin_memory().numpy() for k in range(5)]
    with d.Engine(path) as eng:
        want = [eng.forward(b, classify=True) for b in batches]
        outs = [{k: np.empty_like(v) for k, v in want[0].items()} for _ in range(len(batches))]
        eng.submit(batches[0], outs[0], classify=True)
        for k in range(1, len(batches)):
            eng.submit(batches[k], outs[k], classify=True)
            eng.wait()
        eng.wait()
        with pytest.raises(d.DinoB200Error):
            eng.wait()                                   # nothing in flight
        for k in range(len(batches)):
            for name in want[k]:
                assert np.array_equal(outs[k][name], want[k][name]), (k, name)


_FUSE_LN_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import dinov2_b200 as d
from dinov2_b200 import synth
imgs = synth.lcg_batch(0, int(sys.argv[4]), 518, 518)
with d.Engine(sys.argv[2]) as e:
    out = e.forward(imgs, classify=True)
np.savez(sys.argv[3], **out)
"""


@pytest.mark.parametrize("name,B", [("vits14_reg4", 3), ("vitb14", 2)])
def test_forward_with_fused_layernorm_is_bit_identical(name, B, workdir):
    """DINO_B200_FUSE_LN=1 runs norm2 / the next block's norm1 inside the residual GEMMs (EPI_RESID_LN_F32, LayerNorm worker
    warps).  The option is off by default but must stay exact: same X bits (same TMA reduce-adds), same row
    arithmetic as the stand-alone kernel -> every output of the forward pass is bit-identical.  The switch is read once per
    process, so both variants run in children."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(workdir, f"fuse_{name}.gguf")
    synth.write_synth_gguf(path, synth.CONFIGS[name], seed=5)
    outs = {}
    for flag in ("0", "1"):
        npz = os.path.join(workdir, f"fuse_{name}_{flag}.npz")
        env = dict(os.environ, DINO_B200_FUSE_LN=flag)
        subprocess.run([sys.executable, "-c", _FUSE_LN_CHILD, root, path, npz, str(B)], check=True, env=env, timeout=300)
        outs[flag] = dict(np.load(npz))
    for k in outs["0"]:
        assert np.array_equal(outs["0"][k], outs["1"][k]), k
    assert np.isfinite(outs["1"]["logits"]).all()


# ---------------------------------------------------------------------------------------------------------------------
# Full-size parity on every BASELINE.json config, against the reference itself (oracle/_ref: the unmodified reference
# built from /root/reference, running live on this box's host cores).  The engine runs the WHOLE batch of the config;
# the reference (batch 1 only, dinov2.cpp:630) re-computes the first and the last image of that batch.
# ---------------------------------------------------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(not refmod.available(), reason="oracle/_ref (the reference build) did not travel to this box")


def _report(tag, got, want):
    e = np.abs(got - want)
    inside = float((e <= 1e-3 + 1e-3 * np.abs(want)).mean())
    print(f"[parity] {tag}: nmse {nmse(got, want):.3e} max_abs {float(e.max()):.3e} inside(1e-3+1e-3|ref|) {inside:.5f}")
    return inside


@needs_ref
def test_vitb14_b64_features_vs_reference(workdir):
    """BASELINE.json configs[2]: ViT-B/14, batch 64, feature extraction — patch tokens and cls of images 0 and 63
    (image 63 sits in the last wave of every kernel's tile schedule).  The reference's own noise floor on this checkpoint
    (numpy restatement vs reference build, image 0): NMSE 2.2e-7, max_abs 2.6e-3, 99.49 % of the elements inside
    1e-3 + 1e-3 |ref| — the 99.9 % of SURVEY.md appendix D is a ViT-S figure, so the fraction is gated at 99 % here."""
    p = os.path.join(workdir, "vitb14.gguf")
    synth.write_synth_gguf(p, synth.CONFIGS["vitb14"], seed=0)
    imgs = synth.lcg_batch(100, 64, 518, 518)
    with d.Engine(p) as e:
        out = e.forward(imgs, classify=False)
    assert np.isfinite(out["patch_tokens"]).all()
    R = refmod.Reference(p, classify=False, H=518, W=518)
    for i in (0, 63):
        o = R.forward(imgs[i])
        inside = _report(f"vitb14 b64 image {i} patch", out["patch_tokens"][i], o["patch_tokens"])
        assert nmse(out["patch_tokens"][i], o["patch_tokens"]) < NMSE_F16
        assert np.abs(out["patch_tokens"][i] - o["patch_tokens"]).max() < MAXABS_F16
        assert inside > 0.99
        assert nmse(out["cls"][i], o["cls"]) < NMSE_F16
    R.close()


@needs_ref
def test_vitl14_b64_classify_vs_reference(workdir):
    """BASELINE.json configs[3], the headline: ViT-L/14, batch 64, classify — logits, probabilities, top-1, cls and patch
    tokens of images 0 and 63 of the full batch."""
    p = os.path.join(workdir, "vitl14.gguf")
    if not os.path.exists(p):
        synth.write_synth_gguf(p, synth.CONFIGS["vitl14"], seed=0)
    imgs = synth.lcg_batch(200, 64, 518, 518)
    with d.Engine(p) as e:
        out = e.forward(imgs, classify=True)
    assert np.isfinite(out["patch_tokens"]).all() and np.isfinite(out["logits"]).all()
    # Two reference graphs: the classify graph for logits / probabilities, the features graph for the tokens.  (Reading the
    # final-LayerNorm tokens back out of the reference's CLASSIFY graph is not reliable: they are an intermediate there and
    # ggml's graph allocator re-uses the first row for the pooled vector — restatement and engine agree with each other on
    # that row and disagree with the harness by up to 3.6.)
    Rc = refmod.Reference(p, classify=True, H=518, W=518)
    Rf = refmod.Reference(p, classify=False, H=518, W=518)
    for i in (0, 63):
        o = Rc.forward(imgs[i])
        f = Rf.forward(imgs[i])
        _report(f"vitl14 b64 image {i} logits", out["logits"][i], o["logits"])
        inside = _report(f"vitl14 b64 image {i} patch", out["patch_tokens"][i], f["patch_tokens"])
        assert int(out["probs"][i].argmax()) == int(o["probs"].argmax())
        assert nmse(out["logits"][i], o["logits"]) < NMSE_F16
        assert nmse(out["probs"][i], o["probs"]) < NMSE_F16
        assert nmse(out["patch_tokens"][i], f["patch_tokens"]) < NMSE_F16
        assert np.abs(out["patch_tokens"][i] - f["patch_tokens"]).max() < MAXABS_F16
        assert inside > 0.98     # the reference's own floor at this depth (restatement vs build, image 0): 98.8 %, NMSE 2.4e-7
        assert nmse(out["cls"][i], f["cls"]) < NMSE_F16
    Rc.close()
    Rf.close()


@needs_ref
def test_vitg14_q8_0_b16_classify_vs_reference(workdir):
    """BASELINE.json configs[4]: ViT-g/14 (1536 wide, 40 layers, SwiGLU) from a q8_0 checkpoint, batch 16 — the reference
    runs its int8 x int8 path (ggml-cpu-quants.c:3597), the engine dequantises the weights once and keeps fp16 activations.

    Tolerance: at this depth the int8 re-quantisation of every activation row amplifies last-bit differences, and the
    UNMODIFIED reference no longer agrees with itself: its x86-64-v3 and x86-64-v4 builds (same sources, same flags except
    the ISA level) differ by NMSE 8.7e-3 on the logits and 7.9e-3 on the patch tokens of image 0 (measured; the faithful
    numpy restatement differs from either by the same amount), with equal top-1.  The 1.5e-4 bound of SURVEY.md appendix D is
    a 12-layer ViT-S figure.  So the gate here is the reference's own spread: the engine must be no further from the reference
    than TWICE the distance between the reference's two builds (measured live when both run on this host, else the recorded
    8.7e-3), and top-1 must be equal.  Measured engine distance: 1.1e-2 (logits), 9.4e-3 (patch tokens)."""
    p = os.path.join(workdir, "vitg14_q8_0.gguf")
    synth.write_synth_gguf(p, synth.CONFIGS["vitg14"], seed=0, quant="q8_0")
    imgs = synth.lcg_batch(300, 16, 518, 518)
    with d.Engine(p) as e:
        assert e.ftype == 8
        out = e.forward(imgs, classify=True)
    assert np.isfinite(out["logits"]).all()
    floor_logits, floor_patch = 8.7e-3, 7.9e-3
    if len(refmod.builds()) >= 2:
        a = refmod.forward_in_subprocess("v3", p, imgs[0], True)
        b = refmod.forward_in_subprocess("v4", p, imgs[0], True)
        floor_logits, floor_patch = nmse(a["logits"], b["logits"]), nmse(a["patch_tokens"][1:], b["patch_tokens"][1:])
        assert int(a["probs"].argmax()) == int(b["probs"].argmax())
        print(f"[parity] vitg14 q8_0: reference v3 build vs v4 build: logits nmse {floor_logits:.3e}, patch nmse {floor_patch:.3e}")
    R = refmod.Reference(p, classify=True, H=518, W=518)
    for i in (0, 15):
        o = R.forward(imgs[i])
        _report(f"vitg14 q8_0 b16 image {i} logits", out["logits"][i], o["logits"])
        # (tokens read from the classify graph: row 0 is recycled by ggml's allocator, see the ViT-L test)
        _report(f"vitg14 q8_0 b16 image {i} patch", out["patch_tokens"][i][1:], o["patch_tokens"][1:])
        assert int(out["probs"][i].argmax()) == int(o["probs"].argmax())
        assert nmse(out["patch_tokens"][i][1:], o["patch_tokens"][1:]) < max(NMSE_Q8, 2 * floor_patch)
        assert nmse(out["logits"][i], o["logits"]) < max(NMSE_Q8, 2 * floor_logits)
    R.close()


@pytest.mark.parametrize("name,classify", [("vits14", False), ("vits14", True)])
def test_outlier_weights_vits14(name, classify, workdir):
    """"Massive activation" statistics (synth.make_tensors(outliers=True): LayerNorm gains 8-12x larger in four channels and
    LayerScale factors 8-12x larger in four other channels of every block; outputs reach |y| ~ 14 against a mean of 0.67).
    This stresses exactly what the engine keeps narrower than the reference — fp16 LN output / QKV / MLP hidden buffers and
    the fp16 operands of Q K^T and P V (reference: f32, dinov2.cpp:527-543).  NMSE and top-1 gates as for the benign
    weights; the reference's own noise floor on this checkpoint (numpy restatement vs reference build) is NMSE 2e-7,
    max_abs 1.8e-2, so max_abs is reported, not gated.  (With register tokens, or with both factors in the same channels,
    the reference stops agreeing with itself — restatement vs build NMSE 1e-4 .. 4e-2 — and no parity statement is possible.)"""
    p = os.path.join(workdir, name + "_outliers.gguf")
    synth.write_synth_gguf(p, synth.CONFIGS[name], seed=2, outliers=True)
    imgs = synth.lcg_batch(40, 2, 518, 518)
    with d.Engine(p) as e:
        out = e.forward(imgs, classify=classify)
    assert np.isfinite(out["patch_tokens"]).all()
    m = restate.RefModel(p)
    r = restate.forward(m, imgs[1], classify=classify)
    print(f"[parity] {name} outliers: |ref| mean {float(np.abs(r['patch_tokens']).mean()):.3f} max {float(np.abs(r['patch_tokens']).max()):.2f}")
    _report(f"{name} outliers vs restatement", out["patch_tokens"][1], r["patch_tokens"])
    assert nmse(out["patch_tokens"][1], r["patch_tokens"]) < NMSE_F16
    assert nmse(out["cls"][1], r["cls"]) < NMSE_F16
    if classify:
        assert int(out["probs"][1].argmax()) == int(r["probs"].argmax())
        assert nmse(out["logits"][1], r["logits"]) < NMSE_F16
    if refmod.available():
        R = refmod.Reference(p, classify=classify, H=518, W=518)
        o = R.forward(imgs[0])
        R.close()
        if not classify:               # (tokens of the classify graph are recycled by ggml's allocator: see the ViT-L test)
            inside = _report(f"{name} outliers vs reference", out["patch_tokens"][0], o["patch_tokens"])
            assert nmse(out["patch_tokens"][0], o["patch_tokens"]) < NMSE_F16
            assert inside > 0.99
        if classify:
            assert int(out["probs"][0].argmax()) == int(o["probs"].argmax())
            assert nmse(out["logits"][0], o["logits"]) < NMSE_F16


def test_flash_attn_compat_reproduces_the_phantom_keys(tiny):
    """-fa (DINO_B200_FLASH_ATTN_COMPAT): the reference's flash path zero-pads the tokens to a multiple of 32 and runs
    ggml_flash_attn_ext without a mask (dinov2.cpp:499-525), so every query also attends to the padding keys.  Goldens from the
    reference build run with enable_flash_attn (tests/golden/make_golden_fa.py): 28 tokens -> 4 phantom keys, 45 -> 19.  The
    engine reproduces that semantic difference; what remains is the fp16 running accumulator of ggml's CPU flash kernel
    (ops.cpp:6858-7075), which the engine (fp32 accumulation) deliberately does not imitate."""
    FA = np.load(os.path.join(GOLD, "golden_fa.npz"))
    for img, H, W, key, base in ((0, 70, 70, "fa_feat", "f16_feat"), (3, 98, 84, "fa_nn", "f16_nn")):
        x = synth.lcg_batch(img, 1, H, W)
        plain = tiny.forward(x)
        fa = tiny.forward(x, flash_attn_compat=True)
        d_plain = nmse(plain["patch_tokens"][0], FA[key + "_patch"])
        d_fa = nmse(fa["patch_tokens"][0], FA[key + "_patch"])
        print(f"[parity] -fa {H}x{W}: exact attention vs reference -fa {d_plain:.3e}; compat mode vs reference -fa {d_fa:.3e}; "
              f"reference default vs reference -fa {nmse(G[base + '_patch'], FA[key + '_patch']):.3e}")
        assert d_fa < NMSE_F16                  # measured 2.7e-9 / 2.2e-9: the phantom keys ARE the difference at this size
        assert d_fa < d_plain / 1000            # (exact attention is 3.7e-5 / 2.3e-4 away from the reference's -fa output)
        assert nmse(fa["cls"][0], FA[key + "_cls"]) < NMSE_F16
        # and the flag changes nothing when the token count is already a multiple of 32 ... (no such grid at 14 px patches
        # below 32 tokens with 2 registers; covered by the kernel test below)
    c = tiny.forward(synth.lcg_batch(0, 1, 70, 70), classify=True, flash_attn_compat=True)
    assert int(c["probs"][0].argmax()) == int(FA["fa_cls_probs"].argmax())
    assert nmse(c["logits"][0], FA["fa_cls_logits"]) < NMSE_F16
    again = tiny.forward(synth.lcg_batch(0, 1, 70, 70))
    assert nmse(again["patch_tokens"][0], G["f16_feat_patch"]) < NMSE_F16     # the default path is untouched by earlier -fa calls
