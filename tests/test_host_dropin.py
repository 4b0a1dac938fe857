"""The C++ drop-in for the reference's dinov2.h API (dinov2.cpp_b200/host/): the reference's UNMODIFIED
inference.cpp, built against it by host/Makefile, must link, fail loudly without a GPU, and on a B200 print the
same top-k as the reference's own CPU path run on the same preprocessed image."""
import os
import re
import subprocess

import numpy as np
import pytest

import dinov2_b200 as d
from dinov2_b200 import synth
import ref as refmod

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "dinov2.cpp_b200", "host", "_build")
APP = os.path.join(BUILD, "inference_b200")
GOLD = os.path.join(ROOT, "tests", "golden")

needs_app = pytest.mark.skipif(not os.path.exists(APP), reason="host drop-in not built (needs the reference tree for its headers/apps)")


def _write_ppm(path, img_bgr):
    h, w, _ = img_bgr.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(img_bgr[:, :, ::-1]).tobytes())


@needs_app
def test_reference_apps_link_against_the_dropin():
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(BUILD, "inference_b200")], capture_output=True, text=True).stdout
    assert "dino_model_load" in out and "dino_predict" in out and "ggml_backend_free" in out
    # realtime.cpp: the shim's VideoCapture never opens, so the optimiser drops the loop; parse + link is what is proven
    out = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(BUILD, "realtime_b200")], capture_output=True, text=True).stdout
    assert "dino_params_parse" in out
    lib = subprocess.run(["nm", "-D", "--defined-only", os.path.join(BUILD, "libdinov2_host.so")], capture_output=True, text=True).stdout
    for sym in ("dino_predict", "dino_model_load", "dino_preprocess", "dino_classify_preprocess", "interpolate_pos_embed",
                "dino_params_parse", "ggml_time_ms", "ggml_gallocr_new", "ggml_free", "ggml_backend_synchronize"):
        assert sym in lib, sym
    # the host layer links the engine, never ggml or the oracle
    ldd = subprocess.run(["ldd", os.path.join(BUILD, "libdinov2_host.so")], capture_output=True, text=True).stdout
    assert "libdinov2_b200.so" in ldd and "ggml" not in ldd and "dino_ref" not in ldd


@needs_app
def test_app_fails_loudly_without_gpu(tmp_path):
    if d.device_count() > 0:
        pytest.skip("a B200 is present")
    img = tmp_path / "in.ppm"
    _write_ppm(str(img), np.zeros((64, 64, 3), np.uint8))
    r = subprocess.run([APP, "-m", os.path.join(GOLD, "tiny_f16.gguf"), "-i", str(img), "-c"], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and "failed to load model" in r.stderr


@needs_app
@pytest.mark.gpu
def test_unmodified_inference_app_classify_matches_reference(tmp_path):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(260, 300, 3), dtype=np.uint8)
    ppm = tmp_path / "in.ppm"
    _write_ppm(str(ppm), img)
    gguf = os.path.join(GOLD, "tiny_f16.gguf")
    r = subprocess.run([APP, "-m", gguf, "-i", str(ppm), "-c", "-k", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = re.findall(r"^ > (\S+) : ([0-9.]+)$", r.stdout, flags=re.M)
    assert len(got) == 3
    assert "graph computation took" in r.stderr                      # the line scripts/benchmark.sh scrapes
    if refmod.available():
        R = refmod.Reference(gguf, classify=True, H=224, W=224)
        pre = R.preprocess(img, classify=True)                       # the reference's own dino_classify_preprocess
        o = R.forward(pre)
        R.close()
        order = np.argsort(-o["probs"], kind="stable")[:3]
        assert [g[0] for g in got] == [f"class_{i:04d}" for i in order]
        for (lbl, p), i in zip(got, order):
            assert abs(float(p) - float(o["probs"][i])) <= 0.011     # printed with %.2f


@needs_app
@pytest.mark.gpu
def test_unmodified_inference_app_features_writes_pca_image(tmp_path):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(100, 130, 3), dtype=np.uint8)
    ppm = tmp_path / "in.ppm"
    _write_ppm(str(ppm), img)
    r = subprocess.run([APP, "-m", os.path.join(GOLD, "tiny_f16.gguf"), "-i", str(ppm)], capture_output=True, text=True,
                       timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr
    out = tmp_path / "pca_visual.jpg"
    assert out.exists() and out.stat().st_size > 500
    assert "preprocessed image (112 x 140)" in r.stderr               # dino_preprocess rounds UP to the next patch multiple


BATCH = os.path.join(BUILD, "batch_check")


@pytest.mark.skipif(not os.path.exists(BATCH), reason="host drop-in not built")
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["f", "c"])
def test_dino_predict_batch_equals_one_by_one(mode):
    """dino_predict_batch (host/dinov2_b200_batch.h): three images in one forward pass give exactly what three dino_predict
    calls give (patch tokens bit-identical / same top-k ids)."""
    r = subprocess.run([BATCH, os.path.join(GOLD, "tiny_f16.gguf"), mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "3 images, max difference 0" in r.stderr


@pytest.mark.skipif(not os.path.exists(BATCH), reason="host drop-in not built")
@pytest.mark.gpu
@pytest.mark.parametrize("tag,type_id,block", [("f16", 1, 2), ("q4_0", 2, 18), ("q4_1", 3, 20), ("q5_0", 6, 22), ("q5_1", 7, 24), ("q8_0", 8, 34)])
def test_model_tensors_report_ggml_strides(tag, type_id, block):
    """dino_model::tensors (dinov2.h:54) of the drop-in: nb[0] is the ggml type size of each format (round-1 advisor finding:
    it was 34 for every quantised type), nb[1] the bytes of one row."""
    r = subprocess.run([BATCH, os.path.join(GOLD, f"tiny_{tag}.gguf"), "f"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    m = re.search(r"fc1.weight type (\d+) nb0=(\d+) nb1=(\d+) ne0=(\d+)", r.stderr)
    assert m, r.stderr
    t, nb0, nb1, ne0 = (int(x) for x in m.groups())
    assert (t, nb0) == (type_id, block)
    assert nb1 == (ne0 * 2 if tag == "f16" else ne0 // 32 * block)


@needs_app
def test_dropin_defines_the_whole_reference_surface():
    """Every function the reference's dinov2.h declares is defined by the drop-in (link errors otherwise): including
    get_val_u32 / get_val_str (dinov2.h:20-23), which round 1 left undefined."""
    lib = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(BUILD, "libdinov2_host.so")], capture_output=True, text=True).stdout
    for sym in ("get_val_u32", "get_val_str", "do_quantize", "dino_model_quantize", "print_usage", "print_t_f32", "build_graph",
                "forward_features", "forward_head", "dino_predict_batch"):
        assert sym in lib, sym
