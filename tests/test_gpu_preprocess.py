"""Device preprocessing (dino_preprocess / dino_classify_preprocess, reference dinov2.cpp:106-156)
against (a) golden vectors from real OpenCV (Python cv2, tests/golden/preprocess_cv2.npz) and (b) the reference's own
functions compiled in oracle/_ref; and the u8 -> result path against preprocess + forward done separately."""
import os

import numpy as np
import pytest

import dinov2_b200 as d
import ref as refmod

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
F16 = os.path.join(GOLD, "tiny_f16.gguf")
PG = np.load(os.path.join(GOLD, "preprocess_cv2.npz"))
# float32 bicubic with a different association of the same 16 taps + (v-mean)/std vs OpenCV's fused multiply: ~1e-6 abs
TOL = 2e-5


@pytest.fixture(scope="module")
def eng():
    with d.Engine(F16) as e:
        yield e


def test_preprocess_features_matches_opencv(eng):
    out = eng.preprocess(PG["img"][None], classify=False)[0]
    assert out.shape == PG["feat"].shape == (70, 84, 3)            # 61x83 -> next patch multiple
    assert np.abs(out - PG["feat"]).max() < TOL


def test_preprocess_classify_matches_opencv(eng):
    out = eng.preprocess(PG["img"][None], classify=True)[0]
    assert out.shape == (224, 224, 3)
    assert np.abs(out[:32, :32] - PG["cls_corner"]).max() < TOL
    assert abs(out.astype(np.float64).sum() - float(PG["cls_sum"])) < 0.05
    assert abs(np.abs(out.astype(np.float64)).sum() - float(PG["cls_abs"])) < 0.05


def test_already_patch_multiple_still_rounds_up(eng):
    """reference quirk: 518 -> 532 (dinov2.cpp:140-141)"""
    img = np.random.default_rng(0).integers(0, 256, (2, 28, 42, 3), dtype=np.uint8)
    assert eng.preprocess(img).shape == (2, 42, 56, 3)


@pytest.mark.skipif(not refmod.available(), reason="oracle/_ref not built on this box")
@pytest.mark.parametrize("classify", [False, True])
def test_preprocess_matches_reference_functions(eng, classify):
    img = np.random.default_rng(3).integers(0, 256, (100, 130, 3), dtype=np.uint8)
    R = refmod.Reference(F16, classify=classify, n_threads=2, H=224, W=224)
    want = R.preprocess(img, classify=classify)
    R.close()
    got = eng.preprocess(img[None], classify=classify)[0]
    assert got.shape == want.shape
    assert np.abs(got - want).max() < TOL


@pytest.mark.parametrize("classify", [False, True])
def test_forward_u8_equals_preprocess_then_forward(eng, classify):
    frames = np.random.default_rng(4).integers(0, 256, (3, 90, 75, 3), dtype=np.uint8)
    pre = eng.preprocess(frames, classify=classify)
    a = eng.forward(pre, classify=classify)
    b = eng.forward_u8(frames, classify=classify)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
