"""The header-only OpenCV stand-in (dinov2.cpp_b200/host/cvshim/) against REAL OpenCV.

The reference build in oracle/_ref compiles the unmodified dinov2.cpp against the shim, so the reference's own
dino_preprocess / dino_classify_preprocess / interpolate_pos_embed (dinov2.cpp:106-225) run the shim's cv::resize(INTER_CUBIC),
convertTo and arithmetic.  Golden vectors: tests/golden/preprocess_cv2.npz, written by Python cv2 4.13 (the real library) with
tests/golden/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest

import ref as refmod

GOLD = os.path.join(os.path.dirname(__file__), "golden")
F16 = os.path.join(GOLD, "tiny_f16.gguf")
PG = np.load(os.path.join(GOLD, "preprocess_cv2.npz"))
G = np.load(os.path.join(GOLD, "golden.npz"))

pytestmark = pytest.mark.skipif(not refmod.available(), reason="oracle/_ref (reference + shim build) is not present")


@pytest.fixture(scope="module")
def R():
    r = refmod.Reference(F16, classify=False, n_threads=2, H=70, W=70)
    yield r
    r.close()


def test_shim_resize_cubic_in_dino_preprocess_matches_opencv(R):
    out = R.preprocess(PG["img"], classify=False)
    assert out.shape == PG["feat"].shape == (70, 84, 3)                  # 61 x 83 -> next patch multiple (dinov2.cpp:140-141)
    assert np.abs(out - PG["feat"]).max() < 2e-5


def test_shim_in_dino_classify_preprocess_matches_opencv(R):
    out = R.preprocess(PG["img"], classify=True)                         # squash to 256 x 256, centre crop 224 (dinov2.cpp:106-132)
    assert out.shape == (224, 224, 3)
    assert np.abs(out[:32, :32] - PG["cls_corner"]).max() < 2e-5
    assert abs(out.astype(np.float64).sum() - float(PG["cls_sum"])) < 0.05
    assert abs(np.abs(out.astype(np.float64)).sum() - float(PG["cls_abs"])) < 0.05


def test_shim_resize_in_interpolate_pos_embed_matches_cv2_when_available(R):
    """The per-channel cv::resize of interpolate_pos_embed (dinov2.cpp:196-215) through the shim equals real cv2.resize on
    the same 5 x 5 grid (skipped when the cv2 wheel is not importable; the committed golden covers it otherwise)."""
    pos = R.interpolate_pos_embed(98, 84)
    assert np.abs(pos - G["f16_nn_pos"]).max() < 2e-6                    # committed golden of the same call
    cv2 = pytest.importorskip("cv2")
    from dinov2_b200 import gguf_io
    gg = gguf_io.read_gguf(F16)
    base = gguf_io.to_numpy(gg.tensors["embeddings.position_embeddings"]).reshape(-1, R.hidden_size)
    M = R.img_size // R.patch_size
    grid = base[1:].reshape(M, M, -1)
    want = np.stack([cv2.resize(np.ascontiguousarray(grid[:, :, c]), (6, 7), interpolation=cv2.INTER_CUBIC) for c in range(grid.shape[2])], axis=-1)
    assert np.abs(pos[1:].reshape(7, 6, -1) - want).max() < 2e-6
    assert np.array_equal(pos[0], base[0])
