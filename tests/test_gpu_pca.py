"""PCA colouring on the device (dino_b200_pca_rgb) against numpy's SVD: what inference.cpp:76-86 computes
with cv::PCA(DATA_AS_ROW, 3) + project + cv::normalize(0..255, NORM_MINMAX, CV_8U).  Component signs are arbitrary in any
PCA; the engine's convention (largest-magnitude loading positive) is applied to the numpy result too."""
import os

import numpy as np
import pytest

import dinov2_b200 as d
from dinov2_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def numpy_pca_rgb(x):
    xc = x.astype(np.float64) - x.astype(np.float64).mean(axis=0)
    _, _, vt = np.linalg.svd(xc, full_matrices=False)
    v = vt[:3].T                                           # [D, 3], descending singular values
    for c in range(3):
        if v[np.abs(v[:, c]).argmax(), c] < 0:
            v[:, c] = -v[:, c]
    proj = xc @ v
    lo, hi = proj.min(), proj.max()
    rgb = np.clip(np.rint((proj - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)
    return proj, rgb


@pytest.mark.parametrize("B,NP", [(1, 25), (3, 1369), (2, 200)])
def test_pca_rgb_matches_numpy_svd(B, NP):
    rng = np.random.default_rng(NP)
    with d.Engine(os.path.join(GOLD, "tiny_f16.gguf")) as e:
        D = e.hidden_size
        # features with a clear low-rank structure plus noise (as real patch tokens have): distinct top-3 variances
        basis = np.linalg.qr(rng.standard_normal((D, 5)))[0]
        coef = rng.standard_normal((B, NP, 5)) * np.array([9.0, 5.0, 2.5, 0.6, 0.3])
        x = (coef @ basis.T + 0.05 * rng.standard_normal((B, NP, D)) + rng.standard_normal((1, 1, D))).astype(np.float32)
        rgb, proj = e.pca_rgb(x, want_proj=True)
    assert rgb.shape == (B, NP, 3) and rgb.dtype == np.uint8
    for b in range(B):
        p_ref, rgb_ref = numpy_pca_rgb(x[b])
        scale = np.abs(p_ref).max()
        assert np.abs(proj[b] - p_ref).max() < 2e-3 * scale
        assert np.abs(rgb[b].astype(int) - rgb_ref.astype(int)).max() <= 1
        assert rgb[b].min() == 0 and rgb[b].max() == 255


def test_pca_rgb_on_real_engine_features():
    """End to end as inference.cpp does it: forward (features mode) then PCA of the patch tokens, device path vs host path."""
    import torch
    with d.Engine(os.path.join(GOLD, "tiny_f16.gguf")) as e:
        imgs = synth.lcg_batch(0, 2, 98, 84)
        out = e.forward(imgs)
        rgb_host = e.pca_rgb(out["patch_tokens"])
        NP = out["patch_tokens"].shape[1]
        xd = torch.from_numpy(out["patch_tokens"]).cuda()
        rgbd = torch.empty(2, NP, 3, dtype=torch.uint8, device="cuda")
        e.pca_rgb_device(xd.data_ptr(), 2, NP, rgb_ptr=rgbd.data_ptr())
        e.synchronize()
        torch.cuda.synchronize()
        assert np.array_equal(rgbd.cpu().numpy(), rgb_host)
        with pytest.raises(d.DinoB200Error):
            e.pca_rgb(out["patch_tokens"][:, :2])          # fewer than 3 patches
