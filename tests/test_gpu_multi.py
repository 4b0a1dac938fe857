"""Several engines in one process, the engine group (data-parallel forward from plain C calls) and the fused
final-LayerNorm + peer-store all-gather.  No torch here: the C ABI is driven through ctypes + numpy
only.  Everything except the two-device cases also runs on a one-GPU box: a "group" of two engines on the same device
exercises the whole sharding / gather protocol (peer pointers that happen to be local)."""
import ctypes
import os

import numpy as np
import pytest

import dinov2_b200 as d
from dinov2_b200 import synth

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
F16 = os.path.join(GOLD, "tiny_f16.gguf")


def _devices(n):
    have = d.device_count()
    return [i % have for i in range(n)]


def test_two_engines_in_one_process_are_independent():
    """Two engines alive at once (same or different device): interleaved calls, identical results, separate arenas."""
    imgs = synth.lcg_batch(0, 4, 70, 70)
    devs = _devices(2)
    with d.Engine(F16, device=devs[0]) as a, d.Engine(F16, device=devs[1]) as b:
        ra = a.forward(imgs, classify=True)
        rb = b.forward(imgs[::-1].copy(), classify=True)
        ra2 = a.forward(imgs, classify=True)
    assert np.array_equal(ra["patch_tokens"], rb["patch_tokens"][::-1])
    assert np.array_equal(ra["logits"], rb["logits"][::-1])
    assert np.array_equal(ra["probs"], ra2["probs"])


@pytest.mark.skipif(d.device_count() < 2, reason="needs two B200s in this process")
def test_engines_on_two_devices_match():
    """cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count are per device (round-1 advisor finding): an engine on
    device 1 created after one on device 0 must launch its > 48 KB shared-memory kernels just the same."""
    imgs = synth.lcg_batch(0, 3, 70, 70)
    with d.Engine(F16, device=0) as a, d.Engine(F16, device=1) as b:
        ra, rb = a.forward(imgs, classify=True), b.forward(imgs, classify=True)
    assert np.array_equal(ra["patch_tokens"], rb["patch_tokens"])
    assert np.array_equal(ra["logits"], rb["logits"])


@pytest.mark.parametrize("n,B", [(2, 5), (2, 4), (3, 7), (1, 3)])
def test_group_forward_matches_single_engine(n, B):
    """dino_b200_group_forward: contiguous shards (ragged last shard), outputs assembled in image order, bit-identical to one
    engine running the whole batch."""
    imgs = synth.lcg_batch(20, B, 70, 70)
    with d.Engine(F16) as e:
        want = e.forward(imgs, classify=True)
    with d.Group(F16, _devices(n)) as g:
        got = g.forward(imgs, classify=True)
        again = g.forward(imgs, classify=True)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
        assert np.array_equal(again[k], want[k]), k


@pytest.mark.parametrize("what", ["cls", "patch"])
@pytest.mark.parametrize("n,B", [(2, 5), (3, 6)])
def test_group_allgather_features(n, B, what):
    """The fused final-LayerNorm + all-gather: EVERY rank's buffer holds the features of the whole batch, bit-identical to the
    cls / patch outputs of a plain forward (same LayerNorm arithmetic, stored to n destinations)."""
    imgs = synth.lcg_batch(30, B, 70, 70)
    with d.Engine(F16) as e:
        want = e.forward(imgs, classify=False)
    sel = d.GATHER_CLS if what == "cls" else d.GATHER_PATCH
    ref = want["cls"][:, None, :] if what == "cls" else want["patch_tokens"]
    with d.Group(F16, _devices(n)) as g:
        for src in range(n):                                   # any rank serves the whole batch
            got, bufs = g.allgather_features(imgs, sel, host_from=src)
            assert got.shape == ref.shape
            assert np.array_equal(got, ref), (what, src)
        assert len(bufs) == n and all(bufs)
        # a smaller batch afterwards re-uses the buffers (rows of absent images keep their old contents, never read)
        got2, _ = g.allgather_features(imgs[:3], sel, host_from=n - 1)
        assert np.array_equal(got2, ref[:3])


@pytest.mark.skipif(d.device_count() < 2, reason="needs two B200s in this process")
def test_group_on_two_devices_peer_stores():
    """Same as above with the ranks on different GPUs: the gather rows travel as peer stores over NVLink."""
    imgs = synth.lcg_batch(40, 6, 70, 70)
    with d.Engine(F16) as e:
        want = e.forward(imgs, classify=True)
    with d.Group(F16, [0, 1]) as g:
        got = g.forward(imgs, classify=True)
        for k in want:
            assert np.array_equal(got[k], want[k]), k
        for src in (0, 1):
            f, _ = g.allgather_features(imgs, d.GATHER_PATCH, host_from=src)
            assert np.array_equal(f, want["patch_tokens"])
            c, _ = g.allgather_features(imgs, d.GATHER_CLS, host_from=src)
            assert np.array_equal(c[:, 0], want["cls"])


def test_gather_argument_errors():
    with d.Engine(F16) as e:
        with pytest.raises(d.DinoB200Error):
            e.forward_gather_device(1, d.LAYOUT_BGR_HWC, 1, 70, 70)       # no gather_init
        with pytest.raises(d.DinoB200Error):
            e.gather_init(0, 9, d.GATHER_CLS, 2, 70, 70)                  # world > 8
        e.gather_init(0, 2, d.GATHER_CLS, 2, 70, 70)
        with pytest.raises(d.DinoB200Error):
            e.gather_set_peer(0, dev_ptr=1234)                            # own rank
        with pytest.raises(d.DinoB200Error):
            e.forward_gather_device(1, d.LAYOUT_BGR_HWC, 1, 70, 70)       # rank 1 not registered


def test_submit_u8_pipeline_matches_forward_u8_and_pca():
    """dino_b200_submit_u8 (upload -> device preprocessing -> forward -> PCA colours -> read-back, two frames in flight)
    returns exactly what the synchronous dino_b200_forward_u8 + dino_b200_pca_rgb return for each frame, in order
    (reference loop: realtime.cpp:75-101)."""
    rng = np.random.default_rng(7)
    frames = [rng.integers(0, 256, size=(2, 75, 101, 3), dtype=np.uint8) for _ in range(4)]
    with d.Engine(F16) as e:
        want = []
        for f in frames:
            r = e.forward_u8(f, classify=False)
            r["pca_rgb"] = e.pca_rgb(r["patch_tokens"])
            want.append(r)
        NP, D = want[0]["patch_tokens"].shape[1:]
        outs = [{"cls": np.empty((2, D), np.float32), "patch_tokens": np.empty((2, NP, D), np.float32),
                 "pca_rgb": np.empty((2, NP, 3), np.uint8)} for _ in frames]
        hw = e.submit_u8(frames[0], outs[0])
        assert hw == e.preprocess_size(75, 101, False)
        for k in range(1, len(frames)):
            e.submit_u8(frames[k], outs[k])
            e.wait()
        e.wait()
        for k in range(len(frames)):
            for name in ("cls", "patch_tokens", "pca_rgb"):
                assert np.array_equal(outs[k][name], want[k][name]), (k, name)
        # colours only: the patch tokens never leave the device
        only = {"pca_rgb": np.empty((2, NP, 3), np.uint8)}
        e.submit_u8(frames[1], only)
        e.wait()
        assert np.array_equal(only["pca_rgb"], want[1]["pca_rgb"])
        # classify mode: 256x256 squash + 224 crop preprocessing, top-1 from raw frames
        c_want = e.forward_u8(frames[2], classify=True, want_patch=False)
        c_out = {"probs": np.empty_like(c_want["probs"]), "logits": np.empty_like(c_want["logits"])}
        assert e.submit_u8(frames[2], c_out, classify=True) == (224, 224)
        e.wait()
        assert np.array_equal(c_out["probs"], c_want["probs"])
