#!/usr/bin/env python
"""Regenerates the committed golden fixtures with the UNMODIFIED reference build (oracle/_ref, compiled from
/root/reference by oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py

Fixtures (all small):
  tiny_f16.gguf / tiny_q8_0.gguf   seeded synthetic checkpoints (synth.CONFIGS['tiny'], seed 1); the q8_0 file is
                                   produced by the reference's own `quantize` tool from the f16 file
  golden.npz                       reference outputs for LCG images:
      f16_feat_{cls,patch}     70x70 image 0, features mode        f16_cls_{logits,probs}  classify mode
      f16_nn_{cls,patch}       98x84 image 3 (non-native grid -> bicubic pos-embed), features mode
      f16_nn_pos               the reference's interpolate_pos_embed output for 98x84
      q8_feat_*, q8_cls_*      same for the q8_0 checkpoint
      q4_0_*, q4_1_*, q5_0_*, q5_1_*   same for the other four types the reference's quantize tool offers (README "Quantization");
                                   tiny_q4_0.gguf .. tiny_q5_1.gguf are written by that tool (type ids 2, 3, 6, 7)
The reference is ISA/flag dependent at the 1e-7 NMSE level, which the test tolerances absorb."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import dinov2_b200  # noqa: E402
from dinov2_b200 import synth  # noqa: E402
import ref  # noqa: E402

cfg = synth.CONFIGS["tiny"]
f16 = os.path.join(HERE, "tiny_f16.gguf")
q8 = os.path.join(HERE, "tiny_q8_0.gguf")
synth.write_synth_gguf(f16, cfg, seed=1)
subprocess.run([os.path.join(ROOT, "oracle", "_ref", "quantize"), f16, q8, "8"], check=True, capture_output=True)

cases = [("f16", f16), ("q8", q8)]
for tag, itype in (("q4_0", 2), ("q4_1", 3), ("q5_0", 6), ("q5_1", 7)):
    path = os.path.join(HERE, f"tiny_{tag}.gguf")
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "quantize"), f16, path, str(itype)], check=True, capture_output=True)
    cases.append((tag, path))

out = {}
for tag, path in cases:
    img = synth.lcg_image(0, 70, 70)
    R = ref.Reference(path, classify=False, n_threads=4, H=70, W=70)
    o = R.forward(img)
    out[f"{tag}_feat_cls"], out[f"{tag}_feat_patch"] = o["cls"], o["patch_tokens"]
    assert np.array_equal(R.predict(img), o["patch_tokens"])        # the reference's real dino_predict agrees
    if tag == "f16":
        img_nn = synth.lcg_image(3, 98, 84)
        o = R.forward(img_nn)
        out["f16_nn_cls"], out["f16_nn_patch"] = o["cls"], o["patch_tokens"]
        out["f16_nn_pos"] = R.interpolate_pos_embed(98, 84)
    R.close()
    R = ref.Reference(path, classify=True, n_threads=4, H=70, W=70)
    o = R.forward(img)
    out[f"{tag}_cls_logits"], out[f"{tag}_cls_probs"] = o["logits"], o["probs"]
    R.close()
np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
print({k: v.shape for k, v in out.items()})

# ---- preprocessing goldens come from REAL OpenCV (the Python cv2 wheel, 4.13), not from the shim:
#   rng = np.random.default_rng(7); img = rng.integers(0, 256, (61, 83, 3), uint8)   (BGR)
#   f = img.astype(float32) * float32(1/255); feat = (cv2.resize(f, (84, 70), INTER_CUBIC) - mean_bgr) / std_bgr
#   cls = (cv2.resize(f, (256, 256), INTER_CUBIC)[16:240, 16:240] - mean_bgr) / std_bgr      -> preprocess_cv2.npz
