#!/usr/bin/env python
"""Golden vectors of the reference's `-fa` path (dino_params::enable_flash_attn, dinov2.cpp:499-525) from the UNMODIFIED
reference build (oracle/_ref): python tests/golden/make_golden_fa.py   ->  golden_fa.npz
    fa_feat_{cls,patch}   tiny_f16.gguf, 70x70 LCG image 0   (28 tokens -> padded to 32: 4 phantom keys)
    fa_nn_{cls,patch}     98x84 LCG image 3                   (45 tokens -> 64: 19 phantom keys)
    fa_cls_{logits,probs} classify mode, 70x70 image 0"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from dinov2_b200 import synth  # noqa: E402
import ref  # noqa: E402

f16 = os.path.join(HERE, "tiny_f16.gguf")
out = {}
R = ref.Reference(f16, classify=False, n_threads=4, flash_attn=True, H=70, W=70)
o = R.forward(synth.lcg_image(0, 70, 70))
out["fa_feat_cls"], out["fa_feat_patch"] = o["cls"], o["patch_tokens"]
o = R.forward(synth.lcg_image(3, 98, 84))
out["fa_nn_cls"], out["fa_nn_patch"] = o["cls"], o["patch_tokens"]
R.close()
R = ref.Reference(f16, classify=True, n_threads=4, flash_attn=True, H=70, W=70)
o = R.forward(synth.lcg_image(0, 70, 70))
out["fa_cls_logits"], out["fa_cls_probs"] = o["logits"], o["probs"]
R.close()
np.savez_compressed(os.path.join(HERE, "golden_fa.npz"), **out)
print({k: v.shape for k, v in out.items()})
