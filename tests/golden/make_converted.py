#!/usr/bin/env python
"""Fixtures for the converter path: checkpoints written by the reference's OWN toolchain.

    python tests/golden/make_converted.py          (needs /root/reference and oracle/_ref; run in the build container)

1. random-init HuggingFace Dinov2ForImageClassification / Dinov2WithRegistersForImageClassification (torch.manual_seed(0),
   hidden 128, 2 layers, 2 heads, 70 x 70, 10 labels) saved with save_pretrained into a directory whose name contains
   "imagenet" (the converter's classifier switch, dinov2-to-gguf.py:35),
2. the UNMODIFIED /root/reference/scripts/dinov2-to-gguf.py run on each (HF state_dict -> fused qkv -> gguf, :69-115),
3. the reference build (oracle/_ref) run on the LCG image 0 in features and classify mode.
Outputs: tests/golden/hf_conv_noreg.gguf, hf_conv_reg2.gguf, converted.npz.  Nothing here is imported by the product."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    import torch
    from transformers import (Dinov2Config, Dinov2ForImageClassification, Dinov2WithRegistersConfig,
                              Dinov2WithRegistersForImageClassification)
    from dinov2_b200 import synth
    import ref as refmod
    kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=70, patch_size=14, mlp_ratio=4,
              layerscale_value=1.0, use_swiglu_ffn=False, num_labels=10)
    out = {}
    img = synth.lcg_image(0, 70, 70)
    with tempfile.TemporaryDirectory() as td:
        for tag in ("noreg", "reg2"):
            torch.manual_seed(0)
            if tag == "noreg":
                m = Dinov2ForImageClassification(Dinov2Config(**kw))
            else:
                m = Dinov2WithRegistersForImageClassification(Dinov2WithRegistersConfig(num_register_tokens=2, **kw))
            mdir = os.path.join(td, f"tiny-imagenet-{tag}")
            m.save_pretrained(mdir)
            wd = os.path.join(td, "out_" + tag)
            os.makedirs(wd)
            env = dict(os.environ, PYTHONPATH=os.path.join(REF, "src"))
            subprocess.run([sys.executable, os.path.join(REF, "scripts", "dinov2-to-gguf.py"), "--model_name", mdir], cwd=wd, env=env,
                           check=True, stdout=subprocess.DEVNULL)
            dst = os.path.join(HERE, f"hf_conv_{tag}.gguf")
            shutil.copyfile(os.path.join(wd, "ggml-model.gguf"), dst)
            for classify in (False, True):
                R = refmod.Reference(dst, classify=classify, n_threads=2, H=70, W=70)
                o = R.forward(img)
                R.close()
                mode = "cls" if classify else "feat"
                if classify:
                    out[f"{tag}_{mode}_logits"], out[f"{tag}_{mode}_probs"] = o["logits"], o["probs"]
                else:
                    out[f"{tag}_{mode}_patch"], out[f"{tag}_{mode}_cls"] = o["patch_tokens"], o["cls"]
    np.savez_compressed(os.path.join(HERE, "converted.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
