"""ctypes driver for oracle/_ref/libdino_ref_*.so — the UNMODIFIED reference
(dinov2.cpp + ggml CPU) built by oracle/Makefile.  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product path."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_flags() -> set:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def builds() -> list:
    """The reference builds this host can run, best first: "v4" (x86-64-v4, AVX-512) and / or "v3" (x86-64-v3)."""
    flags = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    return [b for b in ((["v4"] if v4 else []) + ["v3"]) if os.path.exists(os.path.join(_HERE, "_ref", f"libdino_ref_{b}.so"))]


def lib_path() -> Optional[str]:
    forced = os.environ.get("DINO_REF_BUILD")          # "v3" / "v4": pick one build explicitly (self-noise measurements)
    if forced:
        p = os.path.join(_HERE, "_ref", f"libdino_ref_{forced}.so")
        return p if os.path.exists(p) else None
    flags = _cpu_flags()
    v4 = {"avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"} <= flags
    for name in (["libdino_ref_v4.so"] if v4 else []) + ["libdino_ref_v3.so"]:
        p = os.path.join(_HERE, "_ref", name)
        if os.path.exists(p):
            return p
    return None


def available() -> bool:
    return lib_path() is not None


_lib = None


def _load():
    global _lib
    if _lib is None:
        p = lib_path()
        if p is None:
            raise RuntimeError("oracle/_ref is not built (run `make -C oracle ref` where /root/reference exists)")
        L = C.CDLL(p)
        L.ref_load.restype = C.c_void_p
        L.ref_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_hparams.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        fp = C.POINTER(C.c_float)
        L.ref_forward.restype = C.c_double
        L.ref_forward.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, fp, fp, fp, fp]
        L.ref_predict.restype = C.c_double
        L.ref_predict.argtypes = [C.c_void_p, fp, C.c_int, C.c_int, fp]
        L.ref_interpolate_pos_embed.restype = C.c_int
        L.ref_interpolate_pos_embed.argtypes = [C.c_void_p, C.c_int, C.c_int, fp, C.c_int64]
        L.ref_preprocess.restype = C.c_int
        L.ref_preprocess.argtypes = [C.c_void_p, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, fp, C.c_int64,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _fp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


class Reference:
    """One loaded reference model (reference dino_model_load, dinov2.cpp:239)."""

    def __init__(self, gguf_path: str, classify: bool = False, n_threads: Optional[int] = None,
                 flash_attn: bool = False, H: int = 518, W: int = 518):
        L = _load()
        self.n_threads = n_threads or (os.cpu_count() or 1)
        self.classify = classify
        self.h = L.ref_load(gguf_path.encode(), self.n_threads, int(classify), int(flash_attn), H, W)
        if not self.h:
            raise RuntimeError(f"reference dino_model_load failed for {gguf_path}")
        hp = (C.c_uint32 * 8)()
        L.ref_hparams(self.h, hp)
        (self.hidden_size, self.num_hidden_layers, self.num_attention_heads, self.num_classes,
         self.num_register_tokens, self.patch_size, self.img_size, self.ftype) = [int(v) for v in hp]
        self.last_ms = 0.0

    def forward(self, img_bgr_hwc: np.ndarray) -> Dict[str, np.ndarray]:
        """Runs the reference graph (same steps as dino_predict) and returns every output."""
        L = _load()
        img = np.ascontiguousarray(img_bgr_hwc, dtype=np.float32)
        H, W, _ = img.shape
        D, R = self.hidden_size, self.num_register_tokens
        npatch = (H // self.patch_size) * (W // self.patch_size)
        n_out = npatch + (R if self.classify else 0)
        cls = np.empty(D, np.float32)
        patch = np.empty((n_out, D), np.float32)
        logits = np.empty(self.num_classes, np.float32) if self.classify else None
        probs = np.empty(self.num_classes, np.float32) if self.classify else None
        ms = L.ref_forward(self.h, _fp(img), H, W, _fp(cls), _fp(patch), _fp(logits), _fp(probs))
        if ms < 0:
            raise RuntimeError("reference graph compute failed")
        self.last_ms = ms
        out = {"cls": cls, "patch_tokens": patch[R:] if self.classify else patch}
        if self.classify:
            out["pool_tokens"] = patch
            out["logits"], out["probs"] = logits, probs
        return out

    def predict(self, img_bgr_hwc: np.ndarray) -> np.ndarray:
        """The reference's own dino_predict (features mode): returns patch tokens [NP, D]."""
        L = _load()
        img = np.ascontiguousarray(img_bgr_hwc, dtype=np.float32)
        H, W, _ = img.shape
        npatch = (H // self.patch_size) * (W // self.patch_size)
        patch = np.empty((npatch, self.hidden_size), np.float32)
        ms = L.ref_predict(self.h, _fp(img), H, W, _fp(patch))
        if ms < 0:
            raise RuntimeError("reference dino_predict failed")
        self.last_ms = ms
        return patch

    def interpolate_pos_embed(self, H: int, W: int) -> np.ndarray:
        L = _load()
        n = 1 + (H // self.patch_size) * (W // self.patch_size)
        out = np.empty((n, self.hidden_size), np.float32)
        got = L.ref_interpolate_pos_embed(self.h, H, W, _fp(out), out.size)
        assert got == out.size, (got, out.size)
        return out

    def preprocess(self, bgr_u8: np.ndarray, classify: bool) -> np.ndarray:
        L = _load()
        img = np.ascontiguousarray(bgr_u8, dtype=np.uint8)
        H, W, _ = img.shape
        cap = (H + 32) * (W + 32) * 3 if not classify else 224 * 224 * 3
        out = np.empty(cap, np.float32)
        oh, ow = C.c_int(), C.c_int()
        rc = L.ref_preprocess(self.h, img.ctypes.data_as(C.POINTER(C.c_uint8)), H, W, int(classify), _fp(out), cap,
                              C.byref(oh), C.byref(ow))
        assert rc == 0
        return out[: oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()

    def close(self):
        if self.h:
            _load().ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def forward_in_subprocess(build: str, gguf_path: str, img_bgr_hwc: np.ndarray, classify: bool) -> Dict[str, np.ndarray]:
    """Runs ONE forward of the reference build `build` ("v3" / "v4") in a child process (the two builds export the same
    symbols, so they are never loaded side by side) and returns its outputs.  Used to measure how far the unmodified
    reference is from ITSELF across its own compile targets — the noise floor a parity tolerance has to respect."""
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        inp, outp = os.path.join(td, "img.npy"), os.path.join(td, "out.npz")
        np.save(inp, np.ascontiguousarray(img_bgr_hwc, dtype=np.float32))
        env = dict(os.environ, DINO_REF_BUILD=build)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), gguf_path, inp, outp, "1" if classify else "0"], env=env,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference build {build} failed: {r.stderr[-500:]}")
        with np.load(outp) as z:
            return {k: z[k] for k in z.files}


if __name__ == "__main__":
    import sys
    _gguf, _inp, _outp, _cl = sys.argv[1:5]
    _img = np.load(_inp)
    _R = Reference(_gguf, classify=_cl == "1", H=_img.shape[0], W=_img.shape[1])
    _o = _R.forward(_img)
    _R.close()
    np.savez(_outp, **_o)
