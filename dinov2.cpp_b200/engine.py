"""ctypes binding of the C ABI in include/dinov2_b200.h (host side, Python).

Mirrors the reference's operator surface for the hot path — load a gguf
(`dino_model_load`, reference dinov2.cpp:239), run preprocessed images through
the network (`dino_predict`, dinov2.cpp:900) — with batching added.  All
compute happens in libdinov2_b200.so on an sm_100 GPU; there is NO CPU
fallback: if the library or a B200 is missing these calls raise."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DINO_B200_LIB selects another build of the library (A/B variants from tools/build_variant.sh); default: the product build
LIB_PATH = os.environ.get("DINO_B200_LIB") or os.path.join(_HERE, "lib", "libdinov2_b200.so")

LAYOUT_RGB_PLANAR = 0
LAYOUT_BGR_HWC = 1
FLAG_CLASSIFY = 1
FLAG_FLASH_ATTN_COMPAT = 2     # -fa: the reference flash path's unmasked zero-padding keys (dinov2.cpp:499-525)

EPI_BIAS_F16, EPI_GELU_F16, EPI_RESID_F32, EPI_SWIGLU_F16, EPI_PATCH_F32 = range(5)

STATUS = {0: "OK", 1: "ERR_INVALID", 2: "ERR_IO", 3: "ERR_FORMAT", 4: "ERR_CUDA", 5: "ERR_UNSUPPORTED", 6: "ERR_NO_DEVICE"}

# every symbol include/dinov2_b200.h declares (tests check the library exports exactly these)
ABI_SYMBOLS = [
    "dino_b200_device_count", "dino_b200_create", "dino_b200_create_from_gguf", "dino_b200_destroy",
    "dino_b200_get_hparams", "dino_b200_label", "dino_b200_reserve", "dino_b200_set_pos_embed",
    "dino_b200_get_pos_embed", "dino_b200_forward", "dino_b200_forward_device", "dino_b200_synchronize",
    "dino_b200_last_error", "dino_b200_kernel_launches", "dino_b200_set_profiling", "dino_b200_get_profile",
    "dino_b200_kernel_gemm", "dino_b200_kernel_gemm_resid_ln", "dino_b200_kernel_attention", "dino_b200_kernel_layernorm",
    "dino_b200_preprocess", "dino_b200_forward_u8", "dino_b200_submit", "dino_b200_wait",
    "dino_b200_pca_rgb", "dino_b200_pca_rgb_device", "dino_b200_quantize_gguf", "dino_b200_submit_u8",
    "dino_b200_gather_init", "dino_b200_gather_set_peer", "dino_b200_forward_gather_device",
    "dino_b200_group_create_from_gguf", "dino_b200_group_destroy", "dino_b200_group_size", "dino_b200_group_engine",
    "dino_b200_group_forward", "dino_b200_group_allgather_features",
]

GATHER_CLS, GATHER_PATCH = 1, 2
IPC_HANDLE_BYTES = 64


class HParams(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("hidden_size", "num_hidden_layers", "num_attention_heads", "num_classes",
                                           "num_register_tokens", "patch_size", "img_size", "ftype")] + [("eps", C.c_float)]


class DinoB200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"dinov2_b200: {STATUS.get(status, status)}: {msg}")
        self.status = status


_lib = None


def load_library() -> C.CDLL:
    """dlopen libdinov2_b200.so and declare prototypes. Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DinoB200Error(-1, f"{LIB_PATH} is missing — build it with __graft_entry__.build(); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, fp, ip = C.c_void_p, C.c_void_p, C.c_int
    L.dino_b200_device_count.restype = ip
    L.dino_b200_create_from_gguf.argtypes = [C.c_char_p, ip, C.POINTER(vp)]
    L.dino_b200_create.argtypes = [vp, ip, C.POINTER(vp)]
    L.dino_b200_destroy.argtypes = [vp]
    L.dino_b200_destroy.restype = None
    L.dino_b200_get_hparams.argtypes = [vp, C.POINTER(HParams)]
    L.dino_b200_label.argtypes = [vp, ip]
    L.dino_b200_label.restype = C.c_char_p
    L.dino_b200_reserve.argtypes = [vp, ip, ip, ip]
    L.dino_b200_set_pos_embed.argtypes = [vp, ip, ip, fp]
    L.dino_b200_get_pos_embed.argtypes = [vp, ip, ip, fp]
    L.dino_b200_forward.argtypes = [vp, fp, ip, ip, ip, ip, ip, fp, fp, fp, fp]
    L.dino_b200_forward_device.argtypes = [vp, fp, ip, ip, ip, ip, ip, fp, fp, fp, fp, vp]
    L.dino_b200_synchronize.argtypes = [vp]
    L.dino_b200_preprocess.argtypes = [vp, vp, ip, ip, ip, ip, fp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.dino_b200_forward_u8.argtypes = [vp, vp, ip, ip, ip, ip, fp, fp, fp, fp]
    L.dino_b200_submit.argtypes = [vp, fp, ip, ip, ip, ip, ip, fp, fp, fp, fp]
    L.dino_b200_wait.argtypes = [vp]
    L.dino_b200_submit_u8.argtypes = [vp, vp, ip, ip, ip, ip, fp, fp, fp, fp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.dino_b200_gather_init.argtypes = [vp, ip, ip, ip, ip, ip, ip, C.POINTER(vp), vp]
    L.dino_b200_gather_set_peer.argtypes = [vp, ip, vp, vp]
    L.dino_b200_forward_gather_device.argtypes = [vp, fp, ip, ip, ip, ip, ip, fp, fp, fp, fp, vp]
    L.dino_b200_group_create_from_gguf.argtypes = [C.c_char_p, C.POINTER(C.c_int), ip, C.POINTER(vp)]
    L.dino_b200_group_destroy.argtypes = [vp]
    L.dino_b200_group_destroy.restype = None
    L.dino_b200_group_size.argtypes = [vp]
    L.dino_b200_group_engine.argtypes = [vp, ip]
    L.dino_b200_group_engine.restype = vp
    L.dino_b200_group_forward.argtypes = [vp, fp, ip, ip, ip, ip, ip, fp, fp, fp, fp]
    L.dino_b200_group_allgather_features.argtypes = [vp, fp, ip, ip, ip, ip, ip, C.POINTER(vp), fp, ip]
    L.dino_b200_quantize_gguf.argtypes = [C.c_char_p, C.c_char_p, ip]
    L.dino_b200_pca_rgb.argtypes = [vp, fp, ip, ip, vp, fp]
    L.dino_b200_pca_rgb_device.argtypes = [vp, vp, ip, ip, vp, vp, vp]
    L.dino_b200_last_error.argtypes = [vp]
    L.dino_b200_last_error.restype = C.c_char_p
    L.dino_b200_kernel_launches.argtypes = [vp]
    L.dino_b200_kernel_launches.restype = C.c_uint64
    L.dino_b200_set_profiling.argtypes = [vp, ip]
    pf = C.POINTER(C.c_float)
    L.dino_b200_get_profile.argtypes = [vp, pf, pf, pf, pf]
    L.dino_b200_kernel_gemm.argtypes = [ip, vp, ip, vp, ip, ip, ip, ip, vp, vp, vp, ip, vp, ip, ip, ip, vp]
    L.dino_b200_kernel_attention.argtypes = [vp, vp, ip, ip, ip, vp]
    L.dino_b200_kernel_layernorm.argtypes = [vp, vp, vp, vp, ip, ip, C.c_float, ip, vp]
    L.dino_b200_kernel_gemm_resid_ln.argtypes = [vp, ip, vp, ip, ip, ip, ip, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp]
    _lib = L
    return L


def _check(st: int, handle=None):
    if st != 0:
        L = load_library()
        msg = L.dino_b200_last_error(handle)
        raise DinoB200Error(st, (msg or b"").decode(errors="replace"))


def device_count() -> int:
    return int(load_library().dino_b200_device_count())


def _host_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def quantize_gguf(fname_inp: str, fname_out: str, ggml_type: int) -> None:
    """The reference's `quantize` tool (dino_model_quantize, dinov2.cpp:354-452): host-only, works without a GPU.
    ggml_type: 2 q4_0, 3 q4_1, 6 q5_0, 7 q5_1, 8 q8_0."""
    _check(load_library().dino_b200_quantize_gguf(os.fsencode(fname_inp), os.fsencode(fname_out), ggml_type))


class Engine:
    """One model resident on one GPU.  `forward` takes host arrays, `forward_device` raw device pointers."""

    def __init__(self, gguf_path: str, device: int = 0):
        L = load_library()
        h = C.c_void_p()
        _check(L.dino_b200_create_from_gguf(os.fsencode(gguf_path), device, C.byref(h)))
        self._h = h
        self.device = device
        hp = HParams()
        _check(L.dino_b200_get_hparams(self._h, C.byref(hp)), self._h)
        self.hparams = {n: getattr(hp, n) for n, _ in HParams._fields_}
        for k, v in self.hparams.items():
            setattr(self, k, v)

    # -- lifecycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            load_library().dino_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- helpers -----------------------------------------------------------
    def n_patches(self, H: int, W: int) -> int:
        return (H // self.patch_size) * (W // self.patch_size)

    def label(self, class_id: int) -> Optional[str]:
        s = load_library().dino_b200_label(self._h, class_id)
        return s.decode() if s else None

    def reserve(self, max_batch: int, H: int, W: int):
        _check(load_library().dino_b200_reserve(self._h, max_batch, H, W), self._h)

    def set_pos_embed(self, gh: int, gw: int, pos: np.ndarray):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        assert pos.shape == (1 + gh * gw, self.hidden_size)
        _check(load_library().dino_b200_set_pos_embed(self._h, gh, gw, pos.ctypes.data), self._h)

    def get_pos_embed(self, H: int, W: int) -> np.ndarray:
        out = np.empty((1 + self.n_patches(H, W), self.hidden_size), np.float32)
        _check(load_library().dino_b200_get_pos_embed(self._h, H, W, out.ctypes.data), self._h)
        return out

    def synchronize(self):
        _check(load_library().dino_b200_synchronize(self._h), self._h)

    @property
    def kernel_launches(self) -> int:
        return int(load_library().dino_b200_kernel_launches(self._h))

    def set_profiling(self, on: bool):
        _check(load_library().dino_b200_set_profiling(self._h, int(on)), self._h)

    def get_profile(self) -> Dict[str, float]:
        v = [C.c_float() for _ in range(4)]
        _check(load_library().dino_b200_get_profile(self._h, *[C.byref(x) for x in v]), self._h)
        return dict(zip(("gemm_ms", "attn_ms", "other_ms", "total_ms"), [x.value for x in v]))

    # -- the hot path ------------------------------------------------------
    def forward(self, images: np.ndarray, classify: bool = False, layout: int = LAYOUT_BGR_HWC,
                want_patch: bool = True, want_cls: bool = True, out: Optional[Dict[str, np.ndarray]] = None,
                flash_attn_compat: bool = False) -> Dict[str, np.ndarray]:
        """images: float32 [B,H,W,3] (BGR_HWC, what dino_preprocess returns) or [B,3,H,W] (RGB_PLANAR).
        Host in, host out; synchronous.  `out` may carry preallocated (e.g. pinned) result arrays."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        if images.ndim != 4:
            raise ValueError("images must be 4-D")
        if layout == LAYOUT_BGR_HWC:
            B, H, W, ch = images.shape
        else:
            B, ch, H, W = images.shape
        if ch != 3:
            raise ValueError("images must have 3 channels")
        D, NP, Cn = self.hidden_size, self.n_patches(H, W), self.num_classes
        res = dict(out or {})
        if want_cls and "cls" not in res:
            res["cls"] = np.empty((B, D), np.float32)
        if want_patch and "patch_tokens" not in res:
            res["patch_tokens"] = np.empty((B, NP, D), np.float32)
        if classify:
            res.setdefault("logits", np.empty((B, Cn), np.float32))
            res.setdefault("probs", np.empty((B, Cn), np.float32))
        _check(load_library().dino_b200_forward(
            self._h, images.ctypes.data, layout, B, H, W, (FLAG_CLASSIFY if classify else 0) | (FLAG_FLASH_ATTN_COMPAT if flash_attn_compat else 0),
            _host_ptr(res.get("cls")), _host_ptr(res.get("patch_tokens")),
            _host_ptr(res.get("logits")), _host_ptr(res.get("probs"))), self._h)
        return res

    def submit(self, images: np.ndarray, out: Dict[str, np.ndarray], classify: bool = False, layout: int = LAYOUT_BGR_HWC):
        """Pipelined forward: enqueue upload + forward + read-back of one batch and return at once (at most two in flight).
        `images` (float32, C-contiguous) and the arrays in `out` ("cls", "patch_tokens", "logits", "probs" — whichever are
        wanted) must stay alive, and should be pinned, until the matching wait() returns."""
        if images.dtype != np.float32 or not images.flags["C_CONTIGUOUS"] or images.ndim != 4:
            raise ValueError("submit needs a C-contiguous float32 4-D array (no hidden copy may be made)")
        if layout == LAYOUT_BGR_HWC:
            B, H, W, ch = images.shape
        else:
            B, ch, H, W = images.shape
        if ch != 3:
            raise ValueError("images must have 3 channels")
        _check(load_library().dino_b200_submit(
            self._h, images.ctypes.data, layout, B, H, W, FLAG_CLASSIFY if classify else 0, _host_ptr(out.get("cls")),
            _host_ptr(out.get("patch_tokens")), _host_ptr(out.get("logits")), _host_ptr(out.get("probs"))), self._h)

    def wait(self):
        """Block until the oldest submitted batch has completed."""
        _check(load_library().dino_b200_wait(self._h), self._h)

    def submit_u8(self, frames_u8: np.ndarray, out: Dict[str, np.ndarray], classify: bool = False):
        """Pipelined raw-frame path (realtime.cpp's loop): uint8 BGR frames [B,H,W,3] in; the arrays present in `out` ("cls",
        "patch_tokens", "logits", "probs", "pca_rgb" — uint8 [B,NP,3]) are filled when the matching wait() returns.
        Returns the preprocessed (H, W)."""
        if frames_u8.dtype != np.uint8 or not frames_u8.flags["C_CONTIGUOUS"] or frames_u8.ndim != 4 or frames_u8.shape[3] != 3:
            raise ValueError("submit_u8 needs a C-contiguous uint8 [B,H,W,3] array (no hidden copy may be made)")
        B, H, W, _ = frames_u8.shape
        oh, ow = C.c_int(), C.c_int()
        _check(load_library().dino_b200_submit_u8(
            self._h, frames_u8.ctypes.data, B, H, W, FLAG_CLASSIFY if classify else 0, _host_ptr(out.get("cls")),
            _host_ptr(out.get("patch_tokens")), _host_ptr(out.get("logits")), _host_ptr(out.get("probs")),
            _host_ptr(out.get("pca_rgb")), C.byref(oh), C.byref(ow)), self._h)
        return oh.value, ow.value

    # -- feature all-gather (one engine = one rank) -------------------------
    def gather_init(self, rank: int, world: int, what: int, max_batch: int, H: int, W: int):
        """Allocates this rank's gather buffer; returns (device pointer, 64-byte CUDA IPC handle)."""
        buf = C.c_void_p()
        handle = (C.c_ubyte * IPC_HANDLE_BYTES)()
        _check(load_library().dino_b200_gather_init(self._h, rank, world, what, max_batch, H, W, C.byref(buf), handle), self._h)
        return buf.value, bytes(handle)

    def gather_set_peer(self, r: int, dev_ptr: int = 0, ipc_handle: Optional[bytes] = None):
        h = (C.c_ubyte * IPC_HANDLE_BYTES).from_buffer_copy(ipc_handle) if ipc_handle else None
        _check(load_library().dino_b200_gather_set_peer(self._h, r, dev_ptr or None, h), self._h)

    def forward_gather_device(self, images_ptr: int, layout: int, B: int, H: int, W: int, classify: bool = False, cls_ptr: int = 0,
                              patch_ptr: int = 0, logits_ptr: int = 0, probs_ptr: int = 0, stream: int = 0):
        """forward_device + the fused final-LayerNorm / peer-store all-gather into every registered gather buffer."""
        _check(load_library().dino_b200_forward_gather_device(
            self._h, images_ptr, layout, B, H, W, FLAG_CLASSIFY if classify else 0, cls_ptr or None, patch_ptr or None,
            logits_ptr or None, probs_ptr or None, stream or None), self._h)

    def pca_rgb(self, patch_tokens: np.ndarray, want_proj: bool = False):
        """PCA colouring of inference.cpp:76-86 on the device: patch_tokens float32 [B, NP, D] -> uint8 [B, NP, 3]
        (and, optionally, the float projections [B, NP, 3])."""
        x = np.ascontiguousarray(patch_tokens, dtype=np.float32)
        B, NP, D = x.shape
        assert D == self.hidden_size
        rgb = np.empty((B, NP, 3), np.uint8)
        proj = np.empty((B, NP, 3), np.float32) if want_proj else None
        _check(load_library().dino_b200_pca_rgb(self._h, x.ctypes.data, B, NP, rgb.ctypes.data, _host_ptr(proj)), self._h)
        return (rgb, proj) if want_proj else rgb

    def pca_rgb_device(self, patch_ptr: int, B: int, NP: int, rgb_ptr: int = 0, proj_ptr: int = 0, stream: int = 0):
        _check(load_library().dino_b200_pca_rgb_device(self._h, patch_ptr, B, NP, rgb_ptr or None, proj_ptr or None, stream or None), self._h)

    def preprocess_size(self, H: int, W: int, classify: bool):
        """Output size of the reference's preprocessing for an H x W frame (dinov2.cpp:111-116, 140-141)."""
        if classify:
            return 224, 224
        ps = self.patch_size
        return (H // ps + 1) * ps, (W // ps + 1) * ps

    def preprocess(self, frames_u8: np.ndarray, classify: bool = False) -> np.ndarray:
        """uint8 BGR frames [B,H,W,3] -> float32 [B,OH,OW,3] (device version of dino_preprocess / dino_classify_preprocess)."""
        frames = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        B, H, W, ch = frames.shape
        assert ch == 3
        OH, OW = self.preprocess_size(H, W, classify)
        out = np.empty((B, OH, OW, 3), np.float32)
        oh, ow = C.c_int(), C.c_int()
        _check(load_library().dino_b200_preprocess(self._h, frames.ctypes.data, B, H, W, int(classify), out.ctypes.data,
                                                   C.byref(oh), C.byref(ow)), self._h)
        assert (oh.value, ow.value) == (OH, OW)
        return out

    def forward_u8(self, frames_u8: np.ndarray, classify: bool = False, want_patch: bool = True, want_cls: bool = True
                   ) -> Dict[str, np.ndarray]:
        """Raw uint8 BGR frames in, results out: preprocessing and forward pass both on the device."""
        frames = np.ascontiguousarray(frames_u8, dtype=np.uint8)
        B, H, W, ch = frames.shape
        assert ch == 3
        OH, OW = self.preprocess_size(H, W, classify)
        res: Dict[str, np.ndarray] = {}
        if want_cls:
            res["cls"] = np.empty((B, self.hidden_size), np.float32)
        if want_patch:
            res["patch_tokens"] = np.empty((B, self.n_patches(OH, OW), self.hidden_size), np.float32)
        if classify:
            res["logits"] = np.empty((B, self.num_classes), np.float32)
            res["probs"] = np.empty((B, self.num_classes), np.float32)
        _check(load_library().dino_b200_forward_u8(
            self._h, frames.ctypes.data, B, H, W, FLAG_CLASSIFY if classify else 0, _host_ptr(res.get("cls")),
            _host_ptr(res.get("patch_tokens")), _host_ptr(res.get("logits")), _host_ptr(res.get("probs"))), self._h)
        return res

    def forward_device(self, images_ptr: int, layout: int, B: int, H: int, W: int, classify: bool = False,
                       cls_ptr: int = 0, patch_ptr: int = 0, logits_ptr: int = 0, probs_ptr: int = 0, stream: int = 0):
        """All pointers are device addresses (e.g. torch.Tensor.data_ptr()); asynchronous on `stream`."""
        _check(load_library().dino_b200_forward_device(
            self._h, images_ptr, layout, B, H, W, FLAG_CLASSIFY if classify else 0,
            cls_ptr or None, patch_ptr or None, logits_ptr or None, probs_ptr or None, stream or None), self._h)


class Group:
    """n engines on n devices of one process (dino_b200_group_*): data-parallel forward and the fused feature all-gather,
    driven from plain C calls — no torch, no NCCL."""

    def __init__(self, gguf_path: str, devices):
        L = load_library()
        self._g = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        _check(L.dino_b200_group_create_from_gguf(os.fsencode(gguf_path), devs, len(devices), C.byref(self._g)))
        self.n = int(L.dino_b200_group_size(self._g))
        hp = HParams()
        _check(L.dino_b200_get_hparams(L.dino_b200_group_engine(self._g, 0), C.byref(hp)))
        self.hp = hp

    def close(self):
        if self._g:
            load_library().dino_b200_group_destroy(self._g)
            self._g = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def forward(self, images: np.ndarray, classify: bool = False, layout: int = LAYOUT_BGR_HWC, want_patch: bool = True) -> Dict[str, np.ndarray]:
        images = np.ascontiguousarray(images, dtype=np.float32)
        B, H, W = (images.shape[0], images.shape[1], images.shape[2]) if layout == LAYOUT_BGR_HWC else (images.shape[0], images.shape[2], images.shape[3])
        D, ps, Cn = self.hp.hidden_size, self.hp.patch_size, self.hp.num_classes
        NP = (H // ps) * (W // ps)
        res = {"cls": np.empty((B, D), np.float32)}
        if want_patch:
            res["patch_tokens"] = np.empty((B, NP, D), np.float32)
        if classify:
            res["logits"] = np.empty((B, Cn), np.float32)
            res["probs"] = np.empty((B, Cn), np.float32)
        _check(load_library().dino_b200_group_forward(self._g, images.ctypes.data, layout, B, H, W, FLAG_CLASSIFY if classify else 0,
                                                      _host_ptr(res.get("cls")), _host_ptr(res.get("patch_tokens")),
                                                      _host_ptr(res.get("logits")), _host_ptr(res.get("probs"))))
        return res

    def allgather_features(self, images: np.ndarray, what: int = GATHER_CLS, layout: int = LAYOUT_BGR_HWC, host_from: int = 0):
        """Returns (features [B, rows, D] copied from device `host_from`, list of the n per-device gather buffer pointers)."""
        images = np.ascontiguousarray(images, dtype=np.float32)
        B, H, W = (images.shape[0], images.shape[1], images.shape[2]) if layout == LAYOUT_BGR_HWC else (images.shape[0], images.shape[2], images.shape[3])
        D, ps = self.hp.hidden_size, self.hp.patch_size
        rows = 1 if what == GATHER_CLS else (H // ps) * (W // ps)
        out = np.empty((B, rows, D), np.float32)
        bufs = (C.c_void_p * self.n)()
        _check(load_library().dino_b200_group_allgather_features(self._g, images.ctypes.data, layout, B, H, W, what, bufs, out.ctypes.data, host_from))
        return out, [b for b in bufs]


# ---- kernel-level hooks (device pointers) ---------------------------------
def kernel_gemm(epi: int, A: int, lda: int, W: int, ldw: int, M: int, N: int, K: int, bias: int, lscale: int, out: int,
                ldo: int, pos: int = 0, np_: int = 0, ntok: int = 0, tok_off: int = 0, stream: int = 0):
    _check(load_library().dino_b200_kernel_gemm(epi, A, lda, W, ldw, M, N, K, bias, lscale or None, out, ldo, pos or None,
                                                np_, ntok, tok_off, stream or None))


def kernel_gemm_resid_ln(A: int, lda: int, W: int, ldw: int, M: int, N: int, K: int, bias: int, lscale: int, X: int, gamma: int,
                         beta: int, eps: float, ln_out: int, counters: int, stream: int = 0):
    _check(load_library().dino_b200_kernel_gemm_resid_ln(A, lda, W, ldw, M, N, K, bias, lscale, X, gamma, beta, eps, ln_out,
                                                         counters, stream or None))


def kernel_attention(qkv: int, out: int, B: int, n_tok: int, D: int, stream: int = 0):
    _check(load_library().dino_b200_kernel_attention(qkv, out, B, n_tok, D, stream or None))


def kernel_layernorm(X: int, gamma: int, beta: int, out: int, rows: int, D: int, eps: float, out_half: bool, stream: int = 0):
    _check(load_library().dino_b200_kernel_layernorm(X, gamma, beta, out, rows, D, eps, int(out_half), stream or None))
