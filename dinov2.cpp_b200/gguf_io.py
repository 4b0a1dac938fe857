"""GGUF v3 container I/O for DINOv2 checkpoints (host side, numpy only).

The reference reads checkpoints with ggml's gguf.cpp (`gguf_init_from_file`,
reference dinov2.cpp:263-272) and its converter writes them with the `gguf`
Python package (reference scripts/dinov2-to-gguf.py:122-166).  This module is
an independent reader/writer for exactly the subset of the format those two
produce/consume — format described in reference ggml/include/gguf.h:1-46:

    magic "GGUF" | version u32 (=3) | n_tensors u64 | n_kv u64
    KV pairs      : key(str) type(u32) value
    tensor infos  : name(str) n_dims(u32) ne[n_dims](u64) ggml_type(u32) offset(u64)
    padding to `general.alignment` (default 32)
    tensor data   : each tensor at data_start + offset, offsets aligned

Only the tensor types the DINOv2 converter/quantiser emit on the north-star
configs are decoded: F32, F16, Q8_0 (reference ggml-common.h:209-213).
The engine's own C++ loader (csrc/gguf_reader.cpp) implements the same layout;
tests cross-check the two and check both against files written by the
reference converter.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

GGUF_MAGIC = 0x46554747
GGUF_VERSION = 3
DEFAULT_ALIGNMENT = 32

# ggml_type ids (reference ggml/include/ggml.h enum ggml_type)
GGML_TYPE_F32 = 0
GGML_TYPE_F16 = 1
GGML_TYPE_Q4_0 = 2
GGML_TYPE_Q4_1 = 3
GGML_TYPE_Q5_0 = 6
GGML_TYPE_Q5_1 = 7
GGML_TYPE_Q8_0 = 8
QK8_0 = 32
Q8_0_BLOCK_BYTES = 34  # half d + 32 x int8
# bytes per 32-element block of the legacy ggml quant types (reference ggml-common.h:170-213):
#   q4_0 {half d; u8 qs[16]}  q4_1 {half d, m; u8 qs[16]}  q5_0 {half d; u8 qh[4]; u8 qs[16]}  q5_1 {half d, m; u8 qh[4]; u8 qs[16]}
QUANT_BLOCK_BYTES = {GGML_TYPE_Q4_0: 18, GGML_TYPE_Q4_1: 20, GGML_TYPE_Q5_0: 22, GGML_TYPE_Q5_1: 24, GGML_TYPE_Q8_0: 34}

# gguf value types (reference ggml/include/gguf.h enum gguf_type)
T_U8, T_I8, T_U16, T_I16, T_U32, T_I32, T_F32, T_BOOL, T_STR, T_ARR, T_U64, T_I64, T_F64 = range(13)
_SCALAR_FMT = {T_U8: "<B", T_I8: "<b", T_U16: "<H", T_I16: "<h", T_U32: "<I", T_I32: "<i",
               T_F32: "<f", T_BOOL: "<?", T_U64: "<Q", T_I64: "<q", T_F64: "<d"}


@dataclass
class GGUFTensor:
    name: str
    ne: Tuple[int, ...]          # ggml order: fastest-varying dimension first
    ggml_type: int
    data: np.ndarray             # raw bytes (uint8, 1-D), exactly as stored in the file

    @property
    def nelements(self) -> int:
        n = 1
        for d in self.ne:
            n *= d
        return n


@dataclass
class GGUFFile:
    kv: Dict[str, object] = field(default_factory=dict)
    kv_types: Dict[str, int] = field(default_factory=dict)
    tensors: Dict[str, GGUFTensor] = field(default_factory=dict)   # insertion order == file order


def tensor_nbytes(ggml_type: int, ne) -> int:
    n = 1
    for d in ne:
        n *= int(d)
    if ggml_type == GGML_TYPE_F32:
        return n * 4
    if ggml_type == GGML_TYPE_F16:
        return n * 2
    if ggml_type in QUANT_BLOCK_BYTES:
        assert ne[0] % QK8_0 == 0, "quantised rows must be a multiple of 32"
        return n // QK8_0 * QUANT_BLOCK_BYTES[ggml_type]
    raise ValueError(f"unsupported ggml type {ggml_type}")


# ----------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------
class _Cursor:
    def __init__(self, buf: memoryview):
        self.buf, self.pos = buf, 0

    def take(self, fmt: str):
        v = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += struct.calcsize(fmt)
        return v[0]

    def string(self) -> str:
        n = self.take("<Q")
        s = bytes(self.buf[self.pos:self.pos + n]).decode("utf-8")
        self.pos += n
        return s

    def value(self, t: int):
        if t == T_STR:
            return self.string()
        if t == T_ARR:
            et = self.take("<I")
            n = self.take("<Q")
            return [self.value(et) for _ in range(n)]
        return self.take(_SCALAR_FMT[t])


def read_gguf(path: str) -> GGUFFile:
    raw = np.fromfile(path, dtype=np.uint8)
    cur = _Cursor(memoryview(raw))
    if cur.take("<I") != GGUF_MAGIC:
        raise ValueError(f"{path}: not a GGUF file")
    version = cur.take("<I")
    if version not in (2, 3):
        raise ValueError(f"{path}: unsupported GGUF version {version}")
    n_tensors = cur.take("<Q")
    n_kv = cur.take("<Q")
    out = GGUFFile()
    for _ in range(n_kv):
        key = cur.string()
        t = cur.take("<I")
        out.kv[key] = cur.value(t)
        out.kv_types[key] = t
    infos = []
    for _ in range(n_tensors):
        name = cur.string()
        nd = cur.take("<I")
        ne = tuple(cur.take("<Q") for _ in range(nd))
        gt = cur.take("<I")
        off = cur.take("<Q")
        infos.append((name, ne, gt, off))
    align = int(out.kv.get("general.alignment", DEFAULT_ALIGNMENT))
    data_start = (cur.pos + align - 1) // align * align
    for name, ne, gt, off in infos:
        nb = tensor_nbytes(gt, ne)
        out.tensors[name] = GGUFTensor(name, ne, gt, raw[data_start + off: data_start + off + nb])
    return out


# ----------------------------------------------------------------------------
# writer
# ----------------------------------------------------------------------------
def _pack_string(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack("<Q", len(b)) + b


def _pack_value(t: int, v) -> bytes:
    if t == T_STR:
        return _pack_string(v)
    if t == T_ARR:
        raise NotImplementedError("array KVs are not produced by the DINOv2 converter")
    return struct.pack(_SCALAR_FMT[t], v)


def write_gguf(path: str, kv: List[Tuple[str, int, object]], tensors: List[GGUFTensor],
               alignment: int = DEFAULT_ALIGNMENT) -> None:
    head = struct.pack("<IIQQ", GGUF_MAGIC, GGUF_VERSION, len(tensors), len(kv))
    body = bytearray()
    for key, t, v in kv:
        body += _pack_string(key) + struct.pack("<I", t) + _pack_value(t, v)
    offsets, off = [], 0
    for t in tensors:
        nb = tensor_nbytes(t.ggml_type, t.ne)
        assert t.data.dtype == np.uint8 and t.data.size == nb, (t.name, t.data.size, nb)
        offsets.append(off)
        off = (off + nb + alignment - 1) // alignment * alignment
    for t, o in zip(tensors, offsets):
        body += _pack_string(t.name) + struct.pack("<I", len(t.ne))
        for d in t.ne:
            body += struct.pack("<Q", d)
        body += struct.pack("<IQ", t.ggml_type, o)
    meta = head + bytes(body)
    pad = (-len(meta)) % alignment
    with open(path, "wb") as f:
        f.write(meta + b"\0" * pad)
        for t in tensors:
            f.write(t.data.tobytes())
            f.write(b"\0" * ((-t.data.size) % alignment))


# ----------------------------------------------------------------------------
# tensor <-> numpy helpers
# ----------------------------------------------------------------------------
def f32_tensor(name: str, arr: np.ndarray) -> GGUFTensor:
    """arr in numpy (slowest-first) order; ggml ne is the reverse."""
    a = np.ascontiguousarray(arr, dtype=np.float32)
    return GGUFTensor(name, tuple(reversed(a.shape)), GGML_TYPE_F32, a.view(np.uint8).reshape(-1))


def f16_tensor(name: str, arr: np.ndarray) -> GGUFTensor:
    a = np.ascontiguousarray(arr, dtype=np.float16)
    return GGUFTensor(name, tuple(reversed(a.shape)), GGML_TYPE_F16, a.view(np.uint8).reshape(-1))


def quantize_q8_0(w: np.ndarray) -> np.ndarray:
    """Row-wise Q8_0 of a float32 [rows, K] matrix -> raw bytes.

    Restates `quantize_row_q8_0_ref` (reference ggml/src/ggml-quants.c:194-217):
    per 32-block amax, d = amax/127 (stored fp16), q = roundf(x * (1/d)) with
    round-half-away-from-zero.
    """
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    assert k % QK8_0 == 0
    blk = w.reshape(rows, k // QK8_0, QK8_0)
    amax = np.abs(blk).max(axis=2)
    d = (amax / np.float32(127.0)).astype(np.float32)
    inv = np.where(d != 0, np.float32(1.0) / np.where(d != 0, d, 1), np.float32(0.0)).astype(np.float32)
    x0 = blk * inv[..., None]
    q = (np.sign(x0) * np.floor(np.abs(x0) + np.float32(0.5))).astype(np.int8)   # roundf
    out = np.empty((rows, k // QK8_0, Q8_0_BLOCK_BYTES), dtype=np.uint8)
    out[..., 0:2] = d.astype(np.float16)[..., None].view(np.uint8)
    out[..., 2:] = q.view(np.uint8)
    return out.reshape(-1)


def q8_0_tensor(name: str, w: np.ndarray) -> GGUFTensor:
    a = np.ascontiguousarray(w, dtype=np.float32)
    assert a.ndim == 2
    return GGUFTensor(name, (a.shape[1], a.shape[0]), GGML_TYPE_Q8_0, quantize_q8_0(a))


def dequantize_q8_0(raw: np.ndarray, ne) -> np.ndarray:
    """raw Q8_0 bytes -> float32 array in numpy order (reference ggml-quants.c dequantize_row_q8_0)."""
    k = ne[0]
    rows = 1
    for d in ne[1:]:
        rows *= d
    b = raw.reshape(rows, k // QK8_0, Q8_0_BLOCK_BYTES)
    d = b[..., 0:2].copy().view(np.float16).astype(np.float32)        # [rows, nb, 1]
    q = b[..., 2:].view(np.int8).astype(np.float32)
    return (q * d).reshape(tuple(reversed(ne)))


def dequantize_legacy(raw: np.ndarray, ne, ggml_type: int, split: bool = False):
    """raw q4_0 / q4_1 / q5_0 / q5_1 bytes -> float32 in numpy order, restating dequantize_row_q4_0 .. q5_1 (reference
    ggml-quants.c:255-339): element j < 16 of a block comes from the low nibble of qs[j], element j + 16 from the high
    nibble; q5 adds bit j (resp. j + 16) of the 32-bit qh as the fifth bit; q*_0: (q - 8 | 16) * d, q*_1: q * d + m.
    split=True returns (q * d, m per block) instead, which the oracle needs for the q8_1 dot product."""
    k = ne[0]
    rows = 1
    for d in ne[1:]:
        rows *= d
    bb = QUANT_BLOCK_BYTES[ggml_type]
    b = raw.reshape(rows, k // 32, bb)
    d = b[..., 0:2].copy().view(np.float16).astype(np.float32)                     # [rows, nb, 1]
    off = 2
    m = None
    if ggml_type in (GGML_TYPE_Q4_1, GGML_TYPE_Q5_1):
        m = b[..., 2:4].copy().view(np.float16).astype(np.float32)
        off = 4
    hi_bits = None
    if ggml_type in (GGML_TYPE_Q5_0, GGML_TYPE_Q5_1):
        qh = b[..., off:off + 4].copy().view(np.uint32)                            # [rows, nb, 1]
        j = np.arange(32, dtype=np.uint32)
        hi_bits = ((qh >> j) & 1).astype(np.int32) << 4                            # bit j -> element j (j < 16: low half)
        off += 4
    qs = b[..., off:off + 16]
    q = np.concatenate([qs & 0x0F, qs >> 4], axis=-1).astype(np.int32)              # elements 0..15, 16..31
    if hi_bits is not None:
        q = q | hi_bits
    if ggml_type == GGML_TYPE_Q4_0:
        q = q - 8
    elif ggml_type == GGML_TYPE_Q5_0:
        q = q - 16
    qd = q.astype(np.float32) * d
    shape = tuple(reversed(ne))
    if split:
        return qd.reshape(shape), (m if m is not None else np.zeros_like(d)).reshape(rows, k // 32)
    return (qd + m if m is not None else qd).astype(np.float32).reshape(shape)


def to_numpy(t: GGUFTensor) -> np.ndarray:
    """Decode to float32 (F32/Q8_0) or float16 (F16), numpy order (reverse of ggml ne)."""
    shape = tuple(reversed(t.ne))
    if t.ggml_type == GGML_TYPE_F32:
        return t.data.view(np.float32).reshape(shape)
    if t.ggml_type == GGML_TYPE_F16:
        return t.data.view(np.float16).reshape(shape)
    if t.ggml_type == GGML_TYPE_Q8_0:
        return dequantize_q8_0(t.data, t.ne)
    if t.ggml_type in QUANT_BLOCK_BYTES:
        return dequantize_legacy(t.data, t.ne, t.ggml_type)
    raise ValueError(t.ggml_type)
