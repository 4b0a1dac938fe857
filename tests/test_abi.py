"""C-ABI library: loads, exports every symbol include/dinov2_b200.h declares, and fails loudly (no CPU
fallback) when there is no B200.  No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

import dinov2_b200 as d
from dinov2_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "dinov2_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(dino_b200_[a-z_0-9]+)\s*\(", hdr)))
    assert declared == sorted(E.ABI_SYMBOLS), "engine.py's symbol list drifted from the header"
    L = ctypes.CDLL(E.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), f"libdinov2_b200.so does not export {s}"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "dinov2_b200.h"\nint main(void){dino_b200_hparams h; (void)h; return DINO_B200_OK;}\n')
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o",
                        str(tmp_path / "t.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback_without_gpu():
    if d.device_count() > 0:
        pytest.skip("a B200 is present")
    with pytest.raises(d.DinoB200Error) as ei:
        d.Engine(os.path.join(ROOT, "tests", "golden", "tiny_f16.gguf"))
    assert ei.value.status == 6          # DINO_B200_ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_bad_arguments_are_rejected_not_crashed():
    L = d.load_library()
    assert L.dino_b200_create_from_gguf(None, 0, None) == 1
    assert L.dino_b200_forward(None, None, 0, 1, 14, 14, 0, None, None, None, None) == 1
    assert L.dino_b200_kernel_launches(None) == 0
    assert L.dino_b200_last_error(None) is not None


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dinov2.cpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "restate" not in txt and "oracle/" not in txt.replace("oracle/Makefile", "").replace("oracle/_ref", "") or f == "core.hpp", f


def test_malformed_checkpoints_are_reported_before_any_device_work(tmp_path):
    """The engine's own GGUF reader (csrc/gguf_reader.hpp) rejects bad files with the documented status codes — on any
    machine, because the file is parsed before a device is required (reference: gguf_init_from_file failing in
    dino_model_load, dinov2.cpp:263-270)."""
    good = open(os.path.join(ROOT, "tests", "golden", "tiny_f16.gguf"), "rb").read()
    cases = {
        "missing.gguf": (None, 2),                                   # DINO_B200_ERR_IO
        "bad_magic.gguf": (b"GGML" + good[4:], 3),                    # DINO_B200_ERR_FORMAT
        "truncated_meta.gguf": (good[:200], 3),
        "truncated_data.gguf": (good[: len(good) // 2], 3),
        "bad_version.gguf": (good[:4] + (99).to_bytes(4, "little") + good[8:], 3),
    }
    for name, (blob, want) in cases.items():
        path = tmp_path / name
        if blob is not None:
            path.write_bytes(blob)
        with pytest.raises(d.DinoB200Error) as ei:
            d.Engine(str(path))
        assert ei.value.status == want, (name, ei.value.status, str(ei.value))
        assert str(ei.value)                                          # a message, not just a code


def _tensor_info_fields(blob: bytes):
    """Byte offsets of every tensor-info field in a GGUF file: [(name, [ne offsets], type offset, data-offset offset)] and
    the offset of the header's tensor count; plus {key: value offset} of the scalar KVs."""
    import struct
    pos = 8
    n_tensors, n_kv = struct.unpack_from("<QQ", blob, pos)
    pos += 16
    sizes = {0: 1, 1: 1, 2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 7: 1, 10: 8, 11: 8, 12: 8}
    kv_off = {}
    for _ in range(n_kv):
        n = struct.unpack_from("<Q", blob, pos)[0]
        key = blob[pos + 8: pos + 8 + n].decode()
        pos += 8 + n
        t = struct.unpack_from("<I", blob, pos)[0]
        pos += 4
        kv_off[key] = pos
        if t == 8:
            pos += 8 + struct.unpack_from("<Q", blob, pos)[0]
        else:
            pos += sizes[t]
    infos = []
    for _ in range(n_tensors):
        n = struct.unpack_from("<Q", blob, pos)[0]
        name = blob[pos + 8: pos + 8 + n].decode()
        pos += 8 + n
        nd = struct.unpack_from("<I", blob, pos)[0]
        pos += 4
        ne = [pos + 8 * i for i in range(nd)]
        pos += 8 * nd
        infos.append((name, ne, pos, pos + 4))
        pos += 12
    return infos, kv_off


def test_hostile_tensor_tables_are_rejected(tmp_path):
    """Overflow-safe range checks of the GGUF reader (round-1 advisor finding): a data offset that wraps in uint64, zero
    or absurd dimensions, counts larger than the file, a misaligned offset and a class count without a classifier head are
    all format errors — never an out-of-bounds upload, a division by zero or an exception across the C ABI."""
    import struct
    good = bytearray(open(os.path.join(ROOT, "tests", "golden", "tiny_f16.gguf"), "rb").read())
    infos, kv_off = _tensor_info_fields(bytes(good))
    by_name = {i[0]: i for i in infos}
    qkv = by_name["encoder.layer.0.attention.attention.qkv.weight"]
    cases = {}

    b = bytearray(good); struct.pack_into("<Q", b, qkv[3], 2**64 - 4096); cases["wrapping_offset"] = b
    b = bytearray(good); struct.pack_into("<Q", b, qkv[3], len(good)); cases["offset_past_end"] = b
    b = bytearray(good); struct.pack_into("<Q", b, qkv[3], struct.unpack_from("<Q", good, qkv[3])[0] + 2); cases["misaligned_offset"] = b
    b = bytearray(good); struct.pack_into("<Q", b, qkv[1][0], 0); cases["zero_dimension"] = b
    b = bytearray(good); struct.pack_into("<Q", b, qkv[1][1], 2**63); cases["huge_dimension"] = b
    b = bytearray(good); struct.pack_into("<Q", b, qkv[1][1], 2**33); cases["rows_overflow"] = b
    b = bytearray(good); struct.pack_into("<Q", b, 8, 2**60); cases["tensor_count_overflow"] = b
    b = bytearray(good); struct.pack_into("<Q", b, 16, 2**60); cases["kv_count_overflow"] = b
    b = bytearray(good); struct.pack_into("<I", b, kv_off["num_classes"], 0xFFFFFFFF); cases["class_count_without_head"] = b
    for name, blob in cases.items():
        path = tmp_path / (name + ".gguf")
        path.write_bytes(bytes(blob))
        with pytest.raises(d.DinoB200Error) as ei:
            d.Engine(str(path))
        assert ei.value.status == 3, (name, ei.value.status, str(ei.value))       # DINO_B200_ERR_FORMAT
        assert str(ei.value), name
