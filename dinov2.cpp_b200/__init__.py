"""dinov2.cpp_b200 — B200-native DINOv2 forward engine behind the dinov2.cpp API.

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the extern "C" ABI (include/dinov2_b200.h)
  host/      C++ drop-in for the reference's dinov2.h API on top of the C ABI
  engine.py  ctypes binding of the C ABI (tests, bench)
  gguf_io.py GGUF reader/writer, synth.py deterministic synthetic checkpoints / inputs
"""
from . import gguf_io, synth  # noqa: F401
from .engine import (Engine, Group, GATHER_CLS, GATHER_PATCH, DinoB200Error, device_count, load_library, quantize_gguf, LAYOUT_BGR_HWC, LAYOUT_RGB_PLANAR,  # noqa: F401
                     LIB_PATH)
