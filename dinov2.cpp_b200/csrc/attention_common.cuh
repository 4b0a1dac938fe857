// Constants shared by the attention kernels and the engine: every DINOv2 size has head_dim 64 (reference dinov2.cpp:479-494),
// keys / values are streamed in tiles of 128 tokens.
#pragma once

namespace dino {

constexpr int ATT_HD = 64;     // head dimension
constexpr int ATT_BKV = 128;   // K/V (and Q) tile rows = TMA box rows of the QKV tensor map

}  // namespace dino
