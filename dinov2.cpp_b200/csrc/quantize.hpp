// Offline weight quantiser: the reference's `quantize` tool / dino_model_quantize (reference dinov2.cpp:354-452) without ggml.
// Host-only (no CUDA): reads a gguf, re-encodes every 2-D tensor whose name ends in "weight" (do_quantize, dinov2.cpp:227-236)
// as q4_0 / q4_1 / q5_0 / q5_1 / q8_0 with the deterministic reference quantisers (quantize_row_*_ref,
// ggml-quants.c:30-217), copies everything else, sets the "ftype" key, and writes a gguf v3 file laid out the way
// gguf_write_to_file does (KVs in order with the re-set "ftype" moved to the end, tensor data 32-byte aligned).
#pragma once
#include "gguf_reader.hpp"

#include <cfloat>
#include <cmath>
#include <cuda_fp16.h>

namespace dino {
namespace quant_detail {

inline uint16_t f2h(float f) {
    const __half h = __float2half_rn(f);           // IEEE round-to-nearest-even, as GGML_FP32_TO_FP16 (F16C)
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
}
inline float h2f(uint16_t u) {
    __half h;
    std::memcpy(&h, &u, 2);
    return __half2float(h);
}
inline void put16(uint8_t *p, uint16_t v) { std::memcpy(p, &v, 2); }

// one 32-element block each; layouts in ggml-common.h:170-213
inline void q4_0(const float *x, uint8_t *y) {
    float amax = 0.f, mx = 0.f;
    for (int j = 0; j < 32; ++j)
        if (amax < std::fabs(x[j])) { amax = std::fabs(x[j]); mx = x[j]; }
    const float d = mx / -8, id = d ? 1.0f / d : 0.0f;
    put16(y, f2h(d));
    for (int j = 0; j < 16; ++j) {
        const uint8_t a = static_cast<uint8_t>(std::min(15, static_cast<int>(static_cast<int8_t>(x[j] * id + 8.5f))));
        const uint8_t b = static_cast<uint8_t>(std::min(15, static_cast<int>(static_cast<int8_t>(x[j + 16] * id + 8.5f))));
        y[2 + j] = static_cast<uint8_t>(a | (b << 4));
    }
}
inline void q4_1(const float *x, uint8_t *y) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int j = 0; j < 32; ++j) { mn = std::min(mn, x[j]); mx = std::max(mx, x[j]); }
    const float d = (mx - mn) / 15, id = d ? 1.0f / d : 0.0f;
    put16(y, f2h(d));
    put16(y + 2, f2h(mn));
    for (int j = 0; j < 16; ++j) {
        const uint8_t a = static_cast<uint8_t>(std::min(15, static_cast<int>(static_cast<int8_t>((x[j] - mn) * id + 0.5f))));
        const uint8_t b = static_cast<uint8_t>(std::min(15, static_cast<int>(static_cast<int8_t>((x[j + 16] - mn) * id + 0.5f))));
        y[4 + j] = static_cast<uint8_t>(a | (b << 4));
    }
}
inline void q5_0(const float *x, uint8_t *y) {
    float amax = 0.f, mx = 0.f;
    for (int j = 0; j < 32; ++j)
        if (amax < std::fabs(x[j])) { amax = std::fabs(x[j]); mx = x[j]; }
    const float d = mx / -16, id = d ? 1.0f / d : 0.0f;
    put16(y, f2h(d));
    uint32_t qh = 0;
    for (int j = 0; j < 16; ++j) {
        const uint8_t a = static_cast<uint8_t>(std::min(31, static_cast<int>(static_cast<int8_t>(x[j] * id + 16.5f))));
        const uint8_t b = static_cast<uint8_t>(std::min(31, static_cast<int>(static_cast<int8_t>(x[j + 16] * id + 16.5f))));
        y[6 + j] = static_cast<uint8_t>((a & 0x0F) | ((b & 0x0F) << 4));
        qh |= static_cast<uint32_t>((a & 0x10u) >> 4) << j;
        qh |= static_cast<uint32_t>((b & 0x10u) >> 4) << (j + 16);
    }
    std::memcpy(y + 2, &qh, 4);
}
inline void q5_1(const float *x, uint8_t *y) {
    float mn = FLT_MAX, mx = -FLT_MAX;
    for (int j = 0; j < 32; ++j) { mn = std::min(mn, x[j]); mx = std::max(mx, x[j]); }
    const float d = (mx - mn) / 31, id = d ? 1.0f / d : 0.0f;
    put16(y, f2h(d));
    put16(y + 2, f2h(mn));
    uint32_t qh = 0;
    for (int j = 0; j < 16; ++j) {
        const uint8_t a = static_cast<uint8_t>((x[j] - mn) * id + 0.5f);
        const uint8_t b = static_cast<uint8_t>((x[j + 16] - mn) * id + 0.5f);
        y[8 + j] = static_cast<uint8_t>((a & 0x0F) | ((b & 0x0F) << 4));
        qh |= static_cast<uint32_t>((a & 0x10u) >> 4) << j;
        qh |= static_cast<uint32_t>((b & 0x10u) >> 4) << (j + 16);
    }
    std::memcpy(y + 4, &qh, 4);
}
inline void q8_0(const float *x, uint8_t *y) {
    float amax = 0.f;
    for (int j = 0; j < 32; ++j) amax = std::max(amax, std::fabs(x[j]));
    const float d = amax / 127, id = d ? 1.0f / d : 0.0f;
    put16(y, f2h(d));
    for (int j = 0; j < 32; ++j) y[2 + j] = static_cast<uint8_t>(static_cast<int8_t>(std::roundf(x[j] * id)));
}

struct Writer {
    std::vector<uint8_t> b;
    template <typename T> void put(T v) {
        const size_t n = b.size();
        b.resize(n + sizeof(T));
        std::memcpy(b.data() + n, &v, sizeof(T));
    }
    void str(const std::string &s) {
        put<uint64_t>(s.size());
        b.insert(b.end(), s.begin(), s.end());
    }
    void raw(const uint8_t *p, size_t n) { b.insert(b.end(), p, p + n); }
    void pad(size_t align) { b.resize((b.size() + align - 1) / align * align, 0); }
};

}  // namespace quant_detail

// returns the number of tensors that were quantised; throws std::runtime_error on malformed input / unsupported type
inline int quantize_gguf(const std::string &fname_inp, const std::string &fname_out, int itype) {
    using namespace quant_detail;
    static const int kBlockBytes[9] = {0, 0, 18, 20, 0, 0, 22, 24, 34};
    if (!(itype == 2 || itype == 3 || itype == 6 || itype == 7 || itype == 8))
        throw std::runtime_error("quantize: target type must be 2 (q4_0), 3 (q4_1), 6 (q5_0), 7 (q5_1) or 8 (q8_0)");
    GGUFFile in;
    gguf_read(fname_inp, in);

    std::vector<std::vector<uint8_t>> data(in.tensors.size());
    std::vector<int32_t> types(in.tensors.size());
    int n_quant = 0;
    for (size_t i = 0; i < in.tensors.size(); ++i) {
        const GGUFTensorInfo &t = in.tensors[i];
        int dims = 4;
        while (dims > 1 && t.ne[dims - 1] == 1) --dims;          // ggml_n_dims
        const size_t n = t.name.size();
        const bool want = dims == 2 && n >= 6 && t.name.compare(n - 6, 6, "weight") == 0;   // regex ".*weight", 2-D only
        if (!want) {
            types[i] = t.type;
            data[i].assign(t.data, t.data + t.nbytes);
            continue;
        }
        if (t.type != 0 && t.type != 1) throw std::runtime_error("quantize: tensor '" + t.name + "' is neither F32 nor F16");
        if (t.ne[0] % 32) throw std::runtime_error("quantize: row length of '" + t.name + "' is not a multiple of 32");
        const int64_t rows = t.ne[1], k = t.ne[0], nb = k / 32;
        std::vector<float> row(static_cast<size_t>(k));
        data[i].resize(static_cast<size_t>(rows * nb) * kBlockBytes[itype]);
        for (int64_t r = 0; r < rows; ++r) {
            if (t.type == 1) {
                const uint16_t *src = reinterpret_cast<const uint16_t *>(t.data) + r * k;
                for (int64_t j = 0; j < k; ++j) {
                    uint16_t u;
                    std::memcpy(&u, src + j, 2);
                    row[static_cast<size_t>(j)] = h2f(u);
                }
            } else {
                std::memcpy(row.data(), reinterpret_cast<const float *>(t.data) + r * k, static_cast<size_t>(k) * 4);
            }
            uint8_t *dst = data[i].data() + static_cast<size_t>(r * nb) * kBlockBytes[itype];
            for (int64_t bidx = 0; bidx < nb; ++bidx) {
                const float *x = row.data() + bidx * 32;
                uint8_t *y = dst + bidx * kBlockBytes[itype];
                switch (itype) {
                    case 2: q4_0(x, y); break;
                    case 3: q4_1(x, y); break;
                    case 6: q5_0(x, y); break;
                    case 7: q5_1(x, y); break;
                    default: q8_0(x, y); break;
                }
            }
        }
        types[i] = itype;
        ++n_quant;
    }

    Writer w;
    w.put<uint32_t>(0x46554747u);                               // "GGUF"
    w.put<uint32_t>(3);
    w.put<uint64_t>(in.tensors.size());
    bool had_ftype = false;
    for (const auto &kv : in.kv_raw) had_ftype |= kv.key == "ftype";
    w.put<uint64_t>(in.kv_raw.size() + (had_ftype ? 0 : 1));
    for (const auto &kv : in.kv_raw)
        if (kv.key != "ftype") w.raw(in.blob.data() + kv.off, kv.len);
    w.str("ftype");                                             // gguf_set_val_u32 removes the key and appends it again
    w.put<uint32_t>(4);                                         // GGUF_TYPE_UINT32
    w.put<uint32_t>(static_cast<uint32_t>(itype));
    const size_t align = static_cast<size_t>(in.alignment);
    uint64_t off = 0;
    for (size_t i = 0; i < in.tensors.size(); ++i) {
        const GGUFTensorInfo &t = in.tensors[i];
        w.str(t.name);
        int nd = 4;                                             // gguf_add_tensor stores ggml_n_dims(): trailing unit dims dropped
        while (nd > 1 && t.ne[nd - 1] == 1) --nd;
        w.put<uint32_t>(static_cast<uint32_t>(nd));
        for (int d = 0; d < nd; ++d) w.put<uint64_t>(static_cast<uint64_t>(t.ne[d]));
        w.put<uint32_t>(static_cast<uint32_t>(types[i]));
        w.put<uint64_t>(off);
        off += (data[i].size() + align - 1) / align * align;
    }
    w.pad(align);
    for (size_t i = 0; i < in.tensors.size(); ++i) {
        w.raw(data[i].data(), data[i].size());
        w.pad(align);
    }
    FILE *f = std::fopen(fname_out.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot open '" + fname_out + "' for writing");
    const size_t put = std::fwrite(w.b.data(), 1, w.b.size(), f);
    std::fclose(f);
    if (put != w.b.size()) throw std::runtime_error("short write on '" + fname_out + "'");
    return n_quant;
}

}  // namespace dino
