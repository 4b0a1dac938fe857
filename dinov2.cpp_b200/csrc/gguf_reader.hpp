// Stand-alone GGUF v2/v3 reader for DINOv2 checkpoints (no ggml dependency).
// Replaces gguf_init_from_file + the KV/tensor walk in the reference's dino_model_load
// (reference dinov2.cpp:263-336; container layout: reference ggml/include/gguf.h:1-46).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace dino {

struct GGUFTensorInfo {
    std::string name;
    int32_t type = 0;
    int32_t n_dims = 0;
    int64_t ne[4] = {1, 1, 1, 1};
    uint64_t offset = 0;   // relative to the data section
    uint64_t nbytes = 0;
    const uint8_t *data = nullptr;
};

struct GGUFFile {
    std::vector<uint8_t> blob;                       // whole file
    std::map<std::string, uint64_t> kv_u;            // integer-valued KVs (widened)
    std::map<std::string, double> kv_f;
    std::map<std::string, std::string> kv_s;
    std::vector<GGUFTensorInfo> tensors;             // file order
    struct RawKV { std::string key; size_t off, len; };   // byte span of the whole record (key, type, value) in blob
    std::vector<RawKV> kv_raw;                       // file order (the quantiser copies them verbatim)
    uint32_t version = 0;
    uint64_t alignment = 32;

    const GGUFTensorInfo *find(const std::string &n) const {
        for (const auto &t : tensors)
            if (t.name == n) return &t;
        return nullptr;
    }
};

namespace gguf_detail {
struct Cursor {
    const uint8_t *p;
    const uint8_t *end;
    template <typename T> T take() {
        if (static_cast<size_t>(end - p) < sizeof(T)) throw std::runtime_error("gguf: truncated file");
        T v;
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        const uint64_t n = take<uint64_t>();
        if (n > static_cast<uint64_t>(end - p)) throw std::runtime_error("gguf: truncated string");
        std::string s(reinterpret_cast<const char *>(p), n);
        p += n;
        return s;
    }
};

inline uint64_t type_row_bytes(int32_t type, int64_t ne0) {
    switch (type) {
        case 0: return static_cast<uint64_t>(ne0) * 4;          // F32
        case 1: return static_cast<uint64_t>(ne0) * 2;          // F16
        case 2: case 3: case 6: case 7: case 8: {                // Q4_0 / Q4_1 / Q5_0 / Q5_1 / Q8_0: blocks of 32 elements
            static const int kBlockBytes[9] = {0, 0, 18, 20, 0, 0, 22, 24, 34};
            if (ne0 % 32) throw std::runtime_error("gguf: quantised row not a multiple of 32");
            return static_cast<uint64_t>(ne0) / 32 * kBlockBytes[type];
        }
        default: throw std::runtime_error("gguf: unsupported tensor type " + std::to_string(type) +
                                          " (engine handles F32, F16, Q4_0, Q4_1, Q5_0, Q5_1, Q8_0)");
    }
}
}  // namespace gguf_detail

inline void gguf_read(const std::string &path, GGUFFile &out) {
    using namespace gguf_detail;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    std::fseek(f, 0, SEEK_END);
    const long sz = std::ftell(f);
    if (sz < 0) {
        std::fclose(f);
        throw std::runtime_error("cannot open '" + path + "': not a regular file");
    }
    std::fseek(f, 0, SEEK_SET);
    out.blob.resize(static_cast<size_t>(sz));
    const size_t got = std::fread(out.blob.data(), 1, out.blob.size(), f);
    std::fclose(f);
    if (got != out.blob.size()) throw std::runtime_error("short read on '" + path + "'");

    Cursor c{out.blob.data(), out.blob.data() + out.blob.size()};
    if (c.take<uint32_t>() != 0x46554747u) throw std::runtime_error("gguf: bad magic");
    const uint32_t version = c.take<uint32_t>();
    if (version != 2 && version != 3) throw std::runtime_error("gguf: unsupported version");
    out.version = version;
    const uint64_t n_tensors = c.take<uint64_t>();
    const uint64_t n_kv = c.take<uint64_t>();
    // every KV record takes >= 12 bytes (key length, type, one value byte) and every tensor record >= 32: counts that cannot
    // fit in the file are rejected before anything is sized by them
    const uint64_t remaining = static_cast<uint64_t>(c.end - c.p);
    if (n_kv > remaining / 12 || n_tensors > remaining / 32) throw std::runtime_error("gguf: tensor / KV count exceeds the file size");

    auto skip_or_read = [&](auto &&self, const std::string &key, uint32_t t, bool store) -> void {
        switch (t) {
            case 0: { auto v = c.take<uint8_t>(); if (store) out.kv_u[key] = v; break; }
            case 1: { auto v = c.take<int8_t>(); if (store) out.kv_u[key] = static_cast<uint64_t>(v); break; }
            case 2: { auto v = c.take<uint16_t>(); if (store) out.kv_u[key] = v; break; }
            case 3: { auto v = c.take<int16_t>(); if (store) out.kv_u[key] = static_cast<uint64_t>(v); break; }
            case 4: { auto v = c.take<uint32_t>(); if (store) out.kv_u[key] = v; break; }
            case 5: { auto v = c.take<int32_t>(); if (store) out.kv_u[key] = static_cast<uint64_t>(v); break; }
            case 6: { auto v = c.take<float>(); if (store) out.kv_f[key] = v; break; }
            case 7: { auto v = c.take<uint8_t>(); if (store) out.kv_u[key] = v; break; }
            case 8: { auto v = c.str(); if (store) out.kv_s[key] = v; break; }
            case 9: {
                const uint32_t et = c.take<uint32_t>();
                const uint64_t n = c.take<uint64_t>();
                if (et == 9) throw std::runtime_error("gguf: nested arrays are not supported");
                if (n > static_cast<uint64_t>(c.end - c.p)) throw std::runtime_error("gguf: array length exceeds the file size");
                for (uint64_t i = 0; i < n; ++i) self(self, key, et, false);
                break;
            }
            case 10: { auto v = c.take<uint64_t>(); if (store) out.kv_u[key] = v; break; }
            case 11: { auto v = c.take<int64_t>(); if (store) out.kv_u[key] = static_cast<uint64_t>(v); break; }
            case 12: { auto v = c.take<double>(); if (store) out.kv_f[key] = v; break; }
            default: throw std::runtime_error("gguf: unknown KV type");
        }
    };
    for (uint64_t i = 0; i < n_kv; ++i) {
        const uint8_t *rec = c.p;
        const std::string key = c.str();
        const uint32_t t = c.take<uint32_t>();
        skip_or_read(skip_or_read, key, t, true);
        out.kv_raw.push_back({key, static_cast<size_t>(rec - out.blob.data()), static_cast<size_t>(c.p - rec)});
    }
    out.tensors.resize(n_tensors);
    for (auto &t : out.tensors) {
        t.name = c.str();
        t.n_dims = static_cast<int32_t>(c.take<uint32_t>());
        if (t.n_dims < 1 || t.n_dims > 4) throw std::runtime_error("gguf: bad n_dims for " + t.name);
        // dimensions: positive, and small enough that no product below can wrap (the whole file is < 2^63 bytes)
        constexpr uint64_t kMaxDim = 1ull << 40;
        for (int d = 0; d < t.n_dims; ++d) {
            const uint64_t v = c.take<uint64_t>();
            if (v == 0 || v > kMaxDim) throw std::runtime_error("gguf: bad dimension for " + t.name);
            t.ne[d] = static_cast<int64_t>(v);
        }
        t.type = static_cast<int32_t>(c.take<uint32_t>());
        t.offset = c.take<uint64_t>();
        uint64_t rows = 1;
        for (int d = 1; d < t.n_dims; ++d) {
            if (rows > out.blob.size() / static_cast<uint64_t>(t.ne[d]) + 1) throw std::runtime_error("gguf: tensor larger than the file: " + t.name);
            rows *= static_cast<uint64_t>(t.ne[d]);
        }
        const uint64_t row_bytes = type_row_bytes(t.type, t.ne[0]);
        if (row_bytes == 0 || rows > out.blob.size() / row_bytes) throw std::runtime_error("gguf: tensor larger than the file: " + t.name);
        t.nbytes = row_bytes * rows;
    }
    uint64_t align = 32;
    auto it = out.kv_u.find("general.alignment");
    if (it != out.kv_u.end() && it->second) align = it->second;
    if (align > (1u << 20) || (align & (align - 1))) throw std::runtime_error("gguf: general.alignment is not a power of two <= 1 MiB");
    out.alignment = align;
    const uint64_t size = out.blob.size();
    const uint64_t meta = static_cast<uint64_t>(c.p - out.blob.data());
    const uint64_t data_start = (meta + align - 1) / align * align;
    if (data_start > size && !out.tensors.empty()) throw std::runtime_error("gguf: truncated file (no data section)");
    for (auto &t : out.tensors) {
        // overflow-safe: offset <= size - data_start, nbytes <= size - data_start - offset
        if (t.offset > size - data_start || t.nbytes > size - data_start - t.offset)
            throw std::runtime_error("gguf: tensor data out of range: " + t.name);
        if (t.offset % align) throw std::runtime_error("gguf: tensor data is not aligned: " + t.name);
        t.data = out.blob.data() + data_start + t.offset;
    }
}

}  // namespace dino
