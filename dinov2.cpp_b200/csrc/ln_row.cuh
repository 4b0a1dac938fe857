// LayerNorm of one token row by one warp (reference ggml_norm, ops.cpp:3109-3158: mean, then centred variance,
// y = (x-mean)/sqrt(var+eps); followed by *weight + bias, dinov2.cpp:694-700).  Shared by the stand-alone LayerNorm
// kernel (elementwise.cuh) and the fused residual-GEMM epilogue (gemm.cuh) so that both produce identical bits.
// OUT_HALF: writes the fp16 A operand of the next GEMM (the reference rounds it to fp16 inside mul_mat).
// L2_ONLY: the row was just produced by another SM (TMA reduce-add): load with ld.global.cg, never from L1.
#pragma once
#include "ptx.cuh"

namespace dino {

constexpr int LN_MAX_V4 = 12;   // D up to 32 * 4 * 12 = 1536

// NV4 = float4 per lane the row buffer is sized for (D <= 128 * NV4): sizing it to the model width instead of the maximum
// frees registers -> more rows in flight per SM (the kernel is latency-bound per row: load, two shuffle reductions, store).
template <bool OUT_HALF, bool L2_ONLY, int NV4 = LN_MAX_V4>
__device__ __forceinline__ void layernorm_row(const float *__restrict__ xrow, const float *__restrict__ gamma,
                                              const float *__restrict__ beta, void *__restrict__ orow, int D, float eps, int lane) {
    const int nv = D >> 2;   // float4 per row
    const float4 *x4 = reinterpret_cast<const float4 *>(xrow);
    float4 v[NV4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
            v[i] = L2_ONLY ? __ldcg(x4 + idx) : x4[idx];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq / static_cast<float>(D) + eps);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
            const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
            const float y0 = (v[i].x * rstd) * g.x + b.x, y1 = (v[i].y * rstd) * g.y + b.y;
            const float y2 = (v[i].z * rstd) * g.z + b.z, y3 = (v[i].w * rstd) * g.w + b.w;
            if constexpr (OUT_HALF) {
                reinterpret_cast<uint2 *>(orow)[idx] = make_uint2(pack_half2(y0, y1), pack_half2(y2, y3));
            } else {
                reinterpret_cast<float4 *>(orow)[idx] = make_float4(y0, y1, y2, y3);
            }
        }
    }
}

// R rows at once, same arithmetic per row as layernorm_row<true, true> (bit-identical fp16 output): the loads of all R rows are
// issued before any reduction, so a warp that reads its rows from L2 (latency 1-2 us under load) keeps R x D x 4 bytes in
// flight.  The row width is EXACTLY 128 * NV4 and nothing is predicated per element: with `if (idx < nv)` around the loads ptxas
// kept the row buffer in local memory and, worse, put a store behind every load — the in-order warp then waited out one L2
// round trip per load (measured 16 k cycles per pair of rows).  Rows are `ld` elements apart (input) / D apart (output); rows
// >= n_rows (warp-uniform) re-read the last valid row and store nothing.
// Used by the LayerNorm worker warps of the residual GEMM (gemm.cuh, EPI_RESID_LN_F32).
// gamma / beta come from shared memory (gs: gamma as float4[NV4 * 32], beta right behind it): the GEMM's shared-memory
// carve-out leaves almost no L1, so per-row global reads of them were a second serial L2 round trip per row and doubled the
// worker's L2 traffic (ncu: 718 MB of gamma / beta sectors per launch).
template <int NV4, int R>
__device__ __forceinline__ void layernorm_rows_l2(const float *__restrict__ x, size_t ld, const float4 *gs,
                                                  __half *__restrict__ out, float eps, int lane, int n_rows) {
    constexpr int D = NV4 * 128;
    float4 v[R][NV4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x + static_cast<size_t>(min(r, n_rows - 1)) * ld) + lane;
#pragma unroll
        for (int i = 0; i < NV4; ++i) v[r][i] = __ldcg(x4 + 32 * i);
    }
    const float4 *g4 = gs + lane;
    const float4 *b4 = gs + NV4 * 32 + lane;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NV4; ++i) sum += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / static_cast<float>(D);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < NV4; ++i) {
            v[r][i].x -= mean; v[r][i].y -= mean; v[r][i].z -= mean; v[r][i].w -= mean;
            sq += (v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y) + (v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = 1.0f / sqrtf(sq / static_cast<float>(D) + eps);
        if (r < n_rows) {
            uint2 *o2 = reinterpret_cast<uint2 *>(out + static_cast<size_t>(r) * D) + lane;
#pragma unroll
            for (int i = 0; i < NV4; ++i) {
                const float4 g = g4[32 * i], b = b4[32 * i];
                o2[32 * i] = make_uint2(pack_half2((v[r][i].x * rstd) * g.x + b.x, (v[r][i].y * rstd) * g.y + b.y),
                                        pack_half2((v[r][i].z * rstd) * g.z + b.z, (v[r][i].w * rstd) * g.w + b.w));
            }
        }
    }
}

// Destinations of the fused final-LayerNorm + all-gather kernel: the same row is stored into the gather buffer of every rank
// (peer device memory over NVLink / NVSwitch, or this device for the rank's own copy).
struct GatherDst {
    float *p[8];
    int n;
};

// Same arithmetic as layernorm_row<false, ...> (bit-identical values), stored to n destinations.
template <int NV4 = LN_MAX_V4>
__device__ __forceinline__ void layernorm_row_multi(const float *__restrict__ xrow, const float *__restrict__ gamma,
                                                    const float *__restrict__ beta, const GatherDst &dst, size_t dst_off, int D, float eps,
                                                    int lane) {
    const int nv = D >> 2;
    const float4 *x4 = reinterpret_cast<const float4 *>(xrow);
    float4 v[NV4];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
            v[i] = x4[idx];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.0f / sqrtf(sq / static_cast<float>(D) + eps);
    const float4 *g4 = reinterpret_cast<const float4 *>(gamma);
    const float4 *b4 = reinterpret_cast<const float4 *>(beta);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
        const int idx = lane + 32 * i;
        if (idx < nv) {
            const float4 g = __ldg(g4 + idx), b = __ldg(b4 + idx);
            const float4 y = make_float4((v[i].x * rstd) * g.x + b.x, (v[i].y * rstd) * g.y + b.y, (v[i].z * rstd) * g.z + b.z,
                                         (v[i].w * rstd) * g.w + b.w);
            for (int k = 0; k < dst.n; ++k) reinterpret_cast<float4 *>(dst.p[k] + dst_off)[idx] = y;
        }
    }
}

}  // namespace dino
