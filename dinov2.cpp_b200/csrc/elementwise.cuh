// Bandwidth-bound kernels around the tensor-core pipeline: im2col, prefix tokens, LayerNorm,
// classifier head, q8_0 dequantisation, bicubic pos-embed resampling.  All are sized for HBM
// streaming (16-byte vector accesses, one warp per token row, grids that cover every SM).
#pragma once
#include "ptx.cuh"
#include "ln_row.cuh"

namespace dino {

// ---------------------------------------------------------------------------------------------
// im2col for the 14x14 / stride-14 patch embedding (reference ggml_conv_2d -> im2col_f16,
// ggml.c:3995-4017, ops.cpp:5784-5855): A[b*np + y*gw + x][c*ps*ps + ky*ps + kx] = fp16(img[c][y*ps+ky][x*ps+kx]),
// zero-padded from 588 to `kpad` columns so the GEMM's 64-wide K boxes need no tail handling.
// Input either RGB planar [B,3,H,W] (layout 0, what the reference uploads, dinov2.cpp:914-933) or the
// cv::Mat layout [B,H,W,3] BGR-interleaved (layout 1) so the host never has to transpose.
// One thread per input PIXEL (consecutive threads read consecutive pixels: fully coalesced loads of the 3.2 MB / image that
// dominate the traffic; the 2-byte stores of 14 neighbouring threads form one contiguous 28-byte run per channel).  Pixels
// outside the patch grid (H or W not a multiple of ps never reaches here) do not exist; the zero pad columns 588..kpad-1 of
// every patch row are written by the threads of the patch's first pixel row.
__global__ void im2col_patch14_kernel(const float *__restrict__ img, __half *__restrict__ A, int B, int H, int W, int ps,
                                      int gh, int gw, int kpad, int layout) {
    griddep_wait();
    griddep_launch_dependents();
    const int HH = gh * ps, WW = gw * ps;                      // the part of the image covered by patches
    const long long total = static_cast<long long>(B) * HH * WW;
    const int kreal = 3 * ps * ps;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int px = static_cast<int>(idx % WW);
        const int py = static_cast<int>((idx / WW) % HH);
        const int b = static_cast<int>(idx / (static_cast<long long>(WW) * HH));
        const int x = px / ps, kx = px - x * ps, y = py / ps, ky = py - y * ps;
        __half *dst = A + (static_cast<long long>(b) * gh * gw + static_cast<long long>(y) * gw + x) * kpad + ky * ps + kx;
        float r, g, bl;
        if (layout == 0) {                                     // RGB planar [B,3,H,W]
            const float *src = img + (static_cast<size_t>(b) * 3 * H + py) * W + px;
            r = __ldg(src);
            g = __ldg(src + static_cast<size_t>(H) * W);
            bl = __ldg(src + 2 * static_cast<size_t>(H) * W);
        } else {                                               // cv::Mat [B,H,W,3], BGR
            const float *src = img + ((static_cast<size_t>(b) * H + py) * W + px) * 3;
            bl = __ldg(src);
            g = __ldg(src + 1);
            r = __ldg(src + 2);
        }
        dst[0] = __float2half_rn(r);
        dst[ps * ps] = __float2half_rn(g);
        dst[2 * ps * ps] = __float2half_rn(bl);
        if (ky == 0) {                                         // this patch's share of the zero padding: kx-th slice of the pad columns
            __half *pad = dst - kx + kreal;                    // = row start + kreal
            for (int k = kx; k < kpad - kreal; k += ps) pad[k] = __float2half_rn(0.f);
        }
    }
}

// cls + pos[0] and the register tokens (no pos-embed) at the head of every image's token block
// (reference dinov2.cpp:669-685).  grid = B, block covers D.
__global__ void prefix_tokens_kernel(float *__restrict__ X, const float *__restrict__ cls, const float *__restrict__ pos,
                                     const float *__restrict__ reg, int ntok, int D, int R) {
    griddep_wait();
    griddep_launch_dependents();
    float *xb = X + static_cast<size_t>(blockIdx.x) * ntok * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        xb[d] = cls[d] + pos[d];
        for (int r = 0; r < R; ++r) xb[(1 + r) * D + d] = reg[r * D + d];
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the hidden dimension, one warp per token (arithmetic in ln_row.cuh).
template <bool OUT_HALF, int NV4 = LN_MAX_V4>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float *__restrict__ X, const float *__restrict__ gamma, const float *__restrict__ beta,
                 void *__restrict__ out, int rows, int D, float eps, int reverse) {
    griddep_wait();
    griddep_launch_dependents();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    // reverse: the first blocks to be scheduled take the LAST rows (what the producer kernel wrote last is still in L2)
    const size_t off = static_cast<size_t>(reverse ? rows - 1 - warp : warp) * D;
    void *orow = OUT_HALF ? static_cast<void *>(reinterpret_cast<__half *>(out) + off) : static_cast<void *>(reinterpret_cast<float *>(out) + off);
    layernorm_row<OUT_HALF, false, NV4>(X + off, gamma, beta, orow, D, eps, lane);
}

// Final LayerNorm fused with the feature all-gather (the one optional exchange of the data-parallel path):
// one warp per exported token row — the class token (rows_per_image = 1) or the patch tokens (rows_per_image = NP, registers
// stripped) of each of this rank's images — normalises the row of the residual stream (dinov2.cpp:752-761) and stores it into
// the gather buffer of EVERY rank at [rank_slot + image][row][D]: peer stores over NVLink / NVSwitch, no separate collective,
// no staging copy.  Visibility on the peers: kernel completion + the caller's cross-rank synchronisation.
template <int NV4 = LN_MAX_V4>
__global__ void __launch_bounds__(256)
layernorm_gather_kernel(const float *__restrict__ X, const float *__restrict__ gamma, const float *__restrict__ beta, GatherDst dst,
                        int n_images, int ntok, int tok0, int rows_per_image, size_t slot_row0, int D, float eps) {
    griddep_wait();
    griddep_launch_dependents();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_images * rows_per_image) return;
    const int img = warp / rows_per_image, r = warp - img * rows_per_image;
    const float *xrow = X + (static_cast<size_t>(img) * ntok + tok0 + r) * D;
    layernorm_row_multi<NV4>(xrow, gamma, beta, dst, (slot_row0 + static_cast<size_t>(warp)) * D, D, eps, lane);
}

// ---------------------------------------------------------------------------------------------
// Classifier head (reference forward_head, dinov2.cpp:792-821).
// pooled[b][d] = (sum over tokens 1..ntok-1 of Y[b][t][d]) * (1 / n_embd^2): the divisor is the constant
// (img_size/patch)^2 and the sum includes register tokens (dinov2.cpp:770-776, 800-803); ggml_sum_rows
// accumulates in double.  feat[b] = [cls ; pooled]  ([B, 2D] fp32).
// grid = (ceil(D / 128), B), block = 1024: a block owns 128 channels of one image; its 32 warps take every 32nd token (independent
// 16-byte loads, 8 in flight per thread) and the 32 partial sums of a channel are added in warp order, so the result does not
// depend on timing.  (One thread per channel walking all tokens was 130 us at batch 1 — 1369 dependent L2 round trips.)
constexpr int POOL_THREADS = 1024;
__global__ void __launch_bounds__(POOL_THREADS) pool_tokens_kernel(const float *__restrict__ Y, float *__restrict__ feat, int ntok, int D, float inv_div) {
    griddep_wait();
    griddep_launch_dependents();
    __shared__ double red[4][32][32];                        // [component][warp][lane]
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = (blockIdx.x * 32 + lane) * 4;              // first of this thread's four channels
    const bool live = d < D;                                 // D is a multiple of 4
    const float *y = Y + static_cast<size_t>(b) * ntok * D + d;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (live) {
#pragma unroll 8
        for (int t = 1 + warp; t < ntok; t += 32) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(y + static_cast<size_t>(t) * D));
            a0 += static_cast<double>(v.x); a1 += static_cast<double>(v.y); a2 += static_cast<double>(v.z); a3 += static_cast<double>(v.w);
        }
    }
    red[0][warp][lane] = a0; red[1][warp][lane] = a1; red[2][warp][lane] = a2; red[3][warp][lane] = a3;
    __syncthreads();
    if (warp < 4 && live) {                                  // warp c finishes component c of the block's 32 channel quads
        double acc = 0.0;
#pragma unroll 8
        for (int w = 0; w < 32; ++w) acc += red[warp][w][lane];
        feat[static_cast<size_t>(b) * 2 * D + d + warp] = y[warp];                                   // class token (t = 0)
        feat[static_cast<size_t>(b) * 2 * D + D + d + warp] = static_cast<float>(acc) * inv_div;
    }
}

// logits[b][c] = sum_k fp16(feat[b][k]) * W[c][k] + bias[c]  (fp16 x fp16 -> f32, one warp per (b, c))
__global__ void classifier_kernel(const float *__restrict__ feat, const __half *__restrict__ Wc, const float *__restrict__ bias,
                                  float *__restrict__ logits, int B, int K, int C) {
    griddep_wait();
    griddep_launch_dependents();
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= B * C) return;
    const int b = gw / C, c = gw % C;
    const float *f = feat + static_cast<size_t>(b) * K;
    const __half *w = Wc + static_cast<size_t>(c) * K;
    float acc = 0.f;
    for (int k = lane * 2; k < K; k += 64) {
        const float2 wv = __half22float2(*reinterpret_cast<const __half2 *>(w + k));
        acc = fmaf(__half2float(__float2half_rn(f[k])), wv.x, acc);
        acc = fmaf(__half2float(__float2half_rn(f[k + 1])), wv.y, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[static_cast<size_t>(b) * C + c] = acc + bias[c];
}

// probs = softmax(logits) per image (reference ggml_soft_max, ops.cpp:4641-4737). grid = B, block = 256.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float *__restrict__ logits, float *__restrict__ probs, int C) {
    griddep_wait();
    griddep_launch_dependents();
    __shared__ float red[8];
    __shared__ float bcast;
    const float *x = logits + static_cast<size_t>(blockIdx.x) * C;
    float *y = probs + static_cast<size_t>(blockIdx.x) * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < C; i += blockDim.x) mx = fmaxf(mx, x[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = red[0];
        for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
        bcast = m;
    }
    __syncthreads();
    mx = bcast;
    float sum = 0.f;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        const float e = expf(x[i] - mx);
        y[i] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < 8; ++i) s += red[i];
        bcast = 1.0f / s;
    }
    __syncthreads();
    const float inv = bcast;
    for (int i = threadIdx.x; i < C; i += blockDim.x) y[i] *= inv;
}

// ---------------------------------------------------------------------------------------------
// Load-time weight preparation.
// Quantised weights -> fp16, once at load.  Blocks of 32 elements along K in the reference's layouts (ggml-common.h:170-213,
// dequantize_row_q4_0 .. q8_0 in ggml-quants.c:255-353):
//   q8_0 {half d; int8 q[32]}                    x = q d
//   q4_0 {half d; u8 qs[16]}                     x[j] = ((qs[j] & 15) - 8) d,  x[j+16] = ((qs[j] >> 4) - 8) d
//   q4_1 {half d, m; u8 qs[16]}                  x = q d + m
//   q5_0 {half d; u32 qh; u8 qs[16]}             fifth bit of element j = bit j of qh;  x = (q - 16) d
//   q5_1 {half d, m; u32 qh; u8 qs[16]}          x = q d + m
// computed in fp32 exactly as the reference does, then rounded once to fp16 (the GEMM operand type).  One thread per
// block; `row_map` (optional) permutes output rows (used to interleave SwiGLU gate/up rows per N tile).
__device__ __forceinline__ float dq_half_at(const uint8_t *p) {
    return __half2float(__ushort_as_half(static_cast<unsigned short>(p[0] | (p[1] << 8))));
}
template <int TYPE>
__global__ void dequant_kernel(const uint8_t *__restrict__ raw, __half *__restrict__ W, long long n_blocks, int blocks_per_row,
                               int ld_out, const int *__restrict__ row_map) {
    constexpr int kBytes = TYPE == 2 ? 18 : TYPE == 3 ? 20 : TYPE == 6 ? 22 : TYPE == 7 ? 24 : 34;
    constexpr bool kHasMin = TYPE == 3 || TYPE == 7;
    constexpr bool kFive = TYPE == 6 || TYPE == 7;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_blocks;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint8_t *blk = raw + i * kBytes;
        const float d = dq_half_at(blk);
        const long long row = i / blocks_per_row;
        const int kb = static_cast<int>(i % blocks_per_row);
        const long long orow = row_map ? row_map[row] : row;
        __half *dst = W + orow * ld_out + kb * 32;
        if constexpr (TYPE == 8) {
            for (int j = 0; j < 32; ++j) dst[j] = __float2half_rn(d * static_cast<float>(static_cast<int8_t>(blk[2 + j])));
        } else {
            const float m = kHasMin ? dq_half_at(blk + 2) : 0.f;
            const uint8_t *q = blk + (kHasMin ? 4 : 2);
            uint32_t qh = 0;
            if constexpr (kFive) {
                qh = q[0] | (q[1] << 8) | (q[2] << 16) | (static_cast<uint32_t>(q[3]) << 24);
                q += 4;
            }
            constexpr int kBias = kHasMin ? 0 : (kFive ? 16 : 8);
            for (int j = 0; j < 16; ++j) {
                int x0 = q[j] & 0x0F, x1 = q[j] >> 4;
                if constexpr (kFive) {
                    x0 |= ((qh >> j) << 4) & 0x10;
                    x1 |= (qh >> (j + 12)) & 0x10;
                }
                dst[j] = __float2half_rn(static_cast<float>(x0 - kBias) * d + m);
                dst[j + 16] = __float2half_rn(static_cast<float>(x1 - kBias) * d + m);
            }
        }
    }
}

// fp16 [rows, K] -> fp16 [rows, ld_out] with optional row permutation and zero K padding
__global__ void copy_rows_f16_kernel(const __half *__restrict__ src, __half *__restrict__ dst, long long rows, int K, int ld_out,
                                     const int *__restrict__ row_map) {
    const long long total = rows * ld_out;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / ld_out;
        const int k = static_cast<int>(i % ld_out);
        const long long orow = row_map ? row_map[row] : row;
        dst[orow * ld_out + k] = k < K ? src[row * K + k] : __float2half_rn(0.f);
    }
}

__global__ void permute_f32_kernel(const float *__restrict__ src, float *__restrict__ dst, int n, const int *__restrict__ map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[map[i]] = src[i];
}

// ---------------------------------------------------------------------------------------------
// Positional-embedding resampling on device (reference interpolate_pos_embed, dinov2.cpp:159-225:
// per channel cv::resize(INTER_CUBIC) of the M x M grid to gh x gw; cls row copied).  OpenCV convention:
// src = (dst + 0.5) * (M / g) - 0.5, Keys cubic a = -0.75, replicated border, horizontal pass then vertical.
__device__ __forceinline__ void cubic_w(float x, float w[4]) {
    const float A = -0.75f;
    w[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
    w[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    w[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
    w[3] = 1.f - w[0] - w[1] - w[2];
}
__global__ void pos_embed_bicubic_kernel(const float *__restrict__ pos, float *__restrict__ out, int M, int gh, int gw, int D) {
    // grid.x = 1 + gh*gw output rows, threads over D
    const int row = blockIdx.x;
    if (row == 0) {
        for (int d = threadIdx.x; d < D; d += blockDim.x) out[d] = pos[d];
        return;
    }
    const int oy = (row - 1) / gw, ox = (row - 1) % gw;
    const float fy = static_cast<float>((oy + 0.5) * (static_cast<double>(M) / gh) - 0.5);
    const float fx = static_cast<float>((ox + 0.5) * (static_cast<double>(M) / gw) - 0.5);
    const int sy = static_cast<int>(floorf(fy)), sx = static_cast<int>(floorf(fx));
    float wy[4], wx[4];
    cubic_w(fy - sy, wy);
    cubic_w(fx - sx, wx);
    int iy[4], ix[4];
    for (int k = 0; k < 4; ++k) {
        iy[k] = min(max(sy - 1 + k, 0), M - 1);
        ix[k] = min(max(sx - 1 + k, 0), M - 1);
    }
    const float *grid = pos + D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int a = 0; a < 4; ++a) {
            const float *r = grid + static_cast<size_t>(iy[a]) * M * D + d;
            float h = r[static_cast<size_t>(ix[0]) * D] * wx[0];
            h += r[static_cast<size_t>(ix[1]) * D] * wx[1];
            h += r[static_cast<size_t>(ix[2]) * D] * wx[2];
            h += r[static_cast<size_t>(ix[3]) * D] * wx[3];
            acc = a == 0 ? h * wy[0] : acc + h * wy[a];
        }
        out[static_cast<size_t>(row) * D + d] = acc;
    }
}


// ---------------------------------------------------------------------------------------------
// Image preprocessing on device (reference dino_preprocess / dino_classify_preprocess, dinov2.cpp:106-156):
// u8 BGR HWC -> float (x * 1/255) -> cv::resize(INTER_CUBIC) to (RH, RW) -> optional centre crop (OH, OW at offset
// cy, cx) -> per-channel (v - mean) / std, output float BGR HWC (what dino_predict consumes).  Same OpenCV sampling
// convention and summation order (horizontal taps first, then vertical) as pos_embed_bicubic_kernel.
// One thread per output pixel (3 channels).
__global__ void preprocess_bicubic_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst, int B, int H, int W, int RH,
                                          int RW, int OH, int OW, int cy, int cx, float3 mean_bgr, float3 inv_std_bgr) {
    const long long total = static_cast<long long>(B) * OH * OW;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(idx % OW);
        const int oy = static_cast<int>((idx / OW) % OH);
        const int b = static_cast<int>(idx / (static_cast<long long>(OW) * OH));
        const int ry = oy + cy, rx = ox + cx;                       // coordinates in the resized image
        const float fy = static_cast<float>((ry + 0.5) * (static_cast<double>(H) / RH) - 0.5);
        const float fx = static_cast<float>((rx + 0.5) * (static_cast<double>(W) / RW) - 0.5);
        const int sy = static_cast<int>(floorf(fy)), sx = static_cast<int>(floorf(fx));
        float wy[4], wx[4];
        cubic_w(fy - sy, wy);
        cubic_w(fx - sx, wx);
        int iy[4], ix[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            iy[k] = min(max(sy - 1 + k, 0), H - 1);
            ix[k] = min(max(sx - 1 + k, 0), W - 1);
        }
        const uint8_t *img = src + static_cast<size_t>(b) * H * W * 3;
        float acc[3] = {0.f, 0.f, 0.f};
        const float s255 = static_cast<float>(1.0 / 255.0);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const uint8_t *row = img + static_cast<size_t>(iy[a]) * W * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float h = (static_cast<float>(row[ix[0] * 3 + c]) * s255) * wx[0];
                h += (static_cast<float>(row[ix[1] * 3 + c]) * s255) * wx[1];
                h += (static_cast<float>(row[ix[2] * 3 + c]) * s255) * wx[2];
                h += (static_cast<float>(row[ix[3] * 3 + c]) * s255) * wx[3];
                acc[c] = a == 0 ? h * wy[0] : acc[c] + h * wy[a];
            }
        }
        float *o = dst + idx * 3;
        o[0] = (acc[0] - mean_bgr.x) * inv_std_bgr.x;
        o[1] = (acc[1] - mean_bgr.y) * inv_std_bgr.y;
        o[2] = (acc[2] - mean_bgr.z) * inv_std_bgr.z;
    }
}

}  // namespace dino
