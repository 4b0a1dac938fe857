// PCA colouring of patch features on the device — the "callers" row of the hot path:
// inference.cpp:76-92 and realtime.cpp run cv::PCA(patch_tokens, DATA_AS_ROW, 3) + project + cv::normalize(0..255, MINMAX,
// CV_8U) on the host for every image; at > 800 images/s that eigen-decomposition of a D x D covariance is the bottleneck.
//
// Per image (NP x D features, row major):  mean over rows;  top-3 eigenvectors of Xc^T Xc by orthogonal (subspace)
// iteration  V <- orth(Xc^T (Xc V))  — never forms the D x D covariance, every pass streams the 5.6 MB feature block out of
// L2;  projection P = Xc V (NP x 3);  u8 = round((P - min) * 255 / (max - min)) over the whole NP x 3 block.
// The sign of a principal component is arbitrary (OpenCV's Jacobi solver does not fix it either): ours makes the
// largest-magnitude loading of each component positive.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dino {

constexpr int PCA_ITERS = 64;

// mean[b][d] = mean over rows of X[b]
__global__ void pca_mean_kernel(const float *__restrict__ X, float *__restrict__ mean, int NP, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (d >= D) return;
    const float *x = X + static_cast<size_t>(b) * NP * D + d;
    float s0 = 0.f, s1 = 0.f;
    int p = 0;
    for (; p + 1 < NP; p += 2) {
        s0 += x[static_cast<size_t>(p) * D];
        s1 += x[static_cast<size_t>(p + 1) * D];
    }
    if (p < NP) s0 += x[static_cast<size_t>(p) * D];
    mean[static_cast<size_t>(b) * D + d] = (s0 + s1) / static_cast<float>(NP);
}

// deterministic, well-spread start vectors (32-bit LCG), orthonormalised by the first pca_orth_kernel call
__global__ void pca_init_kernel(float *__restrict__ V, int D) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (d >= D) return;
    uint32_t s = 12345u + 747796405u * static_cast<uint32_t>(d);
    for (int c = 0; c < 3; ++c) {
        s = s * 1664525u + 1013904223u;
        V[(static_cast<size_t>(b) * D + d) * 3 + c] = static_cast<float>((s >> 8) & 0xFFFF) / 32768.0f - 1.0f;
    }
}

// Y[b][p][c] = sum_d (X[b][p][d] - mean[b][d]) * V[b][d][c]      one warp per row
__global__ void pca_project_kernel(const float *__restrict__ X, const float *__restrict__ mean, const float *__restrict__ V,
                                   float *__restrict__ Y, int NP, int D) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (warp >= NP) return;
    const float *x = X + (static_cast<size_t>(b) * NP + warp) * D;
    const float *m = mean + static_cast<size_t>(b) * D;
    const float *v = V + static_cast<size_t>(b) * D * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float xc = x[d] - m[d];
        a0 = fmaf(xc, v[d * 3 + 0], a0);
        a1 = fmaf(xc, v[d * 3 + 1], a1);
        a2 = fmaf(xc, v[d * 3 + 2], a2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        float *y = Y + (static_cast<size_t>(b) * NP + warp) * 3;
        y[0] = a0; y[1] = a1; y[2] = a2;
    }
}

// W[b][d][c] = sum_p (X[b][p][d] - mean[b][d]) * Y[b][p][c]      one thread per column, rows split over blockIdx.z
__global__ void pca_backproject_kernel(const float *__restrict__ X, const float *__restrict__ mean, const float *__restrict__ Y,
                                       float *__restrict__ W, int NP, int D, int rows_per_split) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.z * rows_per_split;
    const int p1 = min(NP, p0 + rows_per_split);
    if (d >= D) return;
    const float m = mean[static_cast<size_t>(b) * D + d];
    const float *x = X + static_cast<size_t>(b) * NP * D + d;
    const float *y = Y + static_cast<size_t>(b) * NP * 3;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int p = p0; p < p1; ++p) {
        const float xc = x[static_cast<size_t>(p) * D] - m;
        a0 = fmaf(xc, __ldg(y + p * 3 + 0), a0);
        a1 = fmaf(xc, __ldg(y + p * 3 + 1), a1);
        a2 = fmaf(xc, __ldg(y + p * 3 + 2), a2);
    }
    float *w = W + (static_cast<size_t>(b) * D + d) * 3;
    atomicAdd(w + 0, a0);
    atomicAdd(w + 1, a1);
    atomicAdd(w + 2, a2);
}

__device__ __forceinline__ float pca_block_sum(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

// V[b] = Gram-Schmidt(W[b]) (component 0 first, so the iteration converges to eigenvectors in descending order of eigenvalue);
// W is cleared for the next accumulation.  fix_sign: make the largest-magnitude loading of each component positive.
__global__ void pca_orth_kernel(float *__restrict__ W, float *__restrict__ V, int D, int fix_sign) {
    __shared__ float red[32];
    __shared__ float best_abs[3], best_val[3];
    const int b = blockIdx.x;
    float *w = W + static_cast<size_t>(b) * D * 3;
    float *v = V + static_cast<size_t>(b) * D * 3;
    for (int c = 0; c < 3; ++c) {
        for (int k = 0; k < c; ++k) {                       // w_c -= (w_c . v_k) v_k
            float dot = 0.f;
            for (int d = threadIdx.x; d < D; d += blockDim.x) dot = fmaf(w[d * 3 + c], v[d * 3 + k], dot);
            dot = pca_block_sum(dot, red);
            for (int d = threadIdx.x; d < D; d += blockDim.x) w[d * 3 + c] -= dot * v[d * 3 + k];
            __syncthreads();
        }
        float nn = 0.f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) nn = fmaf(w[d * 3 + c], w[d * 3 + c], nn);
        nn = pca_block_sum(nn, red);
        const float inv = nn > 0.f ? rsqrtf(nn) : 0.f;
        for (int d = threadIdx.x; d < D; d += blockDim.x) v[d * 3 + c] = w[d * 3 + c] * inv;
        __syncthreads();
    }
    if (fix_sign) {
        if (threadIdx.x < 3) { best_abs[threadIdx.x] = -1.f; best_val[threadIdx.x] = 1.f; }
        __syncthreads();
        if (threadIdx.x == 0) {                              // D <= 1536: a serial scan is fine, and deterministic
            for (int c = 0; c < 3; ++c)
                for (int d = 0; d < D; ++d) {
                    const float a = fabsf(v[d * 3 + c]);
                    if (a > best_abs[c]) { best_abs[c] = a; best_val[c] = v[d * 3 + c]; }
                }
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x)
            for (int c = 0; c < 3; ++c)
                if (best_val[c] < 0.f) v[d * 3 + c] = -v[d * 3 + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * 3; i += blockDim.x) w[i] = 0.f;
}

// rgb[b][p][c] = saturate(round((Y - min) * 255 / (max - min)))  with min / max over the whole NP x 3 block of image b
// (cv::normalize(src, dst, 0, 255, NORM_MINMAX, CV_8U), inference.cpp:84)
__global__ void pca_to_u8_kernel(const float *__restrict__ Y, uint8_t *__restrict__ rgb, int n /* NP * 3 */) {
    __shared__ float smin[32], smax[32];
    const int b = blockIdx.x;
    const float *y = Y + static_cast<size_t>(b) * n;
    float lo = INFINITY, hi = -INFINITY;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        lo = fminf(lo, y[i]);
        hi = fmaxf(hi, y[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    lo = INFINITY; hi = -INFINITY;
    for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
    const float scale = hi > lo ? 255.0f / (hi - lo) : 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = rintf((y[i] - lo) * scale);        // cvRound: round half to even
        rgb[static_cast<size_t>(b) * n + i] = static_cast<uint8_t>(fminf(fmaxf(v, 0.f), 255.f));
    }
}

}  // namespace dino
