// Fused multi-head self-attention for the DINOv2 encoder block (head_dim 64, no mask, non-causal):
//     per (image, head):  O = softmax(Q K^T / 8) V
// Replaces the reference's materialised path — ggml_mul_mat(K,Q) -> ggml_soft_max_ext -> ggml_mul_mat(V,P)
// plus five cont/permute copies (reference dinov2.cpp:479-543, softmax ops.cpp:4641-4737) — with one
// flash-style kernel: S and the per-tile P*V products live in TMEM, P goes through shared memory as the
// fp16 A operand of the second MMA, the running max / sum / output row stay in registers (fp32).
// Operands are fp16 (Q, K, V as written by the QKV GEMM epilogue; P rounded to fp16), accumulation fp32;
// SURVEY.md appendix D measured this against the reference's f32 x f32 products at the oracle's own
// noise floor.
//
// One CTA = 128 query rows of one (image, head).  256 threads:
//   warp 0   TMA: Q tile once, then K/V tiles (128 keys x 64) through a 2-stage ring
//   warp 1   MMA issuer:  S[j&1] = Q K_j^T (128x128x64),   Opart[j&1] = P_j V_j (128x64x128)
//   warp 2   TMEM allocator
//   warps 4-7  softmax: thread r owns query row r (TMEM lane r)
// QKV layout: [tokens, 3*D] fp16, token t of image b at row b*N + t, q|k|v thirds, head h at columns 64h..64h+63
// (the reference's split by thirds, dinov2.cpp:479-494).  Output: [tokens, D] fp16, head h at columns 64h..
#pragma once
#include "ptx.cuh"

namespace dino {

constexpr int ATT_BQ = 128;
constexpr int ATT_BKV = 128;
constexpr int ATT_HD = 64;
constexpr int ATT_THREADS = 256;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;                 // 16 KB: one 128 x 64 fp16 tile
constexpr int ATT_SMEM_BYTES = ATT_TILE_BYTES                 // Q
                               + 2 * 2 * ATT_TILE_BYTES       // K,V x 2 stages
                               + 2 * 2 * ATT_TILE_BYTES       // P x 2 buffers (128 x 128 fp16 = 2 tiles)
                               + 256 + 1024;

struct AttnParams {
    int n_tok;        // tokens per image (N)
    int hidden;       // D
    __half *out;      // [B*N, D]
    float scale_log2; // (1/sqrt(64)) * log2(e)
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_tcgen05(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;
    uint8_t *sK = sQ + ATT_TILE_BYTES;             // [2] stages
    uint8_t *sV = sK + 2 * ATT_TILE_BYTES;         // [2] stages
    uint8_t *sP = sV + 2 * ATT_TILE_BYTES;         // [2] buffers x 32 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(sP + 4 * ATT_TILE_BYTES);
    uint64_t *q_full = bars;            // 1
    uint64_t *kv_full = bars + 1;       // 2
    uint64_t *kv_empty = bars + 3;      // 2
    uint64_t *s_full = bars + 5;        // 2
    uint64_t *p_full = bars + 7;        // 2
    uint64_t *o_full = bars + 9;        // 2
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 11);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BQ;
    const int head = blockIdx.y;
    const int img = blockIdx.z;
    const int row_base = img * p.n_tok;
    const int n_kv = (p.n_tok + ATT_BKV - 1) / ATT_BKV;

    if (warp == 0 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 128);
            mbar_init(&o_full[s], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;          // 2 x 128 columns
    const uint32_t tmem_O = tmem_base + 256;    // 2 x 64 columns

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_2d(sQ, &tmQKV, q_full, head * ATT_HD, row_base + q0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], 2 * ATT_TILE_BYTES);
                tma_load_2d(sK + s * ATT_TILE_BYTES, &tmQKV, &kv_full[s], p.hidden + head * ATT_HD, row_base + j * ATT_BKV);
                tma_load_2d(sV + s * ATT_TILE_BYTES, &tmQKV, &kv_full[s], 2 * p.hidden + head * ATT_HD, row_base + j * ATT_BKV);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);   // B = V is MN-major (dims contiguous)
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&kv_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK + s * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_HD / 16; ++k)
                    umma_f16_ss(tmem_S + s * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[s]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                if (j + 1 < n_kv) issue_s(j + 1);
                mbar_wait(&p_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t p_desc = make_smem_desc_sw128(smem_u32(sP + s * 2 * ATT_TILE_BYTES), 16, 1024);
                const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_BKV / 16; ++k) {
                    // P: K-major, two 64-key halves 16 KB apart, +32 B per 16 keys inside a half
                    // V: MN-major, 16 keys = 16 rows of 128 B = +2048 B
                    const uint64_t a = p_desc + static_cast<uint64_t>((k >> 2) * (ATT_TILE_BYTES >> 4) + (k & 3) * 2);
                    const uint64_t b = v_desc + static_cast<uint64_t>(k * (2048 >> 4));
                    umma_f16_ss(tmem_O + s * 64, a, b, idesc_o, k != 0);
                }
                umma_commit(&o_full[s]);
                umma_commit(&kv_empty[s]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        const int r = q * 32 + lane;                     // query row inside the tile == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        float o_acc[ATT_HD];
#pragma unroll
        for (int d = 0; d < ATT_HD; ++d) o_acc[d] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, alpha_pending = 0.f;

        auto accumulate_o = [&](int j, float alpha) {
            const int s = j & 1;
            mbar_wait(&o_full[s], (j >> 1) & 1);
            tc_fence_after();
            uint32_t a[32], b[32];
            tmem_ld_32x32b_x32(tmem_O + lane_addr + s * 64, a);
            tmem_ld_32x32b_x32(tmem_O + lane_addr + s * 64 + 32, b);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) {
                o_acc[d] = fmaf(o_acc[d], alpha, __uint_as_float(a[d]));
                o_acc[d + 32] = fmaf(o_acc[d + 32], alpha, __uint_as_float(b[d]));
            }
        };

        for (int j = 0; j < n_kv; ++j) {
            const int s = j & 1;
            mbar_wait(&s_full[s], (j >> 1) & 1);
            tc_fence_after();
            uint32_t sv[4][32];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tmem_S + lane_addr + s * 128 + c * 32, sv[c]);
            tmem_ld_wait();

            const int kv_valid = p.n_tok - j * ATT_BKV;       // keys of this tile that exist (>= 128: all)
            float mx = m_run;
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float t = __uint_as_float(sv[c][i]) * p.scale_log2;
                    if (c * 32 + i >= kv_valid) t = -INFINITY;
                    sv[c][i] = __float_as_uint(t);
                    mx = fmaxf(mx, t);
                }
            const float alpha = ex2_approx(m_run - mx);        // first tile: ex2(-inf) = 0
            m_run = mx;
            float sum = 0.f;
            // P row -> smem, canonical K-major 128B-swizzle layout: 16-B chunk c of row r lives at chunk c ^ (r & 7)
            uint8_t *p_row = sP + s * 2 * ATT_TILE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float e[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        e[i] = ex2_approx(__uint_as_float(sv[c][g * 8 + i]) - mx);
                        sum += e[i];
                    }
                    const int chunk = (c & 1) * 4 + g;                      // 16-B chunk inside the 64-key half
                    uint8_t *dst = p_row + (c >> 1) * ATT_TILE_BYTES + ((chunk ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]),
                                                                 pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
                }
            }
            l_run = fmaf(l_run, alpha, sum);
            fence_proxy_async_smem();     // generic-proxy P stores -> visible to the tensor core (async proxy)
            tc_fence_before();            // S[s] reads are complete before the issuer may overwrite it
            mbar_arrive(&p_full[s]);

            if (j > 0) accumulate_o(j - 1, alpha_pending);
            alpha_pending = alpha;
        }
        accumulate_o(n_kv - 1, alpha_pending);

        const int tok = q0 + r;
        if (tok < p.n_tok) {
            const float inv = 1.0f / l_run;
            uint4 *dst = reinterpret_cast<uint4 *>(p.out + static_cast<size_t>(row_base + tok) * p.hidden + head * ATT_HD);
#pragma unroll
            for (int v = 0; v < 8; ++v)
                dst[v] = make_uint4(pack_half2(o_acc[8 * v] * inv, o_acc[8 * v + 1] * inv),
                                    pack_half2(o_acc[8 * v + 2] * inv, o_acc[8 * v + 3] * inv),
                                    pack_half2(o_acc[8 * v + 4] * inv, o_acc[8 * v + 5] * inv),
                                    pack_half2(o_acc[8 * v + 6] * inv, o_acc[8 * v + 7] * inv));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace dino
