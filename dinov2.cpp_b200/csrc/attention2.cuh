// Fused multi-head self-attention, second generation (head_dim 64, no mask).  Same contract as attention.cuh
// (replaces the reference's mul_mat(K,Q) -> soft_max_ext -> mul_mat(V,P) chain, dinov2.cpp:479-543) but laid out
// around the three throughput limits of one SM: tensor pipe (512 cycles per 128x128 score tile at hd 64),
// MUFU (16 exp/cycle) and instruction issue.
//
//  * One CTA owns 256 query rows = two 128-row tiles; each tile has its own softmax warpgroup (4 warps, one TMEM
//    lane quarter each), so every SM sub-partition always has a second warp to issue from while the first waits on
//    TMEM / mbarriers, and every K/V tile fetched by TMA is used twice.
//  * exp2 runs as ex2.approx.f16x2: two probabilities per MUFU op, produced directly in the packed fp16 form the
//    P*V tensor-core operand needs.  The fp32 score is scaled and shifted in fp32 ((s - m) * log2e/8), only the
//    (<= 0) exponent argument is rounded to fp16.
//  * The softmax denominator comes from the tensor core: V is extended by a block of ones (MMA N = 80), so column
//    64 of the P*V accumulator is sum_k P[k] of exactly the fp16-rounded probabilities used in the numerator.
//  * Registers are rebalanced with setmaxnreg: 224 for the softmax warpgroups (128 scores + 64 output values live),
//    48 for the TMA / MMA / allocator warpgroup (128*48 + 256*224 <= 384*168, the launch-time pool).
//
// Pipeline per K/V tile j and query tile t (mbarriers; phases flip once per tile):
//   MMA   : S_t(j+1) = Q_t K_{j+1}^T as soon as the softmax warpgroup has pulled S_t(j) into registers (s_free)
//   WG t  : m, alpha, P_t(j) = exp2(..) ; fold Opart_t(j-1) into the register accumulator ; P_t(j) -> smem (p_full)
//   MMA   : Opart_t(j) = P_t(j) [V_j | 1]  (fresh accumulator, o_full) ; K/V stage released after both tiles
#pragma once
#include "ptx.cuh"

namespace dino {

constexpr int AT2_THREADS = 384;
constexpr int AT2_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
constexpr int AT2_KV_STAGES = 3;
constexpr int AT2_SMEM_BYTES = 2 * AT2_TILE                      // Q0, Q1
                               + AT2_KV_STAGES * 2 * AT2_TILE    // K, V ring
                               + 2 * 2 * AT2_TILE                // P0, P1 (128 x 128 fp16 each)
                               + AT2_TILE                        // ones block
                               + 256 + 1024;

struct Attn2Params {
    int n_tok;
    int hidden;
    __half *out;
    float scale_log2;   // log2(e) / sqrt(64)
};

__global__ void __launch_bounds__(AT2_THREADS, 1)
attention_fwd_v2(const __grid_constant__ CUtensorMap tmQKV, const Attn2Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [2]
    uint8_t *sK = sQ + 2 * AT2_TILE;                      // [stages]
    uint8_t *sV = sK + AT2_KV_STAGES * AT2_TILE;          // [stages]
    uint8_t *sP = sV + AT2_KV_STAGES * AT2_TILE;          // [2] x 32 KB
    uint8_t *sOnes = sP + 4 * AT2_TILE;                   // 16 KB of 1.0h
    uint64_t *bars = reinterpret_cast<uint64_t *>(sOnes + AT2_TILE);
    uint64_t *q_full = bars;                              // 1
    uint64_t *kv_full = bars + 1;                         // stages
    uint64_t *kv_empty = kv_full + AT2_KV_STAGES;         // stages
    uint64_t *s_full = kv_empty + AT2_KV_STAGES;          // 2
    uint64_t *s_free = s_full + 2;                        // 2
    uint64_t *p_full = s_free + 2;                        // 2
    uint64_t *o_full = p_full + 2;                        // 2
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(o_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q_base = blockIdx.x * 256;
    const int head = blockIdx.y;
    const int img = blockIdx.z;
    const int row_base = img * p.n_tok;
    const int n_kv = (p.n_tok + 127) / 128;
    const bool has_q1 = q_base + 128 < p.n_tok;           // the last CTA of an image may own a single query tile

    if (warp == 0 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < AT2_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 128);
            mbar_init(&p_full[t], 128);
            mbar_init(&o_full[t], 1);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 512);
    {   // ones block (the extra B columns of the P*V MMA)
        uint4 *o = reinterpret_cast<uint4 *>(sOnes);
        const uint32_t one2 = 0x3C003C00u;
        for (int i = threadIdx.x; i < AT2_TILE / 16; i += AT2_THREADS) o[i] = make_uint4(one2, one2, one2, one2);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;           // S_t at columns 128 t
    const uint32_t tmem_O = tmem_base + 256;     // Opart_t at columns 256 + 128 t (80 used: 64 dims + 16 row-sum copies)

    if (warp < 4) {
        setmaxnreg_dec<48>();
        if (warp == 0 && lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            mbar_arrive_expect_tx(q_full, (has_q1 ? 2 : 1) * AT2_TILE);
            tma_load_2d(sQ, &tmQKV, q_full, head * 64, row_base + q_base);
            if (has_q1) tma_load_2d(sQ + AT2_TILE, &tmQKV, q_full, head * 64, row_base + q_base + 128);
            int s = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&kv_full[s], 2 * AT2_TILE);
                tma_load_2d(sK + s * AT2_TILE, &tmQKV, &kv_full[s], p.hidden + head * 64, row_base + j * 128);
                tma_load_2d(sV + s * AT2_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + head * 64, row_base + j * 128);
                if (++s == AT2_KV_STAGES) { s = 0; ph ^= 1; }
            }
        } else if (warp == 1 && lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 80, 0, 1);     // B = [V | ones], MN-major
            const int n_t = has_q1 ? 2 : 1;
            auto issue_s = [&](int t, int stage) {
                const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ + t * AT2_TILE), 16, 1024);
                const uint64_t k_desc = make_smem_desc_sw128(smem_u32(sK + stage * AT2_TILE), 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S + t * 128, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[t]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            for (int t = 0; t < n_t; ++t) issue_s(t, 0);
            int s = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_kv; ++j) {
                const uint32_t par = j & 1;
                if (j + 1 < n_kv) {
                    int s1 = s + 1;
                    uint32_t ph1 = ph;
                    if (s1 == AT2_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                    mbar_wait(&kv_full[s1], ph1);
                    for (int t = 0; t < n_t; ++t) {
                        mbar_wait(&s_free[t], par);       // softmax warpgroup t holds S_t(j) in registers
                        tc_fence_after();
                        issue_s(t, s1);
                    }
                }
                const uint32_t v_addr = smem_u32(sV + s * AT2_TILE);
                // MN-major B: atom 0 = the V tile (64 dims), atom 1 (leading-dim byte offset away) = the ones block
                const uint64_t v_desc = make_smem_desc_sw128(v_addr, smem_u32(sOnes) - v_addr, 1024);
                for (int t = 0; t < n_t; ++t) {
                    mbar_wait(&p_full[t], par);
                    tc_fence_after();
                    const uint64_t p_desc = make_smem_desc_sw128(smem_u32(sP + t * 2 * AT2_TILE), 16, 1024);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t a = p_desc + static_cast<uint64_t>((k >> 2) * (AT2_TILE >> 4) + (k & 3) * 2);
                        const uint64_t b = v_desc + static_cast<uint64_t>(k * (2048 >> 4));
                        umma_f16_ss(tmem_O + t * 128, a, b, idesc_o, k != 0);
                    }
                    umma_commit(&o_full[t]);
                }
                umma_commit(&kv_empty[s]);
                if (++s == AT2_KV_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        setmaxnreg_inc<224>();
        const int t = (warp - 4) >> 2;                    // query tile / warpgroup
        const int qd = warp & 3;                          // TMEM lane quarter
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        if (t == 0 || has_q1) {
            float o_acc[64];
#pragma unroll
            for (int d = 0; d < 64; ++d) o_acc[d] = 0.f;
            float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
            const float c = p.scale_log2;
            uint8_t *p_row = sP + t * 2 * AT2_TILE + (r >> 3) * 1024 + (r & 7) * 128;
            const uint32_t sw = static_cast<uint32_t>(r & 7);
            if (has_q1 && t == 1) named_bar_arrive(1, 256);     // warpgroup 0 takes the first exp turn

            auto fold_o = [&](int j, float alpha) {
                mbar_wait(&o_full[t], j & 1);
                tc_fence_after();
                uint32_t a[32], b[32], rs;
                tmem_ld_32x32b_x32(tmem_O + lane_addr + t * 128, a);
                tmem_ld_32x32b_x32(tmem_O + lane_addr + t * 128 + 32, b);
                tmem_ld_32x32b_x1(tmem_O + lane_addr + t * 128 + 64, rs);
                tmem_ld_wait();
#pragma unroll
                for (int d = 0; d < 32; ++d) {
                    o_acc[d] = fmaf(o_acc[d], alpha, __uint_as_float(a[d]));
                    o_acc[d + 32] = fmaf(o_acc[d + 32], alpha, __uint_as_float(b[d]));
                }
                l_run = fmaf(l_run, alpha, __uint_as_float(rs));
            };

            for (int j = 0; j < n_kv; ++j) {
                mbar_wait(&s_full[t], j & 1);
                tc_fence_after();
                uint32_t sv[4][32];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(tmem_S + lane_addr + t * 128 + cc * 32, sv[cc]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[t]);                  // S_t may be overwritten by the next Q K^T

                const int kv_valid = p.n_tok - j * 128;
                if (kv_valid < 128) {                     // last tile: keys past the image's tokens do not exist
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (cc * 32 + i >= kv_valid) sv[cc][i] = 0xFF800000u;   // -inf
                }
                float mx4[4] = {m_run, -INFINITY, -INFINITY, -INFINITY};      // four independent chains (ILP)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int i = 0; i < 32; i += 2)
                        mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(sv[cc][i]), __uint_as_float(sv[cc][i + 1]));
                const float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
                const float alpha = ex2_approx((m_run - mx) * c);     // first tile: ex2(-inf) = 0
                m_run = mx;
                const float mc = mx * c;
                // Ping-pong: only one warpgroup at a time runs its MUFU-bound exp phase; the other one meanwhile does its
                // TMEM loads, max, output fold and P stores (named barriers 1 + t, armed by the other warpgroup).
                if (has_q1) named_bar_sync(1 + t, 256);
                // P(j) in registers as packed fp16 pairs: exp2((s - m) * c)
                uint32_t pk[64];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float x0 = fmaf(__uint_as_float(sv[cc][i]), c, -mc);
                        const float x1 = fmaf(__uint_as_float(sv[cc][i + 1]), c, -mc);
                        pk[cc * 16 + (i >> 1)] = ex2_f16x2(cvt_f16x2(x0, x1));
                    }
                if (has_q1) named_bar_arrive(1 + (t ^ 1), 256);
                // the P buffer and the Opart accumulator of this tile are free once P(j-1) V(j-1) has completed
                if (j > 0) fold_o(j - 1, alpha_prev);
                alpha_prev = alpha;
#pragma unroll
                for (int g = 0; g < 16; ++g) {            // 16-byte chunk g of the 128-key row (8 keys)
                    uint8_t *dst = p_row + (g >> 3) * AT2_TILE + (((g & 7) ^ sw) << 4);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                }
                fence_proxy_async_smem();
                mbar_arrive(&p_full[t]);
            }
            fold_o(n_kv - 1, alpha_prev);

            const int tok = q_base + t * 128 + r;
            if (tok < p.n_tok) {
                const float inv = 1.0f / l_run;
                uint4 *dst = reinterpret_cast<uint4 *>(p.out + static_cast<size_t>(row_base + tok) * p.hidden + head * 64);
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    dst[v] = make_uint4(pack_half2(o_acc[8 * v] * inv, o_acc[8 * v + 1] * inv),
                                        pack_half2(o_acc[8 * v + 2] * inv, o_acc[8 * v + 3] * inv),
                                        pack_half2(o_acc[8 * v + 4] * inv, o_acc[8 * v + 5] * inv),
                                        pack_half2(o_acc[8 * v + 6] * inv, o_acc[8 * v + 7] * inv));
            }
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
