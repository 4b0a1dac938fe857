// Fused multi-head self-attention, fourth generation: P never leaves tensor memory.
// Contract as attention3.cuh (replaces the reference's mul_mat(K,Q) -> soft_max_ext -> mul_mat(V,P) chain,
// dinov2.cpp:479-543; head_dim 64, no mask).  What changed against v3, and why (cycle trace of v3, profiles/r01_attn_trace.md):
//
//  * v3 moved P through shared memory (16 STS.128 per row and tile, fence.proxy.async, then the tensor core read the
//    32 KB tile back as the A operand).  With head_dim 64 every tcgen05.mma already reads 128 B/clk of operands from
//    shared memory, so the extra P traffic made the shared-memory port, not the MUFU unit or the tensor pipe, the
//    limiter: the single MMA-issuing thread stalled ~800 cycles per batch of MMAs and both softmax warpgroups sat idle
//    at the same time.  Here the softmax warps write P (packed fp16) straight back into the TMEM columns its scores
//    came from (tcgen05.st) and P V takes its A operand from tensor memory (tcgen05.mma [d], [a], b-desc).
//  * The tensor-pipe program order per query tile t is  P_t(j) V(j)  ->  S_t(j+1) = Q_t K(j+1)^T.  tcgen05.mma executes
//    in issue order, so S_t(j+1) may overwrite the columns P_t(j) lives in without a barrier, and while the tensor
//    pipe works on tile t the other warpgroup is exponentiating tile 1-t: the two warpgroups fall into anti-phase on
//    their own (MUFU and tensor pipe both stay busy) without named-barrier ping-pong.
//  * The row maximum is no longer on the critical path: tile j is exponentiated against the running reference maximum
//    while the new maximum is reduced in the shadow of the MUFU stream; only when some row of the warp grew past the
//    lazy-rescale threshold (2^8) the warp rescales O_t in TMEM and redoes the tile (rare after the first tiles).
//
// Roles: warps 0-3 / 4-7 = softmax warpgroups of query tile 0 / 1 (one thread per query row, TMEM lane quarter =
// warp % 4), warp 8 TMEM allocator, warp 10 TMA producer (Q per item, K/V ring), warp 11 MMA issuer.
// TMEM columns: S_t / P_t at 128 t (P_t = columns [0, 64) of S_t), O_t at 256 + 128 t (64 dims + the ones-trick
// denominator in column 64).
#pragma once
#include "ptx.cuh"

// every AT4_POLY_MOD-th pair of probabilities is computed with a polynomial on the FMA pipe instead of MUFU (0 = none)
#ifndef AT4_POLY_MOD
#define AT4_POLY_MOD 0
#endif

namespace dino {

constexpr int AT4_THREADS = 384;
constexpr int AT4_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
#ifndef AT4_KV_STAGES
#define AT4_KV_STAGES 4
#endif
constexpr int AT4_SMEM_BYTES = 2 * AT4_TILE + AT4_KV_STAGES * 2 * AT4_TILE + AT4_TILE + 256 + 1024;
constexpr float AT4_RESCALE_LOG2 = 8.0f;        // lazy-rescale threshold in the exp2 domain

// exp2(x) without the MUFU unit: round-to-nearest split x = n + f (magic-number add), cubic minimax for 2^f on
// [-0.5, 0.5], exponent field patched by integer add.  Arguments below -30 (masked keys are -inf) clamp to 2^-30, which
// is zero once P is rounded to fp16.
__device__ __forceinline__ float exp2_poly3_v4(float x) {
    const float t = fmaxf(x, -30.0f);
    const float u = t + 12582912.0f;                 // 1.5 * 2^23: low mantissa bits now hold round(t)
    const float f = t - (u - 12582912.0f);
    float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(u) << 23));
}

// Rare path of the lazy running-max correction: scale this warp's 32 rows of O_t (64 numerator columns and the
// denominator column) in TMEM.  Out of line so that its temporaries do not add to the register pressure of the softmax loop.
__device__ __noinline__ void attn4_rescale_rows(uint32_t o_addr, float alpha) {
    uint32_t a[32], b[32], rs;
    tmem_ld_32x32b_x32(o_addr, a);
    tmem_ld_32x32b_x32(o_addr + 32, b);
    tmem_ld_32x32b_x1(o_addr + 64, rs);
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < 32; ++d) {
        a[d] = __float_as_uint(__uint_as_float(a[d]) * alpha);
        b[d] = __float_as_uint(__uint_as_float(b[d]) * alpha);
    }
    tmem_st_32x32b_x32(o_addr, a);
    tmem_st_32x32b_x32(o_addr + 32, b);
    tmem_st_32x32b_x1(o_addr + 64, __float_as_uint(__uint_as_float(rs) * alpha));
    tmem_st_wait();
}

// Optional cycle trace of CTA 0 (compile with -DAT4_TRACE): (event id, index, clock) per role, written to p.trace
// ([role][512][2] uint64).  Roles: 0 = MMA warp, 1 = softmax WG0 thread 0, 2 = softmax WG1 thread 0.
#ifdef AT4_TRACE
#define AT4_EV(ROLE, ID, IDX)                                                                  \
    do {                                                                                       \
        if (blockIdx.x == 0 && p.trace && tr_n < 512) {                                        \
            p.trace[((ROLE) * 512 + tr_n) * 2] = (static_cast<unsigned long long>(ID) << 32) | static_cast<unsigned>(IDX); \
            p.trace[((ROLE) * 512 + tr_n) * 2 + 1] = clock64();                                \
            ++tr_n;                                                                            \
        }                                                                                      \
    } while (0)
#else
#define AT4_EV(ROLE, ID, IDX) do {} while (0)
#endif

struct Attn4Params {
    int n_tok;
    int hidden;
    int n_heads;
    int n_qblk;        // ceil(n_tok / 256)
    int num_items;     // batch * n_heads * n_qblk
    __half *out;
    float scale_log2;  // log2(e) / sqrt(64)
    unsigned long long *trace;   // AT4_TRACE builds only
};

__global__ void __launch_bounds__(AT4_THREADS, 1)
attention_fwd_v4(const __grid_constant__ CUtensorMap tmQKV, const Attn4Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [2]
    uint8_t *sK = sQ + 2 * AT4_TILE;                      // [stages]
    uint8_t *sV = sK + AT4_KV_STAGES * AT4_TILE;          // [stages]
    uint8_t *sOnes = sV + AT4_KV_STAGES * AT4_TILE;       // 16 KB of 1.0h
    uint64_t *bars = reinterpret_cast<uint64_t *>(sOnes + AT4_TILE);
    uint64_t *q_full = bars;                              // 1
    uint64_t *q_empty = bars + 1;                         // 1
    uint64_t *kv_full = bars + 2;                         // stages
    uint64_t *kv_empty = kv_full + AT4_KV_STAGES;         // stages
    uint64_t *s_full = kv_empty + AT4_KV_STAGES;          // 2: S_t(j) is in TMEM (and P_t(j-1) V(j-1) has completed)
    uint64_t *p_full = s_full + 2;                        // 2: P_t(j) is in TMEM
    uint64_t *o_full = p_full + 2;                        // 2: the last P_t V of the item has completed
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(o_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = (p.n_tok + 127) / 128;
    // contiguous, balanced item range of this CTA: consecutive items share K/V (same image and head), so a CTA re-reads
    // them from L2, and every CTA gets the same mix of full and half (single query tile) blocks
    const int item_lo = static_cast<int>(static_cast<long long>(p.num_items) * blockIdx.x / gridDim.x);
    const int item_hi = static_cast<int>(static_cast<long long>(p.num_items) * (blockIdx.x + 1) / gridDim.x);

    if (warp == 10 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 11 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < AT4_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 128);
            mbar_init(&o_full[t], 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_ptr, 512);
    {
        uint4 *o = reinterpret_cast<uint4 *>(sOnes);
        const uint32_t one2 = 0x3C003C00u;
        for (int i = threadIdx.x; i < AT4_TILE / 16; i += AT4_THREADS) o[i] = make_uint4(one2, one2, one2, one2);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;           // S_t / P_t at columns 128 t
    const uint32_t tmem_O = tmem_base + 256;     // O_t at columns 256 + 128 t (64 dims, column 64 = row sum)

    // work item -> (image, head, query block); consecutive items share K/V (same image and head) for L2 reuse
#define AT4_DECODE(ITEM, ROW_BASE, HEAD, Q_BASE, HAS_Q1)             \
    do {                                                            \
        const int qb__ = (ITEM) % p.n_qblk;                         \
        const int ih__ = (ITEM) / p.n_qblk;                         \
        (HEAD) = ih__ % p.n_heads;                                  \
        (ROW_BASE) = (ih__ / p.n_heads) * p.n_tok;                  \
        (Q_BASE) = qb__ * 256;                                      \
        (HAS_Q1) = (Q_BASE) + 128 < p.n_tok;                        \
    } while (0)

    if (warp >= 8) {
        setmaxnreg_dec<80>();
        if (warp == 10) {
            // ---------------------------------------------------------------- TMA producer (warp-uniform; one lane issues)
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                AT4_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_empty, item_ph ^ 1);           // every Q K^T of the previous item has completed
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full, (has_q1 ? 2 : 1) * AT4_TILE);
                    tma_load_2d(sQ, &tmQKV, q_full, head * 64, row_base + q_base);
                    if (has_q1) tma_load_2d(sQ + AT4_TILE, &tmQKV, q_full, head * 64, row_base + q_base + 128);
                }
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[s], 2 * AT4_TILE);
                        tma_load_2d(sK + s * AT4_TILE, &tmQKV, &kv_full[s], p.hidden + head * 64, row_base + j * 128);
                        tma_load_2d(sV + s * AT4_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + head * 64, row_base + j * 128);
                    }
                    __syncwarp();
                    if (++s == AT4_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 11) {
            // ---------------------------------------------------------------- MMA issuer
            // All 32 lanes run the control flow, barrier waits and descriptor arithmetic (warp-uniform -> uniform
            // datapath); one elected lane issues tcgen05.mma / tcgen05.commit.
            constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 80, 0, 1);     // A = P (TMEM), B = [V | ones], MN-major
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            uint32_t np0 = 0, np1 = 0;     // P tiles consumed so far per query tile (phase of p_full)
            int tr_n = 0; (void) tr_n;
            const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t q_desc1 = make_smem_desc_sw128(smem_u32(sQ + AT4_TILE), 16, 1024);
// S_t = Q_t K(stage)^T  (4 k-steps of 16 dims), then s_full[t]
#define AT4_ISSUE_S(QDESC, T, KDESC)                                                                                   \
    do {                                                                                                               \
        _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                                  \
            umma_f16_ss(tmem_S + (T) * 128, (QDESC) + 2 * k, (KDESC) + 2 * k, idesc_s, k != 0);                        \
        umma_commit(&s_full[T]);                                                                                       \
    } while (0)
// O_t (+)= P_t [V | 1]  (8 k-steps of 16 keys; P_t = 8 TMEM columns per step), K/V stage release, then the next S_t
#define AT4_STEP(QDESC, T, CNT, IS_LAST_TILE)                                                                          \
    do {                                                                                                               \
        mbar_wait(&p_full[T], (CNT) & 1);                                                                              \
        tc_fence_after();                                                                                              \
        AT4_EV(0, 7, (CNT));                                                                                           \
        if (elect_one()) {                                                                                             \
            _Pragma("unroll") for (int k = 0; k < 8; ++k)                                                              \
                umma_f16_ts(tmem_O + (T) * 128, tmem_S + (T) * 128 + 8 * k, v_desc + static_cast<uint64_t>(k * (2048 >> 4)), \
                            idesc_o, (j | k) != 0);                                                                    \
            if (j == n_kv - 1) umma_commit(&o_full[T]);                                                                \
            if (IS_LAST_TILE) umma_commit(&kv_empty[s]);                                                               \
            if (j + 1 < n_kv) AT4_ISSUE_S(QDESC, T, k_desc1);                                                          \
            AT4_EV(0, 3 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
    } while (0)
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                AT4_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_full, item_ph);
                mbar_wait(&kv_full[s], ph);
                tc_fence_after();
                {
                    const uint64_t k_desc0 = make_smem_desc_sw128(smem_u32(sK + s * AT4_TILE), 16, 1024);
                    if (elect_one()) {
                        AT4_ISSUE_S(q_desc0, 0, k_desc0);
                        if (has_q1) AT4_ISSUE_S(q_desc1, 1, k_desc0);
                        if (n_kv == 1) umma_commit(q_empty);
                    }
                    __syncwarp();
                }
                for (int j = 0; j < n_kv; ++j) {
                    int s1 = s + 1;
                    uint32_t ph1 = ph;
                    if (s1 == AT4_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                    if (j + 1 < n_kv) {
                        mbar_wait(&kv_full[s1], ph1);
                        tc_fence_after();
                    }
                    AT4_EV(0, 5, j);
                    const uint64_t k_desc1 = make_smem_desc_sw128(smem_u32(sK + s1 * AT4_TILE), 16, 1024);
                    const uint32_t v_addr = smem_u32(sV + s * AT4_TILE);
                    // MN-major B: atom 0 = the V tile (64 dims), atom 1 (leading-dim byte offset away) = the ones block
                    const uint64_t v_desc = make_smem_desc_sw128(v_addr, smem_u32(sOnes) - v_addr, 1024);
                    AT4_STEP(q_desc0, 0, np0, !has_q1);
                    if (has_q1) AT4_STEP(q_desc1, 1, np1, true);
                    if (j + 2 == n_kv) {                       // the last Q K^T of this item has been issued
                        if (elect_one()) umma_commit(q_empty);
                        __syncwarp();
                    }
                    s = s1;
                    ph = ph1;
                }
            }
        }
    } else {
        setmaxnreg_inc<208>();
        const int t = warp >> 2;                          // query tile / warpgroup
        const int qd = warp & 3;                          // TMEM lane quarter
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t s_addr = tmem_S + lane_addr + t * 128;      // S_t row; P_t is written back to its first 64 columns
        const uint32_t o_addr = tmem_O + lane_addr + t * 128;
        const float c = p.scale_log2;
        const float thr = AT4_RESCALE_LOG2 / c;           // threshold in raw-score units
        uint32_t n_tile = 0;                              // tiles processed by this warpgroup (phase of s_full)
        uint32_t n_item = 0;                              // items finished by this warpgroup (phase of o_full)
        int tr_n = 0; (void) tr_n;
#define AT4_SEV(ID) do { if ((threadIdx.x & 127) == 0) AT4_EV(1 + t, ID, n_tile); } while (0)

        for (int item = item_lo; item < item_hi; ++item) {
            int row_base, head, q_base;
            bool has_q1;
            AT4_DECODE(item, row_base, head, q_base, has_q1);
            if (t == 1 && !has_q1) continue;
            float m_used = -INFINITY;

            for (int j = 0; j < n_kv; ++j, ++n_tile) {
                AT4_SEV(10);
                mbar_wait(&s_full[t], n_tile & 1);
                tc_fence_after();
                AT4_SEV(11);
                uint32_t sv[4][32];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(s_addr + cc * 32, sv[cc]);
                tmem_ld_wait();
                AT4_SEV(12);

                const int kv_valid = p.n_tok - j * 128;
                if (kv_valid < 128) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (cc * 32 + i >= kv_valid) sv[cc][i] = 0xFF800000u;   // -inf
                }
                if (j == 0) {
                    // first tile of the item: the reference maximum is this tile's own row maximum
                    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(sv[cc][i]), __uint_as_float(sv[cc][i + 1]));
                    m_used = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
                }
                AT4_SEV(14);
                // P(j) = exp2((s - m_used) * c) as packed fp16 pairs, written back over S_t in TMEM 32 keys at a time; the
                // tile's own row maximum is reduced alongside (FMNMX3 in the shadow of the MUFU stream).  If a row of this
                // warp outgrew the reference maximum by more than the threshold, O_t is rescaled and the tile redone.
                for (;;) {
                    const float mc = m_used * c;
                    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        uint32_t pk[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float s0 = __uint_as_float(sv[h][2 * e]);
                            const float s1 = __uint_as_float(sv[h][2 * e + 1]);
                            mx4[e & 3] = fmax3(mx4[e & 3], s0, s1);
                            const float x0 = fmaf(s0, c, -mc);
                            const float x1 = fmaf(s1, c, -mc);
                            if (AT4_POLY_MOD > 0 && ((h * 16 + e) % (AT4_POLY_MOD > 0 ? AT4_POLY_MOD : 1)) == AT4_POLY_MOD - 1) {
                                // exponentiated on the FMA/ALU pipes (Cody-Waite split + cubic, rel. error 7.7e-5, below
                                // the fp16 rounding of P) so that the MUFU unit is not the only exp2 engine
                                pk[e] = cvt_f16x2(exp2_poly3_v4(x0), exp2_poly3_v4(x1));
                            } else {
                                // ex2.approx.f16x2 (two MUFU.EX2.F16 in SASS); the argument is rounded to fp16 first
                                pk[e] = ex2_f16x2(cvt_f16x2(x0, x1));
                            }
                        }
                        tmem_st_32x32b_x16(s_addr + h * 16, pk);
                    }
                    const float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
                    const bool grow = mx > m_used + thr;
                    if (!__any_sync(0xffffffffu, grow)) break;
                    // rare: s_full(j) implies P(j-1) V(j-1) has completed, so O_t is quiescent; scale this warp's 32 rows
                    const float alpha = grow ? ex2_approx((m_used - mx) * c) : 1.0f;
                    if (grow) m_used = mx;
                    attn4_rescale_rows(o_addr, alpha);
                }
                AT4_SEV(15);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[t]);
                AT4_SEV(17);
            }

            // ---- item epilogue: O_t / rowsum -> fp16 rows
            mbar_wait(&o_full[t], n_item & 1);
            ++n_item;
            tc_fence_after();
            uint32_t a[32], b[32], rs;
            tmem_ld_32x32b_x32(o_addr, a);
            tmem_ld_32x32b_x32(o_addr + 32, b);
            tmem_ld_32x32b_x1(o_addr + 64, rs);
            tmem_ld_wait();
            tc_fence_before();
            const int tok = q_base + t * 128 + r;
            if (tok < p.n_tok) {
                const float inv = 1.0f / __uint_as_float(rs);
                uint4 *dst = reinterpret_cast<uint4 *>(p.out + static_cast<size_t>(row_base + tok) * p.hidden + head * 64);
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    dst[v] = make_uint4(pack_half2(__uint_as_float(a[8 * v]) * inv, __uint_as_float(a[8 * v + 1]) * inv),
                                        pack_half2(__uint_as_float(a[8 * v + 2]) * inv, __uint_as_float(a[8 * v + 3]) * inv),
                                        pack_half2(__uint_as_float(a[8 * v + 4]) * inv, __uint_as_float(a[8 * v + 5]) * inv),
                                        pack_half2(__uint_as_float(a[8 * v + 6]) * inv, __uint_as_float(a[8 * v + 7]) * inv));
                    dst[v + 4] = make_uint4(pack_half2(__uint_as_float(b[8 * v]) * inv, __uint_as_float(b[8 * v + 1]) * inv),
                                            pack_half2(__uint_as_float(b[8 * v + 2]) * inv, __uint_as_float(b[8 * v + 3]) * inv),
                                            pack_half2(__uint_as_float(b[8 * v + 4]) * inv, __uint_as_float(b[8 * v + 5]) * inv),
                                            pack_half2(__uint_as_float(b[8 * v + 6]) * inv, __uint_as_float(b[8 * v + 7]) * inv));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

#undef AT4_DECODE
#undef AT4_SEV
#undef AT4_ISSUE_S
#undef AT4_STEP

}  // namespace dino
