// Fused multi-head self-attention, third generation: persistent, output accumulator resident in TMEM.
// Kept as the previous round's baseline for A/B runs (DINO_B200_ATTN=3).  Contract (every generation): replaces the
// reference's mul_mat(K,Q) -> soft_max_ext -> mul_mat(V,P)
// chain, dinov2.cpp:479-543; head_dim 64, no mask).  What changed against v2, and why:
//
//  * Persistent CTAs (grid = #SMs) walk a static list of work items (image, head, 256-query block).  TMEM allocation,
//    barrier set-up and the ones block are paid once per SM instead of once per item, the K/V ring keeps streaming
//    across item boundaries, and the output write-back of one item overlaps the first Q K^T of the next.
//  * O stays in TMEM for the whole item: P_j [V_j | 1] accumulates in place (tcgen05 accumulate flag), so the softmax
//    warps no longer pull a partial product out of TMEM and FMA it into 64 live registers per K/V tile.  The
//    running-max correction is applied lazily (only when a row maximum grows by more than 2^8 in the exponent
//    domain — probabilities then stay <= 256, exact in fp16/fp32), by the owning warp: tcgen05.ld -> scale ->
//    tcgen05.st on the rare tile that needs it.  Column 64 of the accumulator is the softmax denominator (ones trick),
//    so it is rescaled together with the numerator.
//
// Per K/V tile j and query tile t:
// Warps 0-3 / 4-7 are the softmax warpgroups of query tile 0 / 1 (TMEM lane quarter = warp % 4), warps 8-11 the control
// warpgroup.
//   MMA  : S_t(j+1) = Q_t K_{j+1}^T once WG t has S_t(j) in registers (s_free)
//   WG t : row max -> (rare) rescale of O_t -> P_t(j) = exp2((s - m) * log2e/8) -> smem   (p_full)
//   MMA  : O_t (+)= P_t(j) [V_j | 1]   (o_full; K/V stage released when both tiles are done)
// after the last tile: WG t reads O_t, divides by column 64, writes fp16 rows.
#pragma once
#include "ptx.cuh"

#ifndef AT3_EXP_F32
#define AT3_EXP_F32 0
#endif
// every AT3_POLY_MOD-th pair of probabilities is computed with a polynomial on the FMA pipe instead of MUFU (0 = none)
#ifndef AT3_POLY_MOD
#define AT3_POLY_MOD 0   // measured: 0 -> 933 us, 3 -> 950 us, 2 -> 1052 us per ViT-L layer (B=64): MUFU is not the limiter yet
#endif

namespace dino {

constexpr int AT3_THREADS = 384;
constexpr int AT3_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
constexpr int AT3_KV_STAGES = 3;
constexpr int AT3_SMEM_BYTES = 2 * AT3_TILE + AT3_KV_STAGES * 2 * AT3_TILE + 2 * 2 * AT3_TILE + AT3_TILE + 256 + 1024;
constexpr float AT3_RESCALE_LOG2 = 8.0f;        // lazy-rescale threshold in the exp2 domain


// exp2(x) for x <= 0 without the MUFU unit: round-to-nearest split x = n + f (magic-number add), cubic minimax for 2^f on
// [-0.5, 0.5], exponent field patched by integer add.  Arguments below -30 (masked keys are -inf) clamp to 2^-30, which
// is zero once P is rounded to fp16.
__device__ __forceinline__ float exp2_poly3(float x) {
    const float t = fmaxf(x, -30.0f);
    const float u = t + 12582912.0f;                 // 1.5 * 2^23: low mantissa bits now hold round(t)
    const float f = t - (u - 12582912.0f);
    float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(u) << 23));
}

// Rare path of the lazy running-max correction: scale this warp's 32 rows of O_t (64 numerator columns and the
// denominator column) in TMEM.  Kept out of line so that its 64 temporaries do not add to the register pressure of
// the softmax loop (the caller's live scores are only spilled on the tile that actually takes this path).
__device__ __noinline__ void attn3_rescale_rows(uint32_t o_addr, float alpha) {
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
    tc_fence_before();
}

// Optional cycle trace of CTA 0 (compile with -DAT3_TRACE): (event id, index, clock) triples per role, written to
// p.trace ([role][512][2] uint64).  Roles: 0 = MMA thread, 1 = softmax WG0 thread 0, 2 = softmax WG1 thread 0.
#ifdef AT3_TRACE
#define AT3_EV(ROLE, ID, IDX)                                                                  \
    do {                                                                                       \
        if (blockIdx.x == 0 && p.trace && tr_n < 512) {                                        \
            p.trace[((ROLE) * 512 + tr_n) * 2] = (static_cast<unsigned long long>(ID) << 32) | static_cast<unsigned>(IDX); \
            p.trace[((ROLE) * 512 + tr_n) * 2 + 1] = clock64();                                \
            ++tr_n;                                                                            \
        }                                                                                      \
    } while (0)
#else
#define AT3_EV(ROLE, ID, IDX) do {} while (0)
#endif

struct Attn3Params {
    int n_tok;
    int hidden;
    int n_heads;
    int n_qblk;        // ceil(n_tok / 256)
    int num_items;     // batch * n_heads * n_qblk
    __half *out;
    float scale_log2;  // log2(e) / sqrt(64)
    int pingpong;      // 1: the two softmax warpgroups alternate in the MUFU-bound phase (named barriers)
    unsigned long long *trace;   // AT3_TRACE builds only
};

__global__ void __launch_bounds__(AT3_THREADS, 1)
attention_fwd_v3(const __grid_constant__ CUtensorMap tmQKV, const Attn3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [2]
    uint8_t *sK = sQ + 2 * AT3_TILE;                      // [stages]
    uint8_t *sV = sK + AT3_KV_STAGES * AT3_TILE;          // [stages]
    uint8_t *sP = sV + AT3_KV_STAGES * AT3_TILE;          // [2] x 32 KB
    uint8_t *sOnes = sP + 4 * AT3_TILE;                   // 16 KB of 1.0h
    uint64_t *bars = reinterpret_cast<uint64_t *>(sOnes + AT3_TILE);
    uint64_t *q_full = bars;                              // 1
    uint64_t *q_empty = bars + 1;                         // 1
    uint64_t *kv_full = bars + 2;                         // stages
    uint64_t *kv_empty = kv_full + AT3_KV_STAGES;         // stages
    uint64_t *s_full = kv_empty + AT3_KV_STAGES;          // 2
    uint64_t *s_free = s_full + 2;                        // 2
    uint64_t *p_full = s_free + 2;                        // 2
    uint64_t *o_full = p_full + 2;                        // 2
    uint64_t *turn = o_full + 2;                          // 2: exp-phase hand-over between the two softmax warpgroups
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(turn + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = (p.n_tok + 127) / 128;
    // contiguous, balanced item range of this CTA: consecutive items share K/V (same image and head), so a CTA re-reads
    // them from L2, and every CTA gets the same mix of full and half (single query tile) blocks
    const int item_lo = static_cast<int>(static_cast<long long>(p.num_items) * blockIdx.x / gridDim.x);
    const int item_hi = static_cast<int>(static_cast<long long>(p.num_items) * (blockIdx.x + 1) / gridDim.x);

    // control warpgroup = warps 8-11 (TMEM allocator 8, TMA 10, MMA 11): the sub-partition arbiter prefers the highest
    // eligible warp id, and the MMA issuer must never queue behind the MUFU/FMA streams of the softmax warps
    if (warp == 10 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 11 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < AT3_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 128);
            mbar_init(&p_full[t], 128);
            mbar_init(&o_full[t], 1);
            mbar_init(&turn[t], 128);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_ptr, 512);
    {
        uint4 *o = reinterpret_cast<uint4 *>(sOnes);
        const uint32_t one2 = 0x3C003C00u;
        for (int i = threadIdx.x; i < AT3_TILE / 16; i += AT3_THREADS) o[i] = make_uint4(one2, one2, one2, one2);
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;           // S_t at columns 128 t
    const uint32_t tmem_O = tmem_base + 256;     // O_t at columns 256 + 128 t (64 dims, column 64 = row sum)

    // work item -> (image, head, query block); consecutive items share K/V (same image and head) for L2 reuse
#define decode(ITEM, ROW_BASE, HEAD, Q_BASE, HAS_Q1)                 \
    do {                                                            \
        const int qb__ = (ITEM) % p.n_qblk;                         \
        const int ih__ = (ITEM) / p.n_qblk;                         \
        (HEAD) = ih__ % p.n_heads;                                  \
        (ROW_BASE) = (ih__ / p.n_heads) * p.n_tok;                  \
        (Q_BASE) = qb__ * 256;                                      \
        (HAS_Q1) = (Q_BASE) + 128 < p.n_tok;                        \
    } while (0)

    if (warp >= 8) {
        setmaxnreg_dec<72>();
        if (warp == 10) {
            // ---------------------------------------------------------------- TMA producer (warp-uniform; one lane issues)
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                decode(item, row_base, head, q_base, has_q1);
                mbar_wait(q_empty, item_ph ^ 1);           // every Q K^T of the previous item has completed
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full, (has_q1 ? 2 : 1) * AT3_TILE);
                    tma_load_2d(sQ, &tmQKV, q_full, head * 64, row_base + q_base);
                    if (has_q1) tma_load_2d(sQ + AT3_TILE, &tmQKV, q_full, head * 64, row_base + q_base + 128);
                }
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[s], 2 * AT3_TILE);
                        tma_load_2d(sK + s * AT3_TILE, &tmQKV, &kv_full[s], p.hidden + head * 64, row_base + j * 128);
                        tma_load_2d(sV + s * AT3_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + head * 64, row_base + j * 128);
                    }
                    __syncwarp();
                    if (++s == AT3_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 11) {
            // ---------------------------------------------------------------- MMA issuer
            // All 32 lanes run the control flow, barrier waits and descriptor arithmetic (warp-uniform -> uniform
            // datapath); one elected lane issues tcgen05.mma / tcgen05.commit.
            constexpr uint32_t idesc_s = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 80, 0, 1);     // B = [V | ones], MN-major
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            // tiles issued so far per query tile (phase bookkeeping of s_free / p_full); scalars, not arrays: this thread
            // is on the critical path of both warpgroups and must not touch local memory
            uint32_t ns0 = 0, ns1 = 0, np0 = 0, np1 = 0;
            int tr_n = 0; (void) tr_n;
            const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t q_desc1 = make_smem_desc_sw128(smem_u32(sQ + AT3_TILE), 16, 1024);
            const uint64_t p_desc0 = make_smem_desc_sw128(smem_u32(sP), 16, 1024);
            const uint64_t p_desc1 = make_smem_desc_sw128(smem_u32(sP + 2 * AT3_TILE), 16, 1024);
// (macros, not lambdas: everything must stay in registers of this single issuing thread)
#define AT3_ISSUE_S(QDESC, DTMEM, T, CNT, STAGE)                                                                       \
    do {                                                                                                               \
        if ((CNT) > 0) {                                                                                               \
            mbar_wait(&s_free[T], ((CNT) - 1) & 1);                                                                    \
            tc_fence_after();                                                                                          \
        }                                                                                                              \
        AT3_EV(0, 6, (CNT));                                                                                           \
        const uint64_t k_desc__ = make_smem_desc_sw128(smem_u32(sK + (STAGE) * AT3_TILE), 16, 1024);                   \
        if (elect_one()) {                                                                                             \
            _Pragma("unroll") for (int k = 0; k < 4; ++k) umma_f16_ss((DTMEM), (QDESC) + 2 * k, k_desc__ + 2 * k, idesc_s, k != 0); \
            umma_commit(&s_full[T]);                                                                                   \
            AT3_EV(0, 1 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
    } while (0)
#define AT3_ISSUE_PV(PDESC, VDESC, DTMEM, T, CNT, J, LAST, KVS)                                                                   \
    do {                                                                                                               \
        mbar_wait(&p_full[T], (CNT) & 1);                                                                              \
        tc_fence_after();                                                                                              \
        AT3_EV(0, 7, (CNT));                                                                                           \
        if (elect_one()) {                                                                                             \
            _Pragma("unroll") for (int k = 0; k < 8; ++k) {                                                            \
                const uint64_t a__ = (PDESC) + static_cast<uint64_t>((k >> 2) * (AT3_TILE >> 4) + (k & 3) * 2);        \
                const uint64_t b__ = (VDESC) + static_cast<uint64_t>(k * (2048 >> 4));                                 \
                umma_f16_ss((DTMEM), a__, b__, idesc_o, ((J) | k) != 0);                                               \
            }                                                                                                          \
            umma_commit(&o_full[T]);                                                                                   \
            if (LAST) umma_commit(&kv_empty[KVS]);                                                                     \
            AT3_EV(0, 3 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
    } while (0)
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                decode(item, row_base, head, q_base, has_q1);
                mbar_wait(q_full, item_ph);
                mbar_wait(&kv_full[s], ph);
                tc_fence_after();
                AT3_ISSUE_S(q_desc0, tmem_S, 0, ns0, s);
                if (has_q1) AT3_ISSUE_S(q_desc1, tmem_S + 128, 1, ns1, s);
                if (n_kv == 1 && elect_one()) umma_commit(q_empty);
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) {
                        int s1 = s + 1;
                        uint32_t ph1 = ph;
                        if (s1 == AT3_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                        AT3_EV(0, 8, j);
                        mbar_wait(&kv_full[s1], ph1);
                        tc_fence_after();
                        AT3_EV(0, 5, j);
                        AT3_ISSUE_S(q_desc0, tmem_S, 0, ns0, s1);
                        if (has_q1) AT3_ISSUE_S(q_desc1, tmem_S + 128, 1, ns1, s1);
                        if (j + 2 == n_kv && elect_one()) umma_commit(q_empty);   // last Q K^T of this item is in flight
                        __syncwarp();
                    }
                    const uint32_t v_addr = smem_u32(sV + s * AT3_TILE);
                    // MN-major B: atom 0 = the V tile (64 dims), atom 1 (leading-dim byte offset away) = the ones block
                    const uint64_t v_desc = make_smem_desc_sw128(v_addr, smem_u32(sOnes) - v_addr, 1024);
                    // the K/V stage is released by the last P V product that reads it
                    AT3_ISSUE_PV(p_desc0, v_desc, tmem_O, 0, np0, j, !has_q1, s);
                    if (has_q1) AT3_ISSUE_PV(p_desc1, v_desc, tmem_O + 128, 1, np1, j, true, s);
                    if (++s == AT3_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        setmaxnreg_inc<216>();
        const int t = warp >> 2;                          // query tile / warpgroup
        const int qd = warp & 3;                          // TMEM lane quarter
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t o_addr = tmem_O + lane_addr + t * 128;
        const float c = p.scale_log2;
        const float thr = AT3_RESCALE_LOG2 / c;           // threshold in raw-score units
        uint8_t *p_row = sP + t * 2 * AT3_TILE + (r >> 3) * 1024 + (r & 7) * 128;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        uint32_t n_tile = 0;                              // tiles processed by this warpgroup (phase bookkeeping)
        uint32_t n_pp = 0;                                // ping-pong turns taken so far (phase of turn[])
        int tr_n = 0; (void) tr_n;
#define AT3_SEV(ID) do { if ((threadIdx.x & 127) == 0) AT3_EV(1 + t, ID, n_tile); } while (0)

        for (int item = item_lo; item < item_hi; ++item) {
            int row_base, head, q_base;
            bool has_q1;
            decode(item, row_base, head, q_base, has_q1);
            if (t == 1 && !has_q1) continue;
            const bool pp = has_q1 && p.pingpong;
            float m_used = -INFINITY;

            for (int j = 0; j < n_kv; ++j, ++n_tile) {
                AT3_SEV(10);
                mbar_wait(&s_full[t], n_tile & 1);
                tc_fence_after();
                AT3_SEV(11);
                uint32_t sv[4][32];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(tmem_S + lane_addr + t * 128 + cc * 32, sv[cc]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[t]);
                AT3_SEV(12);

                const int kv_valid = p.n_tok - j * 128;
                if (kv_valid < 128) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (cc * 32 + i >= kv_valid) sv[cc][i] = 0xFF800000u;   // -inf
                }
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int i = 0; i < 32; i += 2)
                        mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(sv[cc][i]), __uint_as_float(sv[cc][i + 1]));
                const float mx = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);

                if (j == 0) {
                    m_used = mx;                          // O_t is overwritten by the first P V of the item
                } else {
                    const bool grow = mx > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {
                        // rare: rescale this warp's 32 rows of O_t (numerator and denominator column) in TMEM
                        mbar_wait(&o_full[t], (n_tile - 1) & 1);     // P(j-1) V(j-1) has landed
                        tc_fence_after();
                        const float alpha = grow ? ex2_approx((m_used - mx) * c) : 1.0f;
                        if (grow) m_used = mx;
                        attn3_rescale_rows(o_addr, alpha);
                    }
                }
                const float mc = m_used * c;
                AT3_SEV(18);
                if (j > 0) mbar_wait(&o_full[t], (n_tile - 1) & 1);   // P buffer is free once P(j-1) V(j-1) has completed

                // Ping-pong: only one warpgroup at a time is in its MUFU-bound exp phase; the other one meanwhile waits
                // for / loads its next S, reduces the row max and (below) streams P to shared memory.  mbarriers, not
                // bar.sync: BAR.SYNC.DEFER_BLOCKING does not hold back register-only ALU/MUFU work.
                AT3_SEV(13);
                if (pp) {
                    if (t == 0) {
                        if (n_pp > 0) mbar_wait(&turn[0], (n_pp - 1) & 1);
                    } else {
                        mbar_wait(&turn[1], n_pp & 1);
                    }
                }
                AT3_SEV(14);
                // P(j) = exp2((s - m) * c) as packed fp16 pairs; each finished 16-byte chunk (8 keys) goes straight to its
                // swizzled slot of the P tile so that the stores overlap the MUFU stream
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    uint32_t pk[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = g * 8 + e * 2;                      // key index inside the tile
                        const float x0 = fmaf(__uint_as_float(sv[i >> 5][i & 31]), c, -mc);
                        const float x1 = fmaf(__uint_as_float(sv[i >> 5][(i & 31) + 1]), c, -mc);
#if AT3_EXP_F32
                        pk[e] = cvt_f16x2(ex2_approx(x0), ex2_approx(x1));     // fp32 exp2, one rounding when packing
#else
                        if (AT3_POLY_MOD > 0 && ((g * 4 + e) % AT3_POLY_MOD) == AT3_POLY_MOD - 1) {
                            // this pair is exponentiated on the FMA/ALU pipes (Cody-Waite split + cubic, rel. error 7.7e-5,
                            // below the fp16 rounding of P) so that the MUFU unit is not the only exp2 engine
                            pk[e] = cvt_f16x2(exp2_poly3(x0), exp2_poly3(x1));
                        } else {
                            // ex2.approx.f16x2 (two MUFU.EX2.F16 + PRMT in SASS); the exponent argument (<= 0) is
                            // rounded to fp16 first
                            pk[e] = ex2_f16x2(cvt_f16x2(x0, x1));
                        }
#endif
                    }
                    uint8_t *dst = p_row + (g >> 3) * AT3_TILE + (((g & 7) ^ sw) << 4);
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                }
                AT3_SEV(15);
                if (pp) {
                    mbar_arrive(&turn[t ^ 1]);
                    ++n_pp;
                }
                AT3_SEV(16);
                fence_proxy_async_smem();
                mbar_arrive(&p_full[t]);
                AT3_SEV(17);
            }

            // ---- item epilogue: O_t / rowsum -> fp16 rows
            mbar_wait(&o_full[t], (n_tile - 1) & 1);
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

#undef decode
#undef AT3_SEV
#undef AT3_ISSUE_S
#undef AT3_ISSUE_PV

}  // namespace dino
