// Fused multi-head self-attention, sixth generation: v5's TMEM ring with FOUR softmax warps per scheduler.
// Contract as attention3.cuh (replaces the reference's mul_mat(K,Q) -> soft_max_ext -> mul_mat(V,P) chain,
// dinov2.cpp:479-543; head_dim 64, no mask).
//
// Why: at head_dim 64 the kernel is bound by the MUFU pipe (16 exp2/clk/SM: 1024 cycles per 128x128 tile against 512 cycles
// of tensor work).  A warp issues in order, so one warp alone only reaches 1 MUFU per ~13 cycles (1600 cycles per tile);
// v5's two softmax warps per scheduler reached ~1200 cycles per tile in their common exp phase and left the pipe idle
// while both reduced the next row maximum (cycle trace profiles/r01_attn_v5_cycle_trace.txt).  Here every query row is
// shared by two threads of different warps (keys 0-63 / 64-127 of each K/V tile), i.e. 16 softmax warps = 4 per
// scheduler: whenever one warp waits, loads scores or reduces a maximum, three others keep the MUFU pipe fed.
//
//  * TMEM as v5: ring of three 64-column slots per query tile (S(n) in slots n%3 and (n+1)%3, P(n) over slot n%3,
//    so S(n+1) is computed while S(n) is exponentiated), O_t in the last 2 x 64 columns.  Half h of a row reads slot
//    (n+h)%3 and writes columns [32h, 32h+32) of P(n).
//  * The two halves of a row agree on the row maximum through shared memory and a 64-thread named barrier per tile
//    (which also orders "both halves have loaded S" before either overwrites slot n%3 with P); each keeps its own
//    partial row sum, added once per item.  Lazy running max as v3/v5 (rescale only after growth by more than 2^8);
//    each half rescales its 32 columns of O_t.
//
// Roles (640 threads): warps 0-15 softmax (warp w: TMEM lane quarter w%4, query tile (w/4)/2, key half (w/4)%2),
// warp 16 TMEM allocator, warp 18 TMA producer (Q per item, K/V ring), warp 19 MMA issuer.
#pragma once
#include "ptx.cuh"

// every AT6_POLY_MOD-th pair of probabilities is computed with a polynomial on the FMA pipe instead of MUFU (0 = none)
#ifndef AT6_POLY_MOD
#define AT6_POLY_MOD 0
#endif

namespace dino {

constexpr int AT6_THREADS = 640;
constexpr int AT6_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
#ifndef AT6_KV_STAGES
#define AT6_KV_STAGES 4
#endif
constexpr int AT6_XCH_BYTES = (2 * 2 * 2 * 128 + 2 * 2 * 128) * 4;   // row-max exchange (double buffered) + row-sum exchange
constexpr int AT6_SMEM_BYTES = 2 * AT6_TILE + AT6_KV_STAGES * 2 * AT6_TILE + AT6_XCH_BYTES + 256 + 1024;
constexpr float AT6_RESCALE_LOG2 = 8.0f;        // lazy-rescale threshold in the exp2 domain

// exp2(x) without the MUFU unit: round-to-nearest split x = n + f (magic-number add), cubic minimax for 2^f on
// [-0.5, 0.5], exponent field patched by integer add.  Arguments below -30 (masked keys are -inf) clamp to 2^-30, which
// is zero once P is rounded to fp16.
__device__ __forceinline__ float exp2_poly3_v6(float x) {
    const float t = fmaxf(x, -30.0f);
    const float u = t + 12582912.0f;                 // 1.5 * 2^23: low mantissa bits now hold round(t)
    const float f = t - (u - 12582912.0f);
    float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(u) << 23));
}

// Rare path of the lazy running-max correction: scale this thread's 32 columns of its O_t row in TMEM.
__device__ __noinline__ void attn6_rescale_rows(uint32_t o_addr, float alpha) {
    uint32_t a[32];
    tmem_ld_32x32b_x32(o_addr, a);
    tmem_ld_wait();
#pragma unroll
    for (int d = 0; d < 32; ++d) a[d] = __float_as_uint(__uint_as_float(a[d]) * alpha);
    tmem_st_32x32b_x32(o_addr, a);
    tmem_st_wait();
}

// Optional cycle trace of CTA 0 (compile with -DAT6_TRACE): (event id, index, clock) per role, written to p.trace
// ([role][512][2] uint64).  Roles: 0 = MMA warp, 1 = thread 0 of warp 0 (tile 0, half 0), 2 = thread 0 of warp 8 (tile 1, half 0).
#ifdef AT6_TRACE
#define AT6_EV(ROLE, ID, IDX)                                                                  \
    do {                                                                                       \
        if (blockIdx.x == 0 && p.trace && tr_n < 512) {                                        \
            p.trace[((ROLE) * 512 + tr_n) * 2] = (static_cast<unsigned long long>(ID) << 32) | static_cast<unsigned>(IDX); \
            p.trace[((ROLE) * 512 + tr_n) * 2 + 1] = clock64();                                \
            ++tr_n;                                                                            \
        }                                                                                      \
    } while (0)
#else
#define AT6_EV(ROLE, ID, IDX) do {} while (0)
#endif

struct Attn6Params {
    int n_tok;
    int hidden;
    int n_heads;
    int n_qblk;        // ceil(n_tok / 256)
    int num_items;     // batch * n_heads * n_qblk
    __half *out;
    float scale_log2;  // log2(e) / sqrt(64)
    unsigned long long *trace;   // AT6_TRACE builds only
};

__global__ void __launch_bounds__(AT6_THREADS, 1)
attention_fwd_v6(const __grid_constant__ CUtensorMap tmQKV, const Attn6Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [2]
    uint8_t *sK = sQ + 2 * AT6_TILE;                      // [stages]
    uint8_t *sV = sK + AT6_KV_STAGES * AT6_TILE;          // [stages]
    float *xmax = reinterpret_cast<float *>(sV + AT6_KV_STAGES * AT6_TILE);   // [parity][tile][half][row]
    float *xsum = xmax + 2 * 2 * 2 * 128;                                       // [tile][half][row]
    uint64_t *bars = reinterpret_cast<uint64_t *>(xsum + 2 * 2 * 128);
    uint64_t *q_full = bars;                              // 1
    uint64_t *q_empty = bars + 1;                         // 1
    uint64_t *kv_full = bars + 2;                         // stages
    uint64_t *kv_empty = kv_full + AT6_KV_STAGES;         // stages
    uint64_t *s_full = kv_empty + AT6_KV_STAGES;          // 2: S_t(n) is in TMEM
    uint64_t *s_free = s_full + 2;                        // 2: S_t(n) is in registers (its second slot may be overwritten)
    // P_t(n) is in TMEM.  Two barriers per tile, used alternately: a warpgroup may finish P_t(n+1) before the MMA warp (held
    // up by the other tile) has looked at P_t(n) — with a single barrier that is two phase flips and the parity wait never
    // returns.  It cannot be two tiles ahead: S_t(n+2) is only issued after the MMA warp has consumed P_t(n).
    uint64_t *p_full = s_free + 2;                        // 2 x 2
    uint64_t *o_full = p_full + 4;                        // 2: P_t(n) V has completed
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(o_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = (p.n_tok + 127) / 128;
    // contiguous, balanced item range of this CTA: consecutive items share K/V (same image and head), so a CTA re-reads
    // them from L2, and every CTA gets the same mix of full and half (single query tile) blocks
    const int item_lo = static_cast<int>(static_cast<long long>(p.num_items) * blockIdx.x / gridDim.x);
    const int item_hi = static_cast<int>(static_cast<long long>(p.num_items) * (blockIdx.x + 1) / gridDim.x);

    if (warp == 18 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 19 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < AT6_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 256);
            mbar_init(&p_full[2 * t], 256);
            mbar_init(&p_full[2 * t + 1], 256);
            mbar_init(&o_full[t], 1);
        }
        fence_mbar_init();
    }
    if (warp == 16) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_R = tmem_base;           // ring of tile t: columns 192 t + 64 slot
    const uint32_t tmem_O = tmem_base + 384;     // O_t at columns 384 + 64 t

    // work item -> (image, head, query block); consecutive items share K/V (same image and head) for L2 reuse
#define AT6_DECODE(ITEM, ROW_BASE, HEAD, Q_BASE, HAS_Q1)             \
    do {                                                            \
        const int qb__ = (ITEM) % p.n_qblk;                         \
        const int ih__ = (ITEM) / p.n_qblk;                         \
        (HEAD) = ih__ % p.n_heads;                                  \
        (ROW_BASE) = (ih__ / p.n_heads) * p.n_tok;                  \
        (Q_BASE) = qb__ * 256;                                      \
        (HAS_Q1) = (Q_BASE) + 128 < p.n_tok;                        \
    } while (0)

    if (warp >= 16) {
        setmaxnreg_dec<48>();
        if (warp == 18) {
            // ---------------------------------------------------------------- TMA producer (warp-uniform; one lane issues)
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                AT6_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_empty, item_ph ^ 1);           // every Q K^T of the previous item has completed
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full, (has_q1 ? 2 : 1) * AT6_TILE);
                    tma_load_2d(sQ, &tmQKV, q_full, head * 64, row_base + q_base);
                    if (has_q1) tma_load_2d(sQ + AT6_TILE, &tmQKV, q_full, head * 64, row_base + q_base + 128);
                }
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[s], 2 * AT6_TILE);
                        tma_load_2d(sK + s * AT6_TILE, &tmQKV, &kv_full[s], p.hidden + head * 64, row_base + j * 128);
                        tma_load_2d(sV + s * AT6_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + head * 64, row_base + j * 128);
                    }
                    __syncwarp();
                    if (++s == AT6_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 19) {
            // ---------------------------------------------------------------- MMA issuer
            // All 32 lanes run the control flow, barrier waits and descriptor arithmetic (warp-uniform -> uniform
            // datapath); one elected lane issues tcgen05.mma / tcgen05.commit.
            constexpr uint32_t idesc_s128 = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_s64 = make_idesc_f16(128, 64, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);     // A = P (TMEM), B = V, MN-major
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            // S / P tiles issued so far per query tile: ring slot = count % 3 (kept as a separate counter), barrier phase = count & 1
            uint32_t ns0 = 0, ns1 = 0, np0 = 0, np1 = 0;
            uint32_t ss0 = 0, ss1 = 0, sp0 = 0, sp1 = 0;   // ring slots of the next S / next P
            int tr_n = 0; (void) tr_n;
            const uint64_t q_desc0 = make_smem_desc_sw128(smem_u32(sQ), 16, 1024);
            const uint64_t q_desc1 = make_smem_desc_sw128(smem_u32(sQ + AT6_TILE), 16, 1024);
// S_t(n) = Q_t K(stage)^T into ring slots (slot, slot+1 mod 3): one N=128 MMA per k-step when the slots are adjacent,
// two N=64 MMAs (keys 0-63 -> slot 2, keys 64-127 -> slot 0; K rows 64.. start 8 KB into the tile) when the ring wraps
#define AT6_ISSUE_S(QDESC, T, CNT, SLOT, STAGE)                                                                        \
    do {                                                                                                               \
        if ((CNT) > 0) {                                                                                               \
            mbar_wait(&s_free[T], ((CNT) - 1) & 1);                                                                    \
            tc_fence_after();                                                                                          \
        }                                                                                                              \
        const uint64_t k_desc__ = make_smem_desc_sw128(smem_u32(sK + (STAGE) * AT6_TILE), 16, 1024);                   \
        const uint32_t ring__ = tmem_R + (T) * 192;                                                                    \
        if (elect_one()) {                                                                                             \
            if ((SLOT) != 2) {                                                                                         \
                _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                          \
                    umma_f16_ss(ring__ + (SLOT) * 64, (QDESC) + 2 * k, k_desc__ + 2 * k, idesc_s128, k != 0);          \
            } else {                                                                                                   \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                                        \
                    umma_f16_ss(ring__ + 128, (QDESC) + 2 * k, k_desc__ + 2 * k, idesc_s64, k != 0);                   \
                    umma_f16_ss(ring__, (QDESC) + 2 * k, k_desc__ + (8192 >> 4) + 2 * k, idesc_s64, k != 0);           \
                }                                                                                                      \
            }                                                                                                          \
            umma_commit(&s_full[T]);                                                                                   \
            AT6_EV(0, 1 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
        (SLOT) = (SLOT) == 2 ? 0u : (SLOT) + 1;                                                                        \
    } while (0)
// O_t (+)= P_t(n) V  (8 k-steps of 16 keys; P = 8 TMEM columns per step in ring slot n % 3), then o_full[t]
#define AT6_ISSUE_PV(VDESC, T, CNT, SLOT, J, LAST, KVS)                                                                \
    do {                                                                                                               \
        mbar_wait(&p_full[2 * (T) + ((CNT) & 1)], ((CNT) >> 1) & 1);                                                   \
        tc_fence_after();                                                                                              \
        AT6_EV(0, 7, (CNT));                                                                                           \
        if (elect_one()) {                                                                                             \
            _Pragma("unroll") for (int k = 0; k < 8; ++k)                                                              \
                umma_f16_ts(tmem_O + (T) * 64, tmem_R + (T) * 192 + (SLOT) * 64 + 8 * k,                               \
                            (VDESC) + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, ((J) | k) != 0);                \
            umma_commit(&o_full[T]);                                                                                   \
            if (LAST) umma_commit(&kv_empty[KVS]);                                                                     \
            AT6_EV(0, 3 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
        (SLOT) = (SLOT) == 2 ? 0u : (SLOT) + 1;                                                                        \
    } while (0)
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                AT6_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_full, item_ph);
                mbar_wait(&kv_full[s], ph);
                tc_fence_after();
                AT6_ISSUE_S(q_desc0, 0, ns0, ss0, s);
                if (has_q1) AT6_ISSUE_S(q_desc1, 1, ns1, ss1, s);
                if (n_kv == 1 && elect_one()) umma_commit(q_empty);
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) {
                        int s1 = s + 1;
                        uint32_t ph1 = ph;
                        if (s1 == AT6_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                        mbar_wait(&kv_full[s1], ph1);
                        tc_fence_after();
                        AT6_EV(0, 5, j);
                        AT6_ISSUE_S(q_desc0, 0, ns0, ss0, s1);
                        if (has_q1) AT6_ISSUE_S(q_desc1, 1, ns1, ss1, s1);
                        if (j + 2 == n_kv && elect_one()) umma_commit(q_empty);   // last Q K^T of this item is in flight
                        __syncwarp();
                    }
                    // MN-major B, N = 64: a single 64-wide atom along MN (leading-dim offset unused)
                    const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * AT6_TILE), 1024, 1024);
                    // the K/V stage is released by the last P V product that reads it
                    AT6_ISSUE_PV(v_desc, 0, np0, sp0, j, !has_q1, s);
                    if (has_q1) AT6_ISSUE_PV(v_desc, 1, np1, sp1, j, true, s);
                    if (++s == AT6_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        setmaxnreg_inc<104>();
        const int qd = warp & 3;                          // TMEM lane quarter
        const int t = warp >> 3;                          // query tile
        const int h = (warp >> 2) & 1;                    // key half of every K/V tile (and dim half of the output row)
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t pair_bar = 1 + t * 4 + qd;         // named barrier shared with the warp that owns the other half
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ring = tmem_R + lane_addr + t * 192;          // this row's ring of three 64-column slots
        const uint32_t o_addr = tmem_O + lane_addr + t * 64 + h * 32;
        const float c = p.scale_log2;
        const float thr = AT6_RESCALE_LOG2 / c;           // threshold in raw-score units
        uint32_t n_tile = 0;                              // tiles processed by this warp (barrier phases)
        uint32_t slot = 0;                                // n_tile % 3
        int tr_n = 0; (void) tr_n;
#define AT6_SEV(ID) do { if ((threadIdx.x & 255) == 0) AT6_EV(1 + t, ID, n_tile); } while (0)

        for (int item = item_lo; item < item_hi; ++item) {
            int row_base, head, q_base;
            bool has_q1;
            AT6_DECODE(item, row_base, head, q_base, has_q1);
            if (t == 1 && !has_q1) continue;
            float m_used = -INFINITY;
            float l_run = 0.f;                            // this half's part of the softmax denominator, relative to m_used

            for (int j = 0; j < n_kv; ++j, ++n_tile) {
                const uint32_t lo = ring + slot * 64;                          // keys 0-63 (P goes back here)
                const uint32_t hi = ring + (slot == 2 ? 0u : slot + 1) * 64;   // keys 64-127
                slot = slot == 2 ? 0u : slot + 1;
                const uint32_t src = h ? hi : lo;
                AT6_SEV(10);
                mbar_wait(&s_full[t], n_tile & 1);
                tc_fence_after();
                AT6_SEV(11);
                uint32_t sv[2][32];
                tmem_ld_32x32b_x32(src, sv[0]);
                tmem_ld_32x32b_x32(src + 32, sv[1]);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[t]);
                AT6_SEV(12);

                const int kv_valid = p.n_tok - j * 128 - h * 64;
                if (kv_valid < 64) {
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (cc * 32 + i >= kv_valid) sv[cc][i] = 0xFF800000u;   // -inf
                }
                // row maximum of this tile: own half, then the other half's through shared memory.  The pair barrier also
                // guarantees that both halves hold their scores in registers before P overwrites slot `lo`.
                float mx;
                {
                    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            mx4[(i >> 1) & 3] = fmax3(mx4[(i >> 1) & 3], __uint_as_float(sv[cc][i]), __uint_as_float(sv[cc][i + 1]));
                    const float mine = fmaxf(fmax3(mx4[0], mx4[1], mx4[2]), mx4[3]);
                    float *slot_x = xmax + (((n_tile & 1) * 2 + t) * 2) * 128;
                    slot_x[h * 128 + r] = mine;
                    named_bar_sync(pair_bar, 64);
                    mx = fmaxf(mine, slot_x[(h ^ 1) * 128 + r]);
                }
                // the reference maximum only moves (and O_t / the running sum are rescaled) when some row grew by more than
                // 2^8 — probabilities then stay <= 256, exact in fp16.  Both halves of a row take the same decision.
                if (j == 0) {
                    m_used = mx;                          // O_t is overwritten by the first P V of the item
                } else {
                    const bool grow = mx > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {
                        // rare: O_t must be quiescent, i.e. P(j-1) V(j-1) complete
                        mbar_wait(&o_full[t], (n_tile - 1) & 1);
                        tc_fence_after();
                        const float alpha = grow ? ex2_approx((m_used - mx) * c) : 1.0f;
                        if (grow) m_used = mx;
                        l_run *= alpha;
                        attn6_rescale_rows(o_addr, alpha);
                    }
                }
                AT6_SEV(14);
                // P(j) = exp2((s - m_used) * c): fp32 exp2 on the MUFU pipe, fp32 row sum, one rounding to packed fp16 when
                // written back to columns [32h, 32h+32) of ring slot `lo`, 32 keys at a time
                {
                    const float mc = m_used * c;
                    float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        uint32_t pk[16];
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float x0 = fmaf(__uint_as_float(sv[g][2 * e]), c, -mc);
                            const float x1 = fmaf(__uint_as_float(sv[g][2 * e + 1]), c, -mc);
                            float p0, p1;
                            if (AT6_POLY_MOD > 0 && ((g * 16 + e) % (AT6_POLY_MOD > 0 ? AT6_POLY_MOD : 1)) == AT6_POLY_MOD - 1) {
                                // exponentiated on the FMA/ALU pipes (Cody-Waite split + cubic, rel. error 7.7e-5, below
                                // the fp16 rounding of P) so that the MUFU unit is not the only exp2 engine
                                p0 = exp2_poly3_v6(x0);
                                p1 = exp2_poly3_v6(x1);
                            } else {
                                p0 = ex2_approx(x0);
                                p1 = ex2_approx(x1);
                            }
                            ls4[e & 3] += p0 + p1;
                            pk[e] = cvt_f16x2(p0, p1);
                        }
                        tmem_st_32x32b_x16(lo + h * 32 + g * 16, pk);
                    }
                    l_run += (ls4[0] + ls4[1]) + (ls4[2] + ls4[3]);
                }
                AT6_SEV(15);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[2 * t + (n_tile & 1)]);
                AT6_SEV(17);
            }

            // ---- item epilogue: O_t / rowsum -> fp16; this thread writes dims [32h, 32h+32) of its row
            mbar_wait(&o_full[t], (n_tile - 1) & 1);
            tc_fence_after();
            uint32_t a[32];
            tmem_ld_32x32b_x32(o_addr, a);
            tmem_ld_wait();
            tc_fence_before();
            xsum[(t * 2 + h) * 128 + r] = l_run;
            named_bar_sync(pair_bar, 64);
            const float l_row = l_run + xsum[(t * 2 + (h ^ 1)) * 128 + r];
            const int tok = q_base + t * 128 + r;
            if (tok < p.n_tok) {
                const float inv = 1.0f / l_row;
                uint4 *dst = reinterpret_cast<uint4 *>(p.out + static_cast<size_t>(row_base + tok) * p.hidden + head * 64 + h * 32);
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    dst[v] = make_uint4(pack_half2(__uint_as_float(a[8 * v]) * inv, __uint_as_float(a[8 * v + 1]) * inv),
                                        pack_half2(__uint_as_float(a[8 * v + 2]) * inv, __uint_as_float(a[8 * v + 3]) * inv),
                                        pack_half2(__uint_as_float(a[8 * v + 4]) * inv, __uint_as_float(a[8 * v + 5]) * inv),
                                        pack_half2(__uint_as_float(a[8 * v + 6]) * inv, __uint_as_float(a[8 * v + 7]) * inv));
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

#undef AT6_DECODE
#undef AT6_SEV
#undef AT6_ISSUE_S
#undef AT6_ISSUE_PV

}  // namespace dino
