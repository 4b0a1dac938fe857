// Fused multi-head self-attention, generation 5 + intra-tile pipelining ("v8").
// Everything as attention5.cuh (TMEM ring of three 64-column slots per query tile, S computed one tile ahead, P back into
// TMEM, fp32 row sums, lazy running maximum), except for how a softmax warp walks through one 128-key tile:
//
//   v5:  load 128 scores -> wait -> row maximum over all 128 -> exponentiate
//   v8:  load the first 32 -> wait -> start the other three loads -> row maximum of the first 32 -> exponentiate them;
//        the loads complete under their first half, the maximum of the other 96 keys is reduced under their second
//        half (independent instruction streams for the scheduler), then the other three chunks are exponentiated.
//
// Only a quarter of the load latency and of the FMNMX3 reduction stays exposed (the two warpgroups run in lock-step, so
// whatever a warp cannot hide itself is time the MUFU pipe idles, see the v5 cycle trace).  The price is that the first
// chunk is exponentiated before the rest of the tile has been looked at: if the other 96 keys push a row more than 2^8
// above the reference maximum, the warp rescales O_t, the sums and the 16 packed P columns already written (rare path;
// helper shared with attention_softmax.cuh).  v7's deeper software pipeline (loads across tile boundaries) measured slower.
//
// Roles / TMEM map / barriers: see attention5.cuh.
#pragma once
#include "ptx.cuh"
#include "attention_softmax.cuh"   // chunk helpers: attn_rowmax32 / attn_mask32 / attn_exp_pairs / attn_rescale

// every AT8_POLY_MOD-th pair of probabilities is computed with a polynomial on the FMA pipe instead of MUFU (0 = none)
#ifndef AT8_POLY_MOD
#define AT8_POLY_MOD 6   // measured per ViT-L layer (B=64): 0 -> 828 us, 6 -> 780, 4 -> 828, 3 -> 857, 2 -> 906
#endif

namespace dino {

constexpr int AT8_THREADS = 384;
constexpr int AT8_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
#ifndef AT8_STAGGER
#define AT8_STAGGER 0
#endif
#ifndef AT8_KV_STAGES
#define AT8_KV_STAGES 4
#endif
constexpr int AT8_SMEM_BYTES = 2 * AT8_TILE + AT8_KV_STAGES * 2 * AT8_TILE + 256 + 1024;
constexpr float AT8_RESCALE_LOG2 = 8.0f;        // lazy-rescale threshold in the exp2 domain

// Optional cycle trace of CTA 0 (compile with -DAT8_TRACE): (event id, index, clock) per role, written to p.trace
// ([role][512][2] uint64).  Roles: 0 = MMA warp, 1 = softmax WG0 thread 0, 2 = softmax WG1 thread 0.
#ifdef AT8_TRACE
#define AT8_EV(ROLE, ID, IDX)                                                                  \
    do {                                                                                       \
        if (blockIdx.x == 0 && p.trace && tr_n < 512) {                                        \
            p.trace[((ROLE) * 512 + tr_n) * 2] = (static_cast<unsigned long long>(ID) << 32) | static_cast<unsigned>(IDX); \
            p.trace[((ROLE) * 512 + tr_n) * 2 + 1] = clock64();                                \
            ++tr_n;                                                                            \
        }                                                                                      \
    } while (0)
#else
#define AT8_EV(ROLE, ID, IDX) do {} while (0)
#endif

struct Attn8Params {
    int n_tok;
    int hidden;
    int n_heads;
    int n_qblk;        // ceil(n_tok / 256)
    int num_items;     // batch * n_heads * n_qblk
    __half *out;
    float scale_log2;  // log2(e) / sqrt(64)
    unsigned long long *trace;   // AT8_TRACE builds only
};

__global__ void __launch_bounds__(AT8_THREADS, 1)
attention_fwd_v8(const __grid_constant__ CUtensorMap tmQKV, const Attn8Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [2]
    uint8_t *sK = sQ + 2 * AT8_TILE;                      // [stages]
    uint8_t *sV = sK + AT8_KV_STAGES * AT8_TILE;          // [stages]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + AT8_KV_STAGES * AT8_TILE);
    uint64_t *q_full = bars;                              // 1
    uint64_t *q_empty = bars + 1;                         // 1
    uint64_t *kv_full = bars + 2;                         // stages
    uint64_t *kv_empty = kv_full + AT8_KV_STAGES;         // stages
    uint64_t *s_full = kv_empty + AT8_KV_STAGES;          // 2: S_t(n) is in TMEM
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

    if (warp == 10 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 11 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 1);
        for (int s = 0; s < AT8_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 128);
            mbar_init(&p_full[2 * t], 128);
            mbar_init(&p_full[2 * t + 1], 128);
            mbar_init(&o_full[t], 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_R = tmem_base;           // ring of tile t: columns 192 t + 64 slot
    const uint32_t tmem_O = tmem_base + 384;     // O_t at columns 384 + 64 t

    // work item -> (image, head, query block); consecutive items share K/V (same image and head) for L2 reuse
#define AT8_DECODE(ITEM, ROW_BASE, HEAD, Q_BASE, HAS_Q1)             \
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
                AT8_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_empty, item_ph ^ 1);           // every Q K^T of the previous item has completed
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full, (has_q1 ? 2 : 1) * AT8_TILE);
                    tma_load_2d(sQ, &tmQKV, q_full, head * 64, row_base + q_base);
                    if (has_q1) tma_load_2d(sQ + AT8_TILE, &tmQKV, q_full, head * 64, row_base + q_base + 128);
                }
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[s], 2 * AT8_TILE);
                        tma_load_2d(sK + s * AT8_TILE, &tmQKV, &kv_full[s], p.hidden + head * 64, row_base + j * 128);
                        tma_load_2d(sV + s * AT8_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + head * 64, row_base + j * 128);
                    }
                    __syncwarp();
                    if (++s == AT8_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 11) {
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
            const uint64_t q_desc1 = make_smem_desc_sw128(smem_u32(sQ + AT8_TILE), 16, 1024);
// S_t(n) = Q_t K(stage)^T into ring slots (slot, slot+1 mod 3): one N=128 MMA per k-step when the slots are adjacent,
// two N=64 MMAs (keys 0-63 -> slot 2, keys 64-127 -> slot 0; K rows 64.. start 8 KB into the tile) when the ring wraps
#define AT8_ISSUE_S(QDESC, T, CNT, SLOT, STAGE)                                                                        \
    do {                                                                                                               \
        if ((CNT) > 0) {                                                                                               \
            mbar_wait(&s_free[T], ((CNT) - 1) & 1);                                                                    \
            tc_fence_after();                                                                                          \
        }                                                                                                              \
        const uint64_t k_desc__ = make_smem_desc_sw128(smem_u32(sK + (STAGE) * AT8_TILE), 16, 1024);                   \
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
            AT8_EV(0, 1 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
        (SLOT) = (SLOT) == 2 ? 0u : (SLOT) + 1;                                                                        \
    } while (0)
// O_t (+)= P_t(n) V  (8 k-steps of 16 keys; P = 8 TMEM columns per step in ring slot n % 3), then o_full[t]
#define AT8_ISSUE_PV(VDESC, T, CNT, SLOT, J, LAST, KVS)                                                                \
    do {                                                                                                               \
        mbar_wait(&p_full[2 * (T) + ((CNT) & 1)], ((CNT) >> 1) & 1);                                                   \
        tc_fence_after();                                                                                              \
        AT8_EV(0, 7, (CNT));                                                                                           \
        if (elect_one()) {                                                                                             \
            _Pragma("unroll") for (int k = 0; k < 8; ++k)                                                              \
                umma_f16_ts(tmem_O + (T) * 64, tmem_R + (T) * 192 + (SLOT) * 64 + 8 * k,                               \
                            (VDESC) + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, ((J) | k) != 0);                \
            umma_commit(&o_full[T]);                                                                                   \
            if (LAST) umma_commit(&kv_empty[KVS]);                                                                     \
            AT8_EV(0, 3 + (T), (CNT));                                                                                 \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        (CNT)++;                                                                                                       \
        (SLOT) = (SLOT) == 2 ? 0u : (SLOT) + 1;                                                                        \
    } while (0)
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                AT8_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_full, item_ph);
                mbar_wait(&kv_full[s], ph);
                tc_fence_after();
                AT8_ISSUE_S(q_desc0, 0, ns0, ss0, s);
                if (has_q1) AT8_ISSUE_S(q_desc1, 1, ns1, ss1, s);
                if (n_kv == 1 && elect_one()) umma_commit(q_empty);
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    if (j + 1 < n_kv) {
                        int s1 = s + 1;
                        uint32_t ph1 = ph;
                        if (s1 == AT8_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                        mbar_wait(&kv_full[s1], ph1);
                        tc_fence_after();
                        AT8_EV(0, 5, j);
                        AT8_ISSUE_S(q_desc0, 0, ns0, ss0, s1);
                        if (has_q1) AT8_ISSUE_S(q_desc1, 1, ns1, ss1, s1);
                        if (j + 2 == n_kv && elect_one()) umma_commit(q_empty);   // last Q K^T of this item is in flight
                        __syncwarp();
                    }
                    // MN-major B, N = 64: a single 64-wide atom along MN (leading-dim offset unused)
                    const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * AT8_TILE), 1024, 1024);
                    // the K/V stage is released by the last P V product that reads it
                    AT8_ISSUE_PV(v_desc, 0, np0, sp0, j, !has_q1, s);
                    if (has_q1) AT8_ISSUE_PV(v_desc, 1, np1, sp1, j, true, s);
                    if (++s == AT8_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        setmaxnreg_inc<208>();
        const int t = warp >> 2;                          // query tile / warpgroup
        const int qd = warp & 3;                          // TMEM lane quarter
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ring = tmem_R + lane_addr + t * 192;          // this row's ring of three 64-column slots
        const uint32_t o_addr = tmem_O + lane_addr + t * 64;
        const float c = p.scale_log2;
        const float thr = AT8_RESCALE_LOG2 / c;           // threshold in raw-score units
        uint32_t n_tile = 0;                              // tiles processed by this warpgroup (barrier phases)
        uint32_t slot = 0;                                // n_tile % 3
        int tr_n = 0; (void) tr_n;
#define AT8_SEV(ID) do { if ((threadIdx.x & 127) == 0) AT8_EV(1 + t, ID, n_tile); } while (0)
#if AT8_STAGGER > 0
        // start the second warpgroup half a tile late: both warpgroups have the same period, so their MUFU phases stay
        // in anti-phase and one of them always feeds the pipe while the other waits for / loads / reduces its next scores
        if (t == 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < AT8_STAGGER) {}
        }
#endif

        for (int item = item_lo; item < item_hi; ++item) {
            int row_base, head, q_base;
            bool has_q1;
            AT8_DECODE(item, row_base, head, q_base, has_q1);
            if (t == 1 && !has_q1) continue;
            float m_used = -INFINITY;
            float l_run = 0.f;                            // softmax denominator relative to m_used

            for (int j = 0; j < n_kv; ++j, ++n_tile) {
                const uint32_t lo = ring + slot * 64;                          // keys 0-63 (P goes back here)
                const uint32_t hi = ring + (slot == 2 ? 0u : slot + 1) * 64;   // keys 64-127
                slot = slot == 2 ? 0u : slot + 1;
                AT8_SEV(10);
                mbar_wait(&s_full[t], n_tile & 1);
                tc_fence_after();
                AT8_SEV(11);
                uint32_t c0[32], c1[32], c2[32], c3[32];
                tmem_ld_32x32b_x32(lo, c0);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(lo + 32, c1);          // in flight under chunk 0's maximum and first exponentials
                tmem_ld_32x32b_x32(hi, c2);
                tmem_ld_32x32b_x32(hi + 32, c3);
                AT8_SEV(12);
                const int kv_valid = p.n_tok - j * 128;
                if (kv_valid < 32) attn_mask32(c0, kv_valid);
                // reference maximum for chunk 0: moves only when a row grew by more than 2^8 (then O_t and the running sum
                // are rescaled) — probabilities stay <= 256, exact in fp16
                const float mx0 = attn_rowmax32(c0);
                if (j == 0) {
                    m_used = mx0;                         // O_t is overwritten by the first P V of the item
                    l_run = 0.f;
                } else {
                    const bool grow = mx0 > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {  // rare: O_t must be quiescent, i.e. P(j-1) V(j-1) complete
                        mbar_wait(&o_full[t], (n_tile - 1) & 1);
                        tc_fence_after();
                        const float alpha = grow ? ex2_approx((m_used - mx0) * c) : 1.0f;
                        if (grow) m_used = mx0;
                        l_run *= alpha;
                        attn_rescale(o_addr, lo, alpha, true, 0);
                    }
                }
                AT8_SEV(14);
                float ls[2] = {0.f, 0.f};
                uint32_t pk[16];
                {
                    const float mc = m_used * c;
                    attn_exp_pairs<0, 8>(c0, pk, c, mc, ls);
                    // chunks 1-3 are in registers: S(n)'s second slot may be overwritten -> the MMA warp starts S(n+1)
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(&s_free[t]);
                    if (kv_valid < 128) {
                        attn_mask32(c1, kv_valid - 32);
                        attn_mask32(c2, kv_valid - 64);
                        attn_mask32(c3, kv_valid - 96);
                    }
                    const float mx123 = fmax3(attn_rowmax32(c1), attn_rowmax32(c2), attn_rowmax32(c3));
                    attn_exp_pairs<8, 16>(c0, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo, pk);
                    const bool grow = mx123 > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {  // rare: O_t, the sums and the 16 columns of P(n) already written move down
                        if (j > 0) {
                            mbar_wait(&o_full[t], (n_tile - 1) & 1);
                            tc_fence_after();
                        }
                        const float alpha = grow ? ex2_approx((m_used - mx123) * c) : 1.0f;
                        if (grow) m_used = mx123;
                        l_run *= alpha;
                        ls[0] *= alpha;
                        ls[1] *= alpha;
                        attn_rescale(o_addr, lo, alpha, j > 0, 16);
                    }
                }
                {
                    const float mc = m_used * c;
                    attn_exp_pairs<0, 16>(c1, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo + 16, pk);
                    attn_exp_pairs<0, 16>(c2, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo + 32, pk);
                    attn_exp_pairs<0, 16>(c3, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo + 48, pk);
                }
                l_run += ls[0] + ls[1];
                AT8_SEV(15);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[2 * t + (n_tile & 1)]);
                AT8_SEV(17);
            }

            // ---- item epilogue: O_t / rowsum -> fp16 rows
            mbar_wait(&o_full[t], (n_tile - 1) & 1);
            tc_fence_after();
            uint32_t a[32], b[32];
            tmem_ld_32x32b_x32(o_addr, a);
            tmem_ld_32x32b_x32(o_addr + 32, b);
            tmem_ld_wait();
            tc_fence_before();
            const int tok = q_base + t * 128 + r;
            if (tok < p.n_tok) {
                const float inv = 1.0f / l_run;
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

#undef AT8_DECODE
#undef AT8_SEV
#undef AT8_ISSUE_S
#undef AT8_ISSUE_PV

}  // namespace dino
