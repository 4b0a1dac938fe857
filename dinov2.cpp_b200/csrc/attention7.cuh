// Fused multi-head self-attention, seventh generation: v5's TMEM ring + a software-pipelined softmax warp.
// Contract as attention3.cuh (replaces the reference's mul_mat(K,Q) -> soft_max_ext -> mul_mat(V,P) chain,
// dinov2.cpp:479-543; head_dim 64, no mask).
//
// v5 (cycle trace profiles/r01_attn_v5_cycle_trace.txt) spends ~2400 cycles of every ~3350-cycle K/V step with both
// softmax warps of a scheduler in their MUFU phase (the pipe's practical limit, ~1200 cycles per 128x128 tile) and the
// other ~950 with BOTH of them waiting for / loading / max-reducing the next scores: the two warpgroups fall into
// lock-step (sharing the MUFU pipe pulls them together), so nobody feeds the pipe in between.  16 softmax warps (v6)
// or a forced stagger did not help.  v7 hides that phase inside each warp instead:
//
//  * a tile's 128 scores per row are handled as four 32-key chunks (c0..c3).  Chunks 2,3 are loaded (tcgen05.ld) while
//    chunk 0 is exponentiated, and chunks 0,1 of the NEXT tile while chunk 3 is — the loads and the FMNMX3 row-max
//    reductions of freshly loaded chunks run in the shadow of the MUFU stream of an older chunk.  s_free (which lets
//    the MMA warp start S(n+1) in the ring) is signalled once chunks 2,3 are in registers, i.e. a quarter into the
//    tile, so S(n+1) is ready when its first chunks are wanted three quarters in.
//  * the lazy running maximum works per half tile: chunks 0,1 are exponentiated against the reference maximum as of
//    their own arrival; if chunks 2,3 (or a later tile) push a row more than 2^8 above it, the owning warp rescales
//    its rows of O_t, the running sum and — mid-tile — the part of P(n) already written (all in TMEM, rare path).
//  * everything else as v5: three 64-column ring slots per query tile (S(n) in slots n%3, (n+1)%3, P(n) over slot
//    n%3), O_t in the last 2 x 64 columns, fp32 row sums in registers, 1/ATS_POLY_MOD of the exponentials on the FMA pipe.
//
// Roles: warps 0-3 / 4-7 = softmax warpgroups of query tile 0 / 1 (one thread per query row, TMEM lane quarter =
// warp % 4), warp 8 TMEM allocator, warp 9 TMA producer (Q per item, K/V ring), warps 11 / 10 MMA issuers of tile 0 / 1.
// TMEM columns: ring of tile t at 192 t (three 64-column slots), O_t at 384 + 64 t.
#pragma once
#include "ptx.cuh"
#include "attention_softmax.cuh"

namespace dino {

constexpr int AT7_THREADS = 384;
constexpr int AT7_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
#ifndef AT7_KV_STAGES
#define AT7_KV_STAGES 4
#endif
constexpr int AT7_SMEM_BYTES = 2 * AT7_TILE + AT7_KV_STAGES * 2 * AT7_TILE + 256 + 1024;
constexpr float AT7_RESCALE_LOG2 = 8.0f;        // lazy-rescale threshold in the exp2 domain

// Optional cycle trace of CTA 0 (compile with -DAT7_TRACE): (event id, index, clock) per role, written to p.trace
// ([role][512][2] uint64).  Roles: 0 = MMA warp, 1 = softmax WG0 thread 0, 2 = softmax WG1 thread 0.
#ifdef AT7_TRACE
#define AT7_EV(ROLE, ID, IDX)                                                                  \
    do {                                                                                       \
        if (blockIdx.x == 0 && p.trace && tr_n < 512) {                                        \
            p.trace[((ROLE) * 512 + tr_n) * 2] = (static_cast<unsigned long long>(ID) << 32) | static_cast<unsigned>(IDX); \
            p.trace[((ROLE) * 512 + tr_n) * 2 + 1] = clock64();                                \
            ++tr_n;                                                                            \
        }                                                                                      \
    } while (0)
#else
#define AT7_EV(ROLE, ID, IDX) do {} while (0)
#endif

struct Attn7Params {
    int n_tok;
    int hidden;
    int n_heads;
    int n_qblk;        // ceil(n_tok / 256)
    int num_items;     // batch * n_heads * n_qblk
    __half *out;
    float scale_log2;  // log2(e) / sqrt(64)
    unsigned long long *trace;   // AT7_TRACE builds only
};

__global__ void __launch_bounds__(AT7_THREADS, 1)
attention_fwd_v7(const __grid_constant__ CUtensorMap tmQKV, const Attn7Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [2]
    uint8_t *sK = sQ + 2 * AT7_TILE;                      // [stages]
    uint8_t *sV = sK + AT7_KV_STAGES * AT7_TILE;          // [stages]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sV + AT7_KV_STAGES * AT7_TILE);
    uint64_t *q_full = bars;                              // 1
    uint64_t *q_empty = bars + 1;                         // 1
    uint64_t *kv_full = bars + 2;                         // stages
    uint64_t *kv_empty = kv_full + AT7_KV_STAGES;         // stages
    uint64_t *s_full = kv_empty + AT7_KV_STAGES;          // 2: S_t(n) is in TMEM
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

    if (warp == 9 && lane == 0) prefetch_tmap(&tmQKV);
    if (warp == 11 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 2);                            // one commit per MMA warp (query tile)
        for (int s = 0; s < AT7_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 2);
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
#define AT7_DECODE(ITEM, ROW_BASE, HEAD, Q_BASE, HAS_Q1)             \
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
        if (warp == 9) {
            // ---------------------------------------------------------------- TMA producer (warp-uniform; one lane issues)
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                int row_base, head, q_base;
                bool has_q1;
                AT7_DECODE(item, row_base, head, q_base, has_q1);
                mbar_wait(q_empty, item_ph ^ 1);           // every Q K^T of the previous item has completed
                if (elect_one()) {
                    mbar_arrive_expect_tx(q_full, (has_q1 ? 2 : 1) * AT7_TILE);
                    tma_load_2d(sQ, &tmQKV, q_full, head * 64, row_base + q_base);
                    if (has_q1) tma_load_2d(sQ + AT7_TILE, &tmQKV, q_full, head * 64, row_base + q_base + 128);
                }
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[s], 2 * AT7_TILE);
                        tma_load_2d(sK + s * AT7_TILE, &tmQKV, &kv_full[s], p.hidden + head * 64, row_base + j * 128);
                        tma_load_2d(sV + s * AT7_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + head * 64, row_base + j * 128);
                    }
                    __syncwarp();
                    if (++s == AT7_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp >= 10) {
            // ---------------------------------------------------------------- MMA issuers: warp 11 = query tile 0, warp 10 = tile 1
            // One issuing warp per query tile: tcgen05.commit tracks the MMAs of the issuing thread only, so the two tiles are
            // independent pipelines (a single in-order issuer made each tile's next Q K^T wait for the OTHER warpgroup's P).
            // All 32 lanes run the control flow, barrier waits and descriptor arithmetic (warp-uniform -> uniform datapath);
            // one elected lane issues tcgen05.mma / tcgen05.commit.  The tiles of a CTA form one flat sequence over (item,
            // K/V tile) and the pipe order per query tile is always  S(next tile)  ->  P V (this tile), also across an item
            // boundary (the softmax warps prefetch the next tile's first scores before they finish P of this one).
            // A warp whose query tile does not exist in an item (last query block) only keeps the barrier protocol going.
            constexpr uint32_t idesc_s128 = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_s64 = make_idesc_f16(128, 64, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);     // A = P (TMEM), B = V, MN-major
            const int t = 11 - warp;
            const uint32_t ring = tmem_R + t * 192;
            const uint32_t o_acc = tmem_O + t * 64;
            const uint64_t q_desc = make_smem_desc_sw128(smem_u32(sQ + t * AT7_TILE), 16, 1024);
            int s = 0;
            uint32_t ph = 0, item_ph = 0;
            uint32_t ns = 0, np = 0;       // S / P tiles issued so far (barrier phases)
            uint32_t ss = 0, sp = 0;       // ring slots of the next S / next P (count % 3)
            int tr_n = 0; (void) tr_n;
// S(n) = Q_t K(stage)^T into ring slots (slot, slot+1 mod 3): one N=128 MMA per k-step when the slots are adjacent,
// two N=64 MMAs (keys 0-63 -> slot 2, keys 64-127 -> slot 0; K rows 64.. start 8 KB into the tile) when the ring wraps
#define AT7_ISSUE_S(STAGE)                                                                                             \
    do {                                                                                                               \
        if (ns > 0) {                                                                                                  \
            mbar_wait(&s_free[t], (ns - 1) & 1);                                                                       \
            tc_fence_after();                                                                                          \
        }                                                                                                              \
        const uint64_t k_desc__ = make_smem_desc_sw128(smem_u32(sK + (STAGE) * AT7_TILE), 16, 1024);                   \
        if (elect_one()) {                                                                                             \
            if (ss != 2) {                                                                                             \
                _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                          \
                    umma_f16_ss(ring + ss * 64, q_desc + 2 * k, k_desc__ + 2 * k, idesc_s128, k != 0);                 \
            } else {                                                                                                   \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                                        \
                    umma_f16_ss(ring + 128, q_desc + 2 * k, k_desc__ + 2 * k, idesc_s64, k != 0);                      \
                    umma_f16_ss(ring, q_desc + 2 * k, k_desc__ + (8192 >> 4) + 2 * k, idesc_s64, k != 0);              \
                }                                                                                                      \
            }                                                                                                          \
            umma_commit(&s_full[t]);                                                                                   \
            if (t == 0) AT7_EV(0, 1, ns);                                                                              \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        ++ns;                                                                                                          \
        ss = ss == 2 ? 0u : ss + 1;                                                                                    \
    } while (0)
            auto has_tile = [&](int it) -> bool { return t == 0 || (it % p.n_qblk) * 256 + 128 < p.n_tok; };
            bool has = false;
            if (item_lo < item_hi) {
                has = has_tile(item_lo);
                mbar_wait(q_full, 0);
                mbar_wait(&kv_full[0], 0);
                tc_fence_after();
                if (has) AT7_ISSUE_S(0);
                if (n_kv == 1 && elect_one()) umma_commit(q_empty);
                __syncwarp();
            }
            for (int item = item_lo; item < item_hi; ++item, item_ph ^= 1) {
                bool has_next = has;
                for (int j = 0; j < n_kv; ++j) {
                    const bool in_item = j + 1 < n_kv;
                    const bool cross = !in_item && item + 1 < item_hi;
                    if (in_item || cross) {
                        int s1 = s + 1;
                        uint32_t ph1 = ph;
                        if (s1 == AT7_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                        if (cross) {
                            has_next = has_tile(item + 1);
                            mbar_wait(q_full, item_ph ^ 1);            // the next item's Q tiles have landed
                        }
                        mbar_wait(&kv_full[s1], ph1);
                        tc_fence_after();
                        if (cross ? has_next : has) AT7_ISSUE_S(s1);
                        // the Q K^T just issued is the last one of its item: Q may be overwritten once it has completed
                        if ((cross ? n_kv == 1 : j + 2 == n_kv) && elect_one()) umma_commit(q_empty);
                        __syncwarp();
                    }
                    if (has) {
                        // O_t (+)= P(n) V: 8 k-steps of 16 keys; A = 8 TMEM columns of ring slot n % 3 per step, B = V (MN-major,
                        // N = 64: a single 64-wide atom along MN, leading-dim offset unused)
                        const uint64_t v_desc = make_smem_desc_sw128(smem_u32(sV + s * AT7_TILE), 1024, 1024);
                        mbar_wait(&p_full[2 * t + (np & 1)], (np >> 1) & 1);
                        tc_fence_after();
                        if (t == 0) AT7_EV(0, 7, np);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                umma_f16_ts(o_acc, ring + sp * 64 + 8 * k, v_desc + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, (j | k) != 0);
                            umma_commit(&o_full[t]);
                        }
                        __syncwarp();
                        ++np;
                        sp = sp == 2 ? 0u : sp + 1;
                    }
                    // this warp is done with the K/V stage once its MMAs so far have retired (the stage is free after both warps)
                    if (elect_one()) umma_commit(&kv_empty[s]);
                    __syncwarp();
                    if (++s == AT7_KV_STAGES) { s = 0; ph ^= 1; }
                }
                has = has_next;
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
        const float thr = AT7_RESCALE_LOG2 / c;           // threshold in raw-score units
        uint32_t n_tile = 0;                              // tiles finished by this warpgroup (barrier phases)
        uint32_t slot = 0;                                // n_tile % 3
        int tr_n = 0; (void) tr_n;
#define AT7_SEV(ID) do { if ((threadIdx.x & 127) == 0) AT7_EV(1 + t, ID, n_tile); } while (0)
        // this warpgroup's tiles form one flat sequence over (item, K/V tile); query tile 1 skips items without one
#define AT7_SKIP(ITEM) while ((ITEM) < item_hi && t == 1 && ((ITEM) % p.n_qblk) * 256 + 128 >= p.n_tok) ++(ITEM)
        const int last_valid = p.n_tok - (n_kv - 1) * 128;            // keys of the last K/V tile that exist (1..128)

        uint32_t c0[32], c1[32], c2[32], c3[32];          // the four 32-key chunks of the current tile's scores
        float m_used = -INFINITY, l_run = 0.f, mx01 = -INFINITY;
        int item = item_lo, j = 0;
        AT7_SKIP(item);
        // One loop body per tile; the pass with primed == false only prefetches chunks 0,1 of the very first tile (so that
        // c0/c1 have a single definition point in the loop — two made ptxas shuffle 64 registers through local memory).
        bool primed = false;
        while (item < item_hi) {
            const uint32_t lo = ring + slot * 64;                          // keys 0-63 of S(n); P(n) goes back here
            const uint32_t hi = ring + (slot == 2 ? 0u : slot + 1) * 64;   // keys 64-127 of S(n) = keys 0-63 of S(n+1)
            const bool last_kv = primed && j == n_kv - 1;
            float ls[2] = {0.f, 0.f};
            uint32_t pk[16];
            if (primed) {
                // A. reference maximum for chunks 0,1 (before the next loads are in flight: few registers live at the rare call)
                if (j == 0) {
                    m_used = mx01;                        // first tile of an item: O_t is overwritten by its first P V
                    l_run = 0.f;
                } else {
                    const bool grow = mx01 > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {  // rare
                        mbar_wait(&o_full[t], (n_tile - 1) & 1);      // P(n-1) V(n-1) complete: O_t is quiescent
                        tc_fence_after();
                        const float alpha = grow ? ex2_approx((m_used - mx01) * c) : 1.0f;
                        if (grow) m_used = mx01;
                        l_run *= alpha;
                        attn_rescale(o_addr, lo, alpha, true, 0);
                    }
                }
                // B. chunks 2,3 start loading (S(n) is complete: chunks 0,1 came out of it)
                tmem_ld_32x32b_x32(hi, c2);
                tmem_ld_32x32b_x32(hi + 32, c3);
                AT7_SEV(14);
                // C. chunk 0 (hides the loads of chunks 2,3)
                {
                    const float mc = m_used * c;
                    attn_exp_pairs<0, 16>(c0, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo, pk);
                }
                // D. chunks 2,3 are in registers: S(n)'s second slot may be overwritten -> the MMA warp starts S(n+1)
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(&s_free[t]);
                if (last_kv) {
                    attn_mask32(c2, last_valid - 64);
                    attn_mask32(c3, last_valid - 96);
                }
                // E. chunk 1, with the row maximum of chunks 2,3 reduced in its shadow
                const float mx23 = fmaxf(attn_rowmax32(c2), attn_rowmax32(c3));
                {
                    const float mc = m_used * c;
                    attn_exp_pairs<0, 16>(c1, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo + 16, pk);
                }
                {
                    const bool grow = mx23 > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {  // rare: O_t, the sums and the half of P(n) already written move down
                        if (j > 0) {
                            mbar_wait(&o_full[t], (n_tile - 1) & 1);
                            tc_fence_after();
                        }
                        const float alpha = grow ? ex2_approx((m_used - mx23) * c) : 1.0f;
                        if (grow) m_used = mx23;
                        l_run *= alpha;
                        ls[0] *= alpha;
                        ls[1] *= alpha;
                        attn_rescale(o_addr, lo, alpha, j > 0, 32);
                    }
                }
                // F. chunk 2
                {
                    const float mc = m_used * c;
                    attn_exp_pairs<0, 16>(c2, pk, c, mc, ls);
                    tmem_st_32x32b_x16(lo + 32, pk);
                }
            }
            // next tile of this warpgroup (the very first one in the priming pass)
            int item_n = item, j_n = primed ? j + 1 : 0;
            if (last_kv) {
                item_n = item + 1;
                j_n = 0;
                AT7_SKIP(item_n);
            }
            const bool has_next = item_n < item_hi;
            // Prefetch only a tile whose Q K^T the MMA warp issues BEFORE this tile's P V.  Query tile 1 skips items that
            // have none: the tile after such a gap is issued much later, waiting for it here (before P of this tile is
            // signalled) would deadlock — prime the pipeline again instead.
            const bool prefetch = has_next && (!last_kv || item_n == item + 1);
            // G. its chunks 0,1 start loading (S(n+1) was started at D, about one and a half chunks ago)
            if (prefetch) {
                AT7_SEV(10);
                mbar_wait(&s_full[t], (primed ? n_tile + 1 : n_tile) & 1);
                tc_fence_after();
                AT7_SEV(11);
                const uint32_t src = primed ? hi : lo;
                tmem_ld_32x32b_x32(src, c0);
                tmem_ld_32x32b_x32(src + 32, c1);
            }
            // H. chunk 3: its first half hides the loads, its second half the row maximum of the new chunks
            const float mc3 = m_used * c;
            if (primed) attn_exp_pairs<0, 8>(c3, pk, c, mc3, ls);
            tmem_ld_wait();
            if (prefetch) {
                if (j_n == n_kv - 1) {
                    attn_mask32(c0, last_valid);
                    attn_mask32(c1, last_valid - 32);
                }
                mx01 = fmaxf(attn_rowmax32(c0), attn_rowmax32(c1));
            }
            if (primed) {
                attn_exp_pairs<8, 16>(c3, pk, c, mc3, ls);
                tmem_st_32x32b_x16(lo + 48, pk);
                l_run += ls[0] + ls[1];
                // I. P(n) complete
                AT7_SEV(15);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[2 * t + (n_tile & 1)]);
                AT7_SEV(17);
                ++n_tile;
                slot = slot == 2 ? 0u : slot + 1;

                if (last_kv) {
                    // ---- item epilogue: O_t / rowsum -> fp16 rows
                    int row_base, head, q_base;
                    bool has_q1;
                    AT7_DECODE(item, row_base, head, q_base, has_q1);
                    (void) has_q1;
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
            if (!has_next) break;
            item = item_n;
            j = j_n;
            primed = prefetch;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

#undef AT7_DECODE
#undef AT7_SEV
#undef AT7_SKIP
#undef AT7_ISSUE_S

}  // namespace dino
