// Fused multi-head self-attention, generation 10: v8's tile loop, run as ONE CONTINUOUS STREAM of key tiles across work items.
// Semantics: reference soft_max_ext(K Q^T / 8) V per head (dinov2.cpp:527-543, ops.cpp:4641-4737); fp16 operands, fp32
// accumulation, exp2 with a lazily updated reference maximum (attention_softmax.cuh).
//
// What the round-2 ncu source page + cycle trace of v8 showed (profiles/r02_ncu_attn.md): inside an item the softmax
// warps run at ~2900 cycles per 128-key tile, but every item boundary (each 11 tiles at 518 px) costs another ~4300 cycles:
//   * the epilogue waits for the last P V, then every thread writes its 128-byte output row with 8 STG.128 whose 32 lanes
//     hit 32 different lines (256 single-sector wavefronts per warp and item),
//   * only then is the next item's first S = Q K^T waited for, which the MMA warp could not issue earlier (one Q buffer),
//   * the item decode does two integer divisions per role.
// v10 removes the boundary from the MUFU stream:
//   1. Q is double-buffered; the MMA warp issues the NEXT item's first S while the current item's last tile is being
//      exponentiated (the tile loop of all three roles just keeps going: stage ring, TMEM ring and barrier phases never reset),
//   2. the epilogue of item i is deferred into the first tile of item i+1 (after its second chunk, when 64 score registers
//      are free): O_t is only overwritten by that tile's P V, which is issued after the warpgroup has arrived on p_full,
//   3. the output tile goes through a 128B-swizzled shared-memory staging tile and ONE TMA store per warpgroup and item
//      (3-D tensor map [image, token, channel]: rows past the image's last token are clipped by the hardware),
//   4. items are decoded incrementally (no divisions after the first).
//
//   5. ONE MMA-ISSUING WARP PER QUERY TILE (warps 11 and 10).  The ncu source page of v8 shows its single issuer busy for ~70 %
//      of every tile period (~350 instructions: barrier-wait loops, elect / reconverge, descriptor arithmetic on the uniform
//      datapath, 24-32 tcgen05.mma, 6 commits), and strictly in order S0, S1, PV0, PV1: the next S of tile 0 queued behind
//      the P V of tile 1, and the softmax warps spun ~350 cycles per tile on s_full.  The hazards of the TMEM ring are per
//      query tile, so each warp's in-order tcgen05 pipeline still covers them; shared K/V stages and Q buffers are released
//      by barriers that count both warps.
//
// Roles (384 threads): warps 0-3 / 4-7 softmax warpgroups (query tile 0 / 1, one thread per row), warp 8 TMEM allocator,
// warp 9 TMA producer, warps 11 / 10 MMA issuers of query tile 0 / 1.  TMEM: per query tile a ring of three 64-column slots (S(n) in slots n%3,
// (n+1)%3; P(n) written back over slot n%3 as the TMEM A operand of P V) at columns 192 t, O_t at columns 384 + 64 t.
#pragma once
#include "ptx.cuh"
#include "attention_softmax.cuh"   // chunk helpers: attn_rowmax32 / attn_mask32 / attn_exp_pairs / attn_rescale
#include <type_traits>

namespace dino {

// AT10_SPLIT = 1: TWO THREADS PER QUERY ROW.  Four softmax warpgroups instead of two: warps 0-3 / 4-7 take keys 0-63 of every
// 128-key tile of query tile 0 / 1, warps 8-11 / 12-15 keys 64-127 (TMEM lane quarter = warp & 3 either way).  The two halves of a
// row agree on the reference maximum through shared memory once per tile (one named barrier per query tile), keep separate row
// sums and each own 32 of the 64 output columns.  Why: with one thread per row there are two softmax warps per scheduler, each
// an in-order stream of ~85 MUFU (8-cycle dispatch) + ~600 other instructions per tile — the XU, FMA and ALU pipes idle in turn.
// Four warps per scheduler with half the work each overlap those streams.  Helper warps move to 16-19.
#ifndef AT10_SPLIT
#define AT10_SPLIT 0
#endif
constexpr int AT10_HW = AT10_SPLIT ? 16 : 8;     // first helper warp: +0 TMEM allocator, +1 TMA producer, +2 / +3 MMA issuers of query tile 1 / 0
constexpr int AT10_THREADS = (AT10_HW + 4) * 32;
constexpr int AT10_SM_THREADS = AT10_SPLIT ? 256 : 128;   // softmax threads per query tile
constexpr int AT10_TILE = 128 * 64 * 2;          // 16 KB: a 128 x 64 fp16 tile
#ifndef AT10_KV_STAGES
#define AT10_KV_STAGES (AT10_SPLIT ? 3 : 4)      // 3 measured equal to 4 (619 vs 620 us); the split variant needs the space for its exchange buffers
#endif
constexpr int AT10_XCHG_BYTES = AT10_SPLIT ? 2 * 2 * 2 * 128 * 4 : 0;   // [max | sum][query tile][half][row] floats
#ifndef AT10_STAGGER
#define AT10_STAGGER 0
#endif
// 1: the two softmax warpgroups take turns on the exponential phase of a tile (named-barrier baton, as in FlashAttention-3's
// ping-pong schedule).  One warp per scheduler saturates the XU pipe on its own (85 MUFU.EX2 + 64 F2FP per 128 keys = 1192 cycles,
// measured 1192); free-running, the warpgroups fall into lock-step within ~10 tiles, exponentiate at the same time (2000-2200
// cycles each) and leave the pipe idle while both wait for / load / max-reduce their next scores (~900 cycles per tile).
#ifndef AT10_PINGPONG
#define AT10_PINGPONG 0     // measured: strict alternation costs 684 us per ViT-L layer against 616 free-running — see DESIGN.md
#endif
// 1: fp16 rounding of the MUFU pairs by integer arithmetic instead of F2FP (attn_exp_pairs_ip in attention_softmax.cuh), which
// takes a third of the work off the XU pipe.  Same accuracy, measured SLOWER (659 vs 634 us per ViT-L layer; without any polynomial
// pairs 706): two in-order warps per scheduler are bound by their own issue timeline — every instruction added costs, whichever
// pipe it runs on — not by the XU pipe alone.  Off.
#ifndef AT10_INTPACK
#define AT10_INTPACK 0
#endif
// 1: chunks that lie entirely beyond the image's last token are neither max-reduced nor exponentiated.  As branches in EVERY tile
// this measured 7 % slower (724 vs 675 us per ViT-L layer: they cut the body into separate scheduling regions for ptxas); with the
// tile body specialised per position they only exist in the item's last tile: 647 -> 636 us.
#ifndef AT10_SKIP_MASKED
#define AT10_SKIP_MASKED 1
#endif
// How the softmax warps learn that P(n-1) V(n-1) has completed (growth path, epilogue) — see the race described at the end of the
// tile body.  2 (default): the query tile's MMA warp waits for every product it has issued, one tile later, and publishes the
// count in shared memory; the softmax warps poll the count (no wait in the common path: 632 us per ViT-L layer).  1: every
// softmax warp waits on o_full at the end of every tile (642-658 us).  0: parity wait in the rare path only — the race.
#ifndef AT10_OBSERVE_EVERY_PV
#define AT10_OBSERVE_EVERY_PV 2
#endif
// Q: 2 buffers x 2 tiles; K, V: stages; O staging: 2 tiles; barriers; alignment slack
constexpr int AT10_SMEM_BYTES = 4 * AT10_TILE + AT10_KV_STAGES * 2 * AT10_TILE + 2 * AT10_TILE + 256 + AT10_XCHG_BYTES + 1024;
#ifndef AT10_RESCALE_LOG2_VALUE
#define AT10_RESCALE_LOG2_VALUE 8.0f
#endif
constexpr float AT10_RESCALE_LOG2 = AT10_RESCALE_LOG2_VALUE;        // lazy-rescale threshold in the exp2 domain

// Optional cycle trace of CTA 0 (compile with -DAT10_TRACE): (event id, index, clock) per role, written to p.trace
// ([role][512][2] uint64).  Roles: 0 = MMA warp, 1 = softmax WG0 thread 0, 2 = softmax WG1 thread 0.
#ifdef AT10_TRACE
#define AT10_EV(ROLE, ID, IDX)                                                                 \
    do {                                                                                       \
        if (blockIdx.x == 0 && p.trace && tr_n < 512) {                                        \
            p.trace[((ROLE) * 512 + tr_n) * 2] = (static_cast<unsigned long long>(ID) << 32) | static_cast<unsigned>(IDX); \
            p.trace[((ROLE) * 512 + tr_n) * 2 + 1] = clock64();                                \
            ++tr_n;                                                                            \
        }                                                                                      \
    } while (0)
#else
#define AT10_EV(ROLE, ID, IDX) do {} while (0)
#endif

// -DAT10_PROF: CTA 0 accumulates the cycles its roles spend in each class of barrier wait (two clock reads per wait) and
// writes them to p.trace: [0] producer kv_empty, [1] producer q_empty, [2] MMA0 kv_full, [3] MMA0 s_free, [4] MMA0 p_full,
// [5] MMA0 q_full, [6] WG0 s_full, [7] WG0 o_full (epilogue), [8] WG0 total, [9] MMA0 total, [10] WG0 epilogue total.
#ifdef AT10_PROF
#define AT10_PWAIT(SLOT, STMT)                                        \
    do {                                                              \
        const long long t0__ = clock64();                             \
        STMT;                                                         \
        prof_acc[SLOT] += clock64() - t0__;                           \
    } while (0)
#else
#define AT10_PWAIT(SLOT, STMT) do { STMT; } while (0)
#endif

struct Attn10Params {
    int n_tok;
    int hidden;
    int n_heads;
    int n_qblk;        // ceil(n_tok / 256)
    int num_items;     // batch * n_heads * n_qblk
    __half *out;       // (unused by the kernel: the output is written through tmOut)
    float scale_log2;  // log2(e) / sqrt(64)
    // -fa compatibility (reference dinov2.cpp:499-525): the flash path zero-pads Q, K, V to a multiple of 32 tokens and calls
    // ggml_flash_attn_ext WITHOUT a mask, so every query also attends to n_phantom all-zero keys (score 0, value 0): they add
    // n_phantom * exp(0 - max) to the softmax denominator and nothing to the numerator.  0 = exact attention (default path).
    int n_phantom;
    int reverse;       // walk each CTA's work items from last to first (L2 reuse across consecutive kernels, see GemmParams::reverse)
    unsigned long long *trace;   // AT10_TRACE builds only
};

// 3-D tiled store shared -> global; parts of the box outside the tensor are clipped
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *m, const void *smem_src, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// Work item -> (image, head, query block), walked incrementally: consecutive items share K/V (same image and head).
struct Attn10Item {
    int qb, head, img;
    __device__ __forceinline__ void init(int item, int n_qblk, int n_heads) {
        qb = item % n_qblk;
        const int ih = item / n_qblk;
        head = ih % n_heads;
        img = ih / n_heads;
    }
    __device__ __forceinline__ void next(int n_qblk, int n_heads) {
        if (++qb == n_qblk) {
            qb = 0;
            if (++head == n_heads) {
                head = 0;
                ++img;
            }
        }
    }
    __device__ __forceinline__ void prev(int n_qblk, int n_heads) {
        if (--qb < 0) {
            qb = n_qblk - 1;
            if (--head < 0) {
                head = n_heads - 1;
                --img;
            }
        }
    }
    // step in the walking direction of the kernel
    __device__ __forceinline__ void step(int n_qblk, int n_heads, int reverse) {
        if (reverse) prev(n_qblk, n_heads);
        else next(n_qblk, n_heads);
    }
    __device__ __forceinline__ bool has_q1(int n_tok) const { return qb * 256 + 128 < n_tok; }
};

__global__ void __launch_bounds__(AT10_THREADS, 1)
attention_fwd_v10(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmOut, const Attn10Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;                                   // [buffer][tile]
    uint8_t *sK = sQ + 4 * AT10_TILE;                     // [stages]
    uint8_t *sV = sK + AT10_KV_STAGES * AT10_TILE;        // [stages]
    uint8_t *sO = sV + AT10_KV_STAGES * AT10_TILE;        // [tile]: output staging
    uint64_t *bars = reinterpret_cast<uint64_t *>(sO + 2 * AT10_TILE);
    uint64_t *q_full = bars;                              // 2
    uint64_t *q_empty = bars + 2;                         // 2
    uint64_t *kv_full = bars + 4;                         // stages
    uint64_t *kv_empty = kv_full + AT10_KV_STAGES;        // stages
    uint64_t *s_full = kv_empty + AT10_KV_STAGES;         // 2: S_t(n) is in TMEM
    uint64_t *s_free = s_full + 2;                        // 2: S_t(n) is in registers (its second slot may be overwritten)
    // P_t(n) is in TMEM.  Two barriers per tile, used alternately: a warpgroup may finish P_t(n+1) before the MMA warp (held
    // up by the other tile) has looked at P_t(n) — with a single barrier that is two phase flips and the parity wait never
    // returns.  It cannot be two tiles ahead: S_t(n+2) is only issued after the MMA warp has consumed P_t(n).
    uint64_t *p_full = s_free + 2;                        // 2 x 2
    uint64_t *o_full = p_full + 4;                        // 2: P_t(n) V has completed
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(o_full + 2);
    [[maybe_unused]] float *xchg = reinterpret_cast<float *>(sO + 2 * AT10_TILE + 256);   // (AT10_SPLIT) [kind][query tile][half][row]
    // (AT10_SPLIT) the two warps that share 32 query rows (same query tile and lane quarter, key halves 0 / 1) meet here once per tile
    [[maybe_unused]] uint64_t *pair_bar = reinterpret_cast<uint64_t *>(tmem_ptr + 2);     // [query tile][lane quarter], 64 arrivals
    // (AT10_OBSERVE_EVERY_PV == 2) number of completed P V products per query tile, published by the tile's MMA warp
    [[maybe_unused]] uint32_t *pv_done = reinterpret_cast<uint32_t *>(pair_bar + 8);      // bytes 248..255 of the barrier block

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_kv = (p.n_tok + 127) / 128;
    // contiguous, balanced item range of this CTA: consecutive items share K/V (same image and head), so a CTA re-reads
    // them from L2, and every CTA gets the same mix of full and half (single query tile) blocks
    const int item_lo = static_cast<int>(static_cast<long long>(p.num_items) * blockIdx.x / gridDim.x);
    const int item_hi = static_cast<int>(static_cast<long long>(p.num_items) * (blockIdx.x + 1) / gridDim.x);

#ifdef AT10_PROF
    long long prof_acc[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_t0 = clock64();
#endif
    if (warp == AT10_HW + 1 && lane == 0) {
        prefetch_tmap(&tmQKV);
        prefetch_tmap(&tmOut);
    }
    if (warp == AT10_HW + 3 && lane == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(&q_full[b], 1);
            mbar_init(&q_empty[b], 2);                    // both MMA warps have issued their last Q K^T of the item
        }
        for (int s = 0; s < AT10_KV_STAGES; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 2);                   // both MMA warps' P V products that read the stage have completed
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], AT10_SM_THREADS);
            mbar_init(&p_full[2 * t], AT10_SM_THREADS);
            mbar_init(&p_full[2 * t + 1], AT10_SM_THREADS);
            mbar_init(&o_full[t], 1);
        }
#if AT10_SPLIT
        for (int i = 0; i < 8; ++i) mbar_init(&pair_bar[i], 64);
#endif
        pv_done[0] = 0;
        pv_done[1] = 0;
        fence_mbar_init();
    }
    if (warp == AT10_HW) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_R = tmem_base;           // ring of tile t: columns 192 t + 64 slot
    const uint32_t tmem_O = tmem_base + 384;     // O_t at columns 384 + 64 t
    griddep_wait();                              // prologue above overlaps the previous kernel's tail (programmatic dependent launch)
    griddep_launch_dependents();

    if (warp >= AT10_HW) {
        setmaxnreg_dec<AT10_SPLIT ? 64 : 80>();
        if (warp == AT10_HW + 1 && item_lo < item_hi) {
            // ---------------------------------------------------------------- TMA producer (warp-uniform; one lane issues)
            int s = 0;
            uint32_t ph = 0, li = 0;                       // K/V stage + phase; local item index (Q buffer = li & 1)
            Attn10Item it;
            it.init(p.reverse ? item_hi - 1 : item_lo, p.n_qblk, p.n_heads);
            for (int item = item_lo; item < item_hi; ++item, ++li, it.step(p.n_qblk, p.n_heads, p.reverse)) {
                const int row_base = it.img * p.n_tok, q_base = it.qb * 256;
                const bool has_q1 = it.has_q1(p.n_tok);
                const uint32_t qb = li & 1;
                AT10_PWAIT(1, mbar_wait(&q_empty[qb], ((li >> 1) & 1) ^ 1));   // every Q K^T of the item that used this buffer last has completed
                if (elect_one()) {
                    mbar_arrive_expect_tx(&q_full[qb], (has_q1 ? 2 : 1) * AT10_TILE);
                    tma_load_2d(sQ + qb * 2 * AT10_TILE, &tmQKV, &q_full[qb], it.head * 64, row_base + q_base);
                    if (has_q1) tma_load_2d(sQ + (qb * 2 + 1) * AT10_TILE, &tmQKV, &q_full[qb], it.head * 64, row_base + q_base + 128);
                }
                __syncwarp();
                for (int j = 0; j < n_kv; ++j) {
                    AT10_PWAIT(0, mbar_wait(&kv_empty[s], ph ^ 1));
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&kv_full[s], 2 * AT10_TILE);
                        tma_load_2d(sK + s * AT10_TILE, &tmQKV, &kv_full[s], p.hidden + it.head * 64, row_base + j * 128);
                        tma_load_2d(sV + s * AT10_TILE, &tmQKV, &kv_full[s], 2 * p.hidden + it.head * 64, row_base + j * 128);
                    }
                    __syncwarp();
                    if (++s == AT10_KV_STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp >= AT10_HW + 2 && item_lo < item_hi) {
            // ---------------------------------------------------------------- MMA issuers: warp 11 -> query tile 0, warp 10 -> tile 1
            // All 32 lanes run the control flow, barrier waits and descriptor arithmetic (warp-uniform -> uniform
            // datapath); one elected lane issues tcgen05.mma / tcgen05.commit.  Both warps walk every item and tile; the
            // tile-1 warp skips the MMAs of items without a second query tile but still takes part in the stage / Q-buffer
            // hand-backs (a commit with nothing outstanding arrives at once).
            const int T = AT10_HW + 3 - warp;
            constexpr uint32_t idesc_s128 = make_idesc_f16(128, 128, 0, 0);
            constexpr uint32_t idesc_s64 = make_idesc_f16(128, 64, 0, 0);
            constexpr uint32_t idesc_o = make_idesc_f16(128, 64, 0, 1);     // A = P (TMEM), B = V, MN-major
            int s = 0;                                     // K/V stage of the tile whose P V comes next
            uint32_t ph = 0, li = 0;
            uint32_t n_s = 0, n_p = 0;                     // S / P tiles issued so far: barrier phase = count & 1
            uint32_t slot_s = 0, slot_p = 0;               // ring slots (count % 3) of the next S / next P
            [[maybe_unused]] int tr_n = 0;
            const uint32_t ring = tmem_R + T * 192;
            const uint32_t o_acc = tmem_O + T * 64;
            const uint64_t q_desc_base = make_smem_desc_sw128(smem_u32(sQ + T * AT10_TILE), 16, 1024);
            const uint64_t k_desc_base = make_smem_desc_sw128(smem_u32(sK), 16, 1024);
            // MN-major B, N = 64: a single 64-wide atom along MN (leading-dim offset unused)
            const uint64_t v_desc_base = make_smem_desc_sw128(smem_u32(sV), 1024, 1024);
// S(n) = Q K(stage)^T into ring slots (slot, slot+1 mod 3): one N=128 MMA per k-step when the slots are adjacent,
// two N=64 MMAs (keys 0-63 -> slot 2, keys 64-127 -> slot 0; K rows 64.. start 8 KB into the tile) when the ring wraps
#define AT10_ISSUE_S(QDESC, STAGE)                                                                                     \
    do {                                                                                                               \
        if (n_s > 0) {                                                                                                 \
            AT10_PWAIT(3, mbar_wait(&s_free[T], (n_s - 1) & 1));                                                       \
            tc_fence_after();                                                                                          \
        }                                                                                                              \
        const uint64_t k_desc__ = k_desc_base + static_cast<uint64_t>((STAGE) * (AT10_TILE >> 4));                     \
        if (elect_one()) {                                                                                             \
            if (slot_s != 2) {                                                                                         \
                _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                          \
                    umma_f16_ss(ring + slot_s * 64, (QDESC) + 2 * k, k_desc__ + 2 * k, idesc_s128, k != 0);            \
            } else {                                                                                                   \
                _Pragma("unroll") for (int k = 0; k < 4; ++k) {                                                        \
                    umma_f16_ss(ring + 128, (QDESC) + 2 * k, k_desc__ + 2 * k, idesc_s64, k != 0);                     \
                    umma_f16_ss(ring, (QDESC) + 2 * k, k_desc__ + (8192 >> 4) + 2 * k, idesc_s64, k != 0);             \
                }                                                                                                      \
            }                                                                                                          \
            umma_commit(&s_full[T]);                                                                                   \
            if (T == 0) AT10_EV(0, 1 + T, n_s);                                                                                    \
        }                                                                                                              \
        __syncwarp();                                                                                                  \
        n_s++;                                                                                                         \
        slot_s = slot_s == 2 ? 0u : slot_s + 1;                                                                        \
    } while (0)
// (AT10_OBSERVE_EVERY_PV == 2) This warp waits for every P V it has issued — one tile later, when the product has long
// completed — and publishes the count.  It sees every phase of o_full in order, so its parity waits are unambiguous; the softmax
// warps, which need "P(n-1) V(n-1) has completed" only in the rare growth path and once per item in the epilogue, poll the count.
#define AT10_PUBLISH_PV()                                                                                              \
    do {                                                                                                               \
        if (n_pub < n_p) {                                                                                             \
            mbar_wait(&o_full[T], (n_p - 1) & 1);                                                                      \
            tc_fence_after();                                                                                          \
            if (lane == 0) st_release_cta_shared(&pv_done[T], n_p);                                                    \
            __syncwarp();                                                                                              \
            n_pub = n_p;                                                                                               \
        }                                                                                                              \
    } while (0)
            [[maybe_unused]] uint32_t n_pub = 0;
            Attn10Item it;
            it.init(p.reverse ? item_hi - 1 : item_lo, p.n_qblk, p.n_heads);
            bool act = T == 0 || it.has_q1(p.n_tok);       // this warp's query tile exists in the current item
            // the very first S of this CTA
            mbar_wait(&q_full[0], 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            if (act) AT10_ISSUE_S(q_desc_base, 0);
            if (n_kv == 1 && elect_one()) umma_commit(&q_empty[0]);
            __syncwarp();
            for (int item = item_lo; item < item_hi; ++item, ++li) {
                const bool has_next = item + 1 < item_hi;
                it.step(p.n_qblk, p.n_heads, p.reverse);
                const bool act_next = has_next && (T == 0 || it.has_q1(p.n_tok));
                const uint64_t q_cur = q_desc_base + static_cast<uint64_t>((li & 1) * ((2 * AT10_TILE) >> 4));
                const uint64_t q_nxt = q_desc_base + static_cast<uint64_t>(((li + 1) & 1) * ((2 * AT10_TILE) >> 4));
                for (int j = 0; j < n_kv; ++j) {
#if AT10_OBSERVE_EVERY_PV == 2
                    AT10_PUBLISH_PV();
#endif
                    int s1 = s + 1;
                    uint32_t ph1 = ph;
                    if (s1 == AT10_KV_STAGES) { s1 = 0; ph1 ^= 1; }
                    if (j + 1 < n_kv) {
                        // S(j+1) of this item: computed while S(j) is being exponentiated
                        AT10_PWAIT(2, mbar_wait(&kv_full[s1], ph1));
                        tc_fence_after();
                        if (T == 0) AT10_EV(0, 5, j);
                        if (act) AT10_ISSUE_S(q_cur, s1);
                        if (j + 2 == n_kv && elect_one()) umma_commit(&q_empty[li & 1]);   // last Q K^T of this item is in flight
                        __syncwarp();
                    } else if (has_next) {
                        // first S of the NEXT item (other Q buffer): the tile stream does not stop at the item boundary
                        AT10_PWAIT(5, mbar_wait(&q_full[(li + 1) & 1], ((li + 1) >> 1) & 1));
                        AT10_PWAIT(2, mbar_wait(&kv_full[s1], ph1));
                        tc_fence_after();
                        if (T == 0) AT10_EV(0, 6, j);
                        if (act_next) AT10_ISSUE_S(q_nxt, s1);
                        if (n_kv == 1 && elect_one()) umma_commit(&q_empty[(li + 1) & 1]);
                        __syncwarp();
                    }
                    // O (+)= P(n) V  (8 k-steps of 16 keys; P = 8 TMEM columns per step in ring slot n % 3), then o_full
                    if (act) {
                        AT10_PWAIT(4, mbar_wait(&p_full[2 * T + (n_p & 1)], (n_p >> 1) & 1));
                        tc_fence_after();
                        if (T == 0) AT10_EV(0, 7, n_p);
                    }
                    const uint64_t v_desc = v_desc_base + static_cast<uint64_t>(s * (AT10_TILE >> 4));
                    if (elect_one()) {
                        if (act) {
#pragma unroll
                            for (int k = 0; k < 8; ++k)
                                umma_f16_ts(o_acc, ring + slot_p * 64 + 8 * k, v_desc + static_cast<uint64_t>(k * (2048 >> 4)), idesc_o, (j | k) != 0);
                            umma_commit(&o_full[T]);
                            if (T == 0) AT10_EV(0, 3 + T, n_p);
                        }
                        umma_commit(&kv_empty[s]);         // this warp's share of the stage hand-back
                    }
                    __syncwarp();
                    if (act) {
                        n_p++;
                        slot_p = slot_p == 2 ? 0u : slot_p + 1;
                    }
                    s = s1;
                    ph = ph1;
                }
                act = act_next;
            }
#if AT10_OBSERVE_EVERY_PV == 2
            AT10_PUBLISH_PV();
#endif
        }
    } else {
#if AT10_SPLIT
        // ---------------------------------------------------------------- softmax, two threads per query row (see AT10_SPLIT above)
        setmaxnreg_inc<104>();
        const int t = (warp >> 2) & 1;                    // query tile
        const int h = warp >> 3;                          // key half of every 128-key tile (and output-column half of the row)
        const int qd = warp & 3;                          // TMEM lane quarter
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ring = tmem_R + lane_addr + t * 192;
        const uint32_t o_addr = tmem_O + lane_addr + t * 64 + 32 * h;        // this thread's 32 of the row's 64 output columns
        const float c = p.scale_log2;
        const float thr = AT10_RESCALE_LOG2 / c;
        uint32_t n_tile = 0;
        uint32_t slot = 0;
        [[maybe_unused]] int tr_n = 0;
        uint8_t *stage_row = sO + t * AT10_TILE + r * 128;
        const uint32_t sw = static_cast<uint32_t>(r & 7);
        const bool tile_leader = h == 0 && (threadIdx.x & 127) == 0;         // issues the TMA store of query tile t
        float *xmax_own = xchg + ((0 * 2 + t) * 2 + h) * 128 + r, *xmax_other = xchg + ((0 * 2 + t) * 2 + (1 - h)) * 128 + r;
        float *xsum_own = xchg + ((1 * 2 + t) * 2 + h) * 128 + r, *xsum_other = xchg + ((1 * 2 + t) * 2 + (1 - h)) * 128 + r;
#define AT10_SEV(ID) do { if (tile_leader) AT10_EV(1 + t, ID, n_tile); } while (0)
        bool pending = false;
        float pend_l = 1.f;
        int pend_c0 = 0, pend_c1 = 0, pend_c2 = 0;
// This half's 32 columns of O_t / (row sum of both halves) -> fp16 -> its 64 bytes of the row in the swizzled staging tile;
// one TMA store per query tile.  The two barriers also order the exchange of the halves' row sums.
#define AT10_EPILOGUE()                                                                                                \
    do {                                                                                                               \
        AT10_PWAIT(7, mbar_wait(&o_full[t], (n_tile - 1) & 1));                                                        \
        tc_fence_after();                                                                                              \
        uint32_t a__[32];                                                                                              \
        tmem_ld_32x32b_x32(o_addr, a__);                                                                               \
        tmem_ld_wait();                                                                                                \
        tc_fence_before();                                                                                             \
        *xsum_own = pend_l;                                                                                            \
        if (tile_leader) bulk_wait_read<0>();    /* the previous store has finished reading the staging tile */        \
        named_bar_sync(1 + t, 256);                                                                                    \
        const float inv__ = 1.0f / (pend_l + *xsum_other);                                                             \
        _Pragma("unroll") for (int v = 0; v < 4; ++v) {                                                                \
            *reinterpret_cast<uint4 *>(stage_row + ((static_cast<uint32_t>(v + 4 * h) ^ sw) << 4)) =                   \
                make_uint4(pack_half2(__uint_as_float(a__[8 * v]) * inv__, __uint_as_float(a__[8 * v + 1]) * inv__),   \
                           pack_half2(__uint_as_float(a__[8 * v + 2]) * inv__, __uint_as_float(a__[8 * v + 3]) * inv__), \
                           pack_half2(__uint_as_float(a__[8 * v + 4]) * inv__, __uint_as_float(a__[8 * v + 5]) * inv__), \
                           pack_half2(__uint_as_float(a__[8 * v + 6]) * inv__, __uint_as_float(a__[8 * v + 7]) * inv__)); \
        }                                                                                                              \
        fence_proxy_async_smem();                                                                                      \
        named_bar_sync(1 + t, 256);                                                                                    \
        if (tile_leader) {                                                                                             \
            tma_store_3d(&tmOut, sO + t * AT10_TILE, pend_c0, pend_c1, pend_c2);                                       \
            bulk_commit();                                                                                             \
        }                                                                                                              \
        pending = false;                                                                                               \
    } while (0)

        Attn10Item it;
        it.init(p.reverse ? item_hi - 1 : item_lo, p.n_qblk, p.n_heads);
        for (int item = item_lo; item < item_hi; ++item, it.step(p.n_qblk, p.n_heads, p.reverse)) {
            if (t == 1 && !it.has_q1(p.n_tok)) continue;
            float m_used = -INFINITY;
            float l_run = 0.f;                            // this half's part of the softmax denominator, relative to m_used
            [[maybe_unused]] float l_poly = 0.f;          // (AT10_INTPACK) its polynomial-pair part; l_run is then in the 2^-112 scale
#if AT10_INTPACK
#define AT10_EXP(E0, E1, CH, MC) attn_exp_pairs_ip<E0, E1>(CH, pk, c, MC, ls, lp)
#else
#define AT10_EXP(E0, E1, CH, MC) attn_exp_pairs<E0, E1>(CH, pk, c, MC, ls)
#endif
#if AT10_PINGPONG
            // baton between the two query tiles for the exponential segments (see the one-thread-per-row path below): all four
            // warpgroups take part, the two of a query tile take and pass together
            const bool pp = it.has_q1(p.n_tok);
            const int pp_segs = n_kv + 1;
            int pp_seg = 0;
            // one baton PER SCHEDULER (lane quarter qd = warp & 3 = the SM sub-partition the warp lives on): the XU pipe is shared per
            // sub-partition, and a baton over all warps of a query tile waits for the slowest of its eight warps (hand-over ~285
            // cycles measured).  Barriers 7 + qd: "tile 1 is done here, tile 0 may go", 11 + qd: the reverse; 128 threads each.
            if (pp && t == 1) named_bar_arrive(7 + qd, 128);
            auto baton_take = [&]() {
                if (pp) named_bar_sync((t == 0 ? 7 : 11) + qd, 128);
            };
            auto baton_pass = [&]() {
                ++pp_seg;
                if (pp && !(t == 1 && pp_seg == pp_segs)) named_bar_arrive((t == 0 ? 11 : 7) + qd, 128);
            };
#else
            auto baton_take = [&]() {};
            auto baton_pass = [&]() {};
#endif
            auto tile = [&](auto first_c, auto last_c) {
                constexpr bool FIRST = decltype(first_c)::value, LAST = decltype(last_c)::value;
                const uint32_t lo = ring + slot * 64;                          // keys 0-63 (P of the whole tile goes back here)
                const uint32_t hi = ring + (slot == 2 ? 0u : slot + 1) * 64;   // keys 64-127
                slot = slot == 2 ? 0u : slot + 1;
                const uint32_t src = h ? hi : lo;                              // this half's 64 scores
                const uint32_t pdst = lo + 32 * h;                             // this half's 32 columns of packed P
                AT10_SEV(10);
                AT10_PWAIT(6, mbar_wait(&s_full[t], n_tile & 1));
                tc_fence_after();
                AT10_SEV(11);
                uint32_t c0[32], c1[32];
                tmem_ld_32x32b_x32(src, c0);
                tmem_ld_32x32b_x32(src + 32, c1);
                tmem_ld_wait();
                AT10_SEV(12);
                // keys of this half that exist (only an item's last tile is ragged; <= 0: the whole half lies beyond the image)
                const int kv_valid = LAST ? p.n_tok - (n_kv - 1) * 128 - 64 * h : 64;
                if constexpr (LAST) {
                    if (kv_valid < 32) attn_mask32(c0, kv_valid);
                    if (kv_valid < 64) attn_mask32(c1, kv_valid - 32);
                }
                // the row's maximum over all 128 keys: exchanged with the thread that holds the other half (the two warps of a lane
                // quarter meet on an mbarrier: a named barrier over the query tile's eight warps waits for the slowest).  s_free is announced
                // only after the partner's value has been read, so its next write (after S(n+1), which waits for s_free) cannot
                // overtake this read; the barrier also keeps the P columns written below from clobbering scores the partner
                // has not loaded yet (P of keys 64-127 lands on the columns of the scores of keys 32-63).
                const float mxh = fmaxf(attn_rowmax32(c0), attn_rowmax32(c1));
                *xmax_own = mxh;
                mbar_arrive(&pair_bar[t * 4 + qd]);
                mbar_wait(&pair_bar[t * 4 + qd], n_tile & 1);
                const float mx = fmaxf(mxh, *xmax_other);
                tc_fence_before();
                mbar_arrive(&s_free[t]);
                if constexpr (FIRST) {
                    m_used = p.n_phantom > 0 ? fmaxf(mx, 0.f) : mx;
                    l_run = 0.f;
                    l_poly = 0.f;
                } else {
                    const bool grow = mx > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {  // rare: O_t (this half's columns) and the sum move down to the new reference
                        mbar_wait(&o_full[t], (n_tile - 1) & 1);      // O_t must be quiescent, i.e. P(n-1) V(n-1) complete
                        tc_fence_after();
                        const float alpha = grow ? ex2_approx((m_used - mx) * c) : 1.0f;
                        if (grow) m_used = mx;
                        l_run *= alpha;
                        l_poly *= alpha;
#pragma unroll 1
                        for (int i = 0; i < 32; i += 8) {
                            uint32_t a[8];
                            tmem_ld_32x32b_x8(o_addr + i, a);
                            tmem_ld_wait();
#pragma unroll
                            for (int d = 0; d < 8; ++d) a[d] = __float_as_uint(__uint_as_float(a[d]) * alpha);
                            tmem_st_32x32b_x8(o_addr + i, a);
                        }
                        tmem_st_wait();
                    }
                }
                baton_take();
                AT10_SEV(14);
                float ls[2] = {0.f, 0.f};
                [[maybe_unused]] float lp[2] = {0.f, 0.f};
                uint32_t pk[16];
                const float mc = m_used * c;
                if (!LAST || kv_valid > 0) {
                    AT10_EXP(0, 16, c0, mc);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;
                }
                tmem_st_32x32b_x16(pdst, pk);
                // the previous item's output: O_t stays untouched until this tile's P V, issued only after the p_full arrive below
                if constexpr (FIRST) {
                    baton_pass();
                    if (pending) {
                        AT10_SEV(18);
                        AT10_PWAIT(10, AT10_EPILOGUE());
                        AT10_SEV(19);
                    }
                    baton_take();
                }
                if (!LAST || kv_valid > 32) {
                    AT10_EXP(0, 16, c1, mc);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) pk[i] = 0u;
                }
                tmem_st_32x32b_x16(pdst + 16, pk);
                baton_pass();
                l_run += ls[0] + ls[1];
                l_poly += lp[0] + lp[1];
                AT10_SEV(15);
#if AT10_OBSERVE_EVERY_PV
                if constexpr (!FIRST) mbar_wait(&o_full[t], (n_tile - 1) & 1);      // see the one-thread-per-row path
#endif
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[2 * t + (n_tile & 1)]);
                AT10_SEV(17);
                ++n_tile;
            };
            using yes_t = std::integral_constant<bool, true>;
            using no_t = std::integral_constant<bool, false>;
            if (n_kv == 1) {
                tile(yes_t{}, yes_t{});
            } else {
                tile(yes_t{}, no_t{});
#pragma unroll 1
                for (int j = 1; j + 1 < n_kv; ++j) tile(no_t{}, no_t{});
                tile(no_t{}, yes_t{});
            }
            pending = true;
#if AT10_INTPACK
            if (p.n_phantom > 0 && h == 0) l_poly += static_cast<float>(p.n_phantom) * ex2_approx(ATS_IP_SHIFT - m_used * c);   // the zero keys of the -fa path
            pend_l = fmaf(l_run, ATS_IP_UNBIAS, l_poly);
#else
            if (p.n_phantom > 0 && h == 0) l_run += static_cast<float>(p.n_phantom) * ex2_approx(-m_used * c);   // the zero keys of the -fa path
            pend_l = l_run;
#endif
            pend_c0 = it.head * 64;
            pend_c1 = it.qb * 256 + t * 128;
            pend_c2 = it.img;
        }
        if (pending) AT10_EPILOGUE();
        if (tile_leader) bulk_wait<0>();                  // the staging tile must outlive the last TMA store
#else
        setmaxnreg_inc<208>();
        const int t = warp >> 2;                          // query tile / warpgroup
        const int qd = warp & 3;                          // TMEM lane quarter
        const int r = qd * 32 + lane;                     // row inside the tile
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ring = tmem_R + lane_addr + t * 192;          // this row's ring of three 64-column slots
        const uint32_t o_addr = tmem_O + lane_addr + t * 64;
        const float c = p.scale_log2;
        const float thr = AT10_RESCALE_LOG2 / c;          // threshold in raw-score units
        uint32_t n_tile = 0;                              // tiles processed by this warpgroup (barrier phases)
        uint32_t slot = 0;                                // n_tile % 3
        [[maybe_unused]] int tr_n = 0;
        uint8_t *stage_row = sO + t * AT10_TILE + r * 128;           // this row of the warpgroup's output staging tile
        const uint32_t sw = static_cast<uint32_t>(r & 7);            // 128-B swizzle: chunk c of row r lives at c ^ (r & 7)
        const bool wg_leader = (threadIdx.x & 127) == 0;
#define AT10_SEV(ID) do { if ((threadIdx.x & 127) == 0) AT10_EV(1 + t, ID, n_tile); } while (0)
// "the first N products P V of this query tile have completed"
#if AT10_OBSERVE_EVERY_PV == 2
#define AT10_WAIT_PV(N) do { while (ld_acquire_cta_shared(&pv_done[t]) < (N)) {} } while (0)
#else
#define AT10_WAIT_PV(N) mbar_wait(&o_full[t], ((N) - 1) & 1)
#endif

        // Deferred epilogue state: the finished item whose O_t is still in tensor memory
        bool pending = false;
        float pend_l = 1.f;
        int pend_c0 = 0, pend_c1 = 0, pend_c2 = 0;        // TMA store coordinates: channel, token, image
// O_t / rowsum -> fp16 rows -> swizzled staging tile -> one TMA store.  Runs with n_tile = index of the tile AFTER the item's last.
#define AT10_EPILOGUE()                                                                                                \
    do {                                                                                                               \
        AT10_PWAIT(7, AT10_WAIT_PV(n_tile));                                                                           \
        tc_fence_after();                                                                                              \
        uint32_t a__[32], b__[32];                                                                                     \
        tmem_ld_32x32b_x32(o_addr, a__);                                                                               \
        tmem_ld_32x32b_x32(o_addr + 32, b__);                                                                          \
        tmem_ld_wait();                                                                                                \
        tc_fence_before();                                                                                             \
        if (wg_leader) bulk_wait_read<0>();      /* the previous store has finished reading the staging tile */        \
        named_bar_sync(1 + t, 128);                                                                                    \
        const float inv__ = 1.0f / pend_l;                                                                             \
        _Pragma("unroll") for (int v = 0; v < 4; ++v) {                                                                \
            *reinterpret_cast<uint4 *>(stage_row + ((static_cast<uint32_t>(v) ^ sw) << 4)) =                           \
                make_uint4(pack_half2(__uint_as_float(a__[8 * v]) * inv__, __uint_as_float(a__[8 * v + 1]) * inv__),   \
                           pack_half2(__uint_as_float(a__[8 * v + 2]) * inv__, __uint_as_float(a__[8 * v + 3]) * inv__), \
                           pack_half2(__uint_as_float(a__[8 * v + 4]) * inv__, __uint_as_float(a__[8 * v + 5]) * inv__), \
                           pack_half2(__uint_as_float(a__[8 * v + 6]) * inv__, __uint_as_float(a__[8 * v + 7]) * inv__)); \
            *reinterpret_cast<uint4 *>(stage_row + ((static_cast<uint32_t>(v + 4) ^ sw) << 4)) =                       \
                make_uint4(pack_half2(__uint_as_float(b__[8 * v]) * inv__, __uint_as_float(b__[8 * v + 1]) * inv__),   \
                           pack_half2(__uint_as_float(b__[8 * v + 2]) * inv__, __uint_as_float(b__[8 * v + 3]) * inv__), \
                           pack_half2(__uint_as_float(b__[8 * v + 4]) * inv__, __uint_as_float(b__[8 * v + 5]) * inv__), \
                           pack_half2(__uint_as_float(b__[8 * v + 6]) * inv__, __uint_as_float(b__[8 * v + 7]) * inv__)); \
        }                                                                                                              \
        fence_proxy_async_smem();                                                                                      \
        named_bar_sync(1 + t, 128);                                                                                    \
        if (wg_leader) {                                                                                               \
            tma_store_3d(&tmOut, sO + t * AT10_TILE, pend_c0, pend_c1, pend_c2);                                       \
            bulk_commit();                                                                                             \
        }                                                                                                              \
        pending = false;                                                                                               \
    } while (0)

        Attn10Item it;
        it.init(p.reverse ? item_hi - 1 : item_lo, p.n_qblk, p.n_heads);
        for (int item = item_lo; item < item_hi; ++item, it.step(p.n_qblk, p.n_heads, p.reverse)) {
            if (t == 1 && !it.has_q1(p.n_tok)) continue;
            float m_used = -INFINITY;
            float l_run = 0.f;                            // softmax denominator relative to m_used
            float l_poly = 0.f;                           // (AT10_INTPACK) its polynomial-pair part; l_run is then in the 2^-112 scale
#if AT10_PINGPONG
            // Baton between the warpgroups for the exponential segments of this item (only when both query tiles exist): barrier
            // 3 = "warpgroup 1 is done, 0 may go", barrier 4 = the reverse.  Every tile is one segment, an item's first tile two
            // (the previous item's epilogue sits between them, outside the baton).  Warpgroup 1 hands over first and keeps the
            // baton it would pass after its last segment, so both barriers are balanced per item.
            const bool pp = it.has_q1(p.n_tok);
            const int pp_segs = n_kv + 1;
            int pp_seg = 0;
            if (pp && t == 1) named_bar_arrive(3, 256);
            auto baton_take = [&]() {
                if (pp) named_bar_sync(3 + t, 256);
            };
            auto baton_pass = [&]() {
                ++pp_seg;
                if (pp && !(t == 1 && pp_seg == pp_segs)) named_bar_arrive(4 - t, 256);
            };
#else
            auto baton_take = [&]() {};
            auto baton_pass = [&]() {};
#endif

            // One 128-key tile.  The body is instantiated per position in the item — FIRST (sets the reference maximum, hosts the
            // previous item's deferred epilogue), LAST (ragged: keys past the image's last token are masked), middle (neither) —
            // so that the nine middle tiles of an 11-tile item carry no mask / first-tile / epilogue branches: every branch cuts
            // the unrolled body into separate scheduling regions for ptxas and drains the MUFU pipe (measured with three extra
            // branches: 724 vs 675 us per ViT-L layer).
#if AT10_INTPACK
#define AT10_EXP(E0, E1, CH, MC) attn_exp_pairs_ip<E0, E1>(CH, pk, c, MC, ls, lp)
#else
#define AT10_EXP(E0, E1, CH, MC) attn_exp_pairs<E0, E1>(CH, pk, c, MC, ls)
#endif
            auto tile = [&](auto first_c, auto last_c) {
                constexpr bool FIRST = decltype(first_c)::value, LAST = decltype(last_c)::value;
                const uint32_t lo = ring + slot * 64;                          // keys 0-63 (P goes back here)
                const uint32_t hi = ring + (slot == 2 ? 0u : slot + 1) * 64;   // keys 64-127
                slot = slot == 2 ? 0u : slot + 1;
                AT10_SEV(10);
                AT10_PWAIT(6, mbar_wait(&s_full[t], n_tile & 1));
#if AT10_STAGGER > 0
                if (t == 1 && n_tile == 0) {                                   // experiment: hold the second warpgroup back once, after its first S has arrived
                    const long long t0 = clock64();
                    while (clock64() - t0 < AT10_STAGGER) {}
                }
#endif
                tc_fence_after();
                AT10_SEV(11);
                uint32_t c0[32], c1[32], c2[32], c3[32];
                tmem_ld_32x32b_x32(lo, c0);
                tmem_ld_wait();
                tmem_ld_32x32b_x32(lo + 32, c1);          // in flight under chunk 0's maximum and first exponentials
                tmem_ld_32x32b_x32(hi, c2);
                tmem_ld_32x32b_x32(hi + 32, c3);
                AT10_SEV(12);
                const int kv_valid = LAST ? p.n_tok - (n_kv - 1) * 128 : 128;   // keys of this tile that exist (only the last tile is ragged)
                if constexpr (LAST) {
                    if (kv_valid < 32) attn_mask32(c0, kv_valid);
                }
                // Reference maximum: moves only when a row grew by more than 2^8 over it (then O_t and the running sums are rescaled)
                // — probabilities stay <= 256, exact in fp16.  ONE growth test per tile, after the maximum of all 128 keys is known;
                // chunk 0 is exponentiated before that with the maximum carried over from the previous tiles (for an item's first
                // tile: with its own maximum).  If the tile then turns out to have grown, the sums and the 16 P columns of chunk 0
                // are rescaled — or, when the stale reference was so low that they may have overflowed fp16 (growth > 2^15),
                // chunk 0 is simply exponentiated again.
                const float mx0 = attn_rowmax32(c0);
                if constexpr (FIRST) {
                    m_used = p.n_phantom > 0 ? fmaxf(mx0, 0.f) : mx0;   // O_t is overwritten by the first P V of the item
                    l_run = 0.f;                                        // (phantom keys have score 0: the reference covers them)
                    l_poly = 0.f;
                }
                baton_take();
                AT10_SEV(14);
                float ls[2] = {0.f, 0.f};
                [[maybe_unused]] float lp[2] = {0.f, 0.f};      // (AT10_INTPACK) polynomial pairs, true scale; ls: MUFU pairs, 2^-112 scale
                uint32_t pk[16];
                {
                    const float mc = m_used * c;
                    AT10_EXP(0, 8, c0, mc);
                    // chunks 1-3 are in registers: S(n)'s second slot may be overwritten -> the MMA warp starts S(n+1)
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(&s_free[t]);
                    // keys past the image's last token (only in an item's last tile): a chunk that straddles the boundary is masked
                    // element-wise, a chunk that lies entirely beyond it is neither reduced nor exponentiated (its P is zero)
#if AT10_SKIP_MASKED
                    if constexpr (LAST) {
                        if (kv_valid > 32 && kv_valid < 64) attn_mask32(c1, kv_valid - 32);
                        if (kv_valid > 64 && kv_valid < 96) attn_mask32(c2, kv_valid - 64);
                        if (kv_valid > 96 && kv_valid < 128) attn_mask32(c3, kv_valid - 96);
                    }
                    const float mx123 = fmax3(kv_valid > 32 ? attn_rowmax32(c1) : -INFINITY, kv_valid > 64 ? attn_rowmax32(c2) : -INFINITY,
                                              kv_valid > 96 ? attn_rowmax32(c3) : -INFINITY);
#else
                    if constexpr (LAST) {
                        if (kv_valid < 128) {
                            attn_mask32(c1, kv_valid - 32);
                            attn_mask32(c2, kv_valid - 64);
                            attn_mask32(c3, kv_valid - 96);
                        }
                    }
                    const float mx123 = fmax3(attn_rowmax32(c1), attn_rowmax32(c2), attn_rowmax32(c3));
#endif
                    AT10_EXP(8, 16, c0, mc);
                    tmem_st_32x32b_x16(lo, pk);
                    const float mx = fmaxf(mx0, mx123);
                    const bool grow = mx > m_used + thr;
                    if (__any_sync(0xffffffffu, grow)) {  // rare: O_t, the sums and the 16 columns of P(n) already written move down
                        if constexpr (!FIRST) {           // O_t must be quiescent, i.e. P(n-1) V(n-1) complete
                            AT10_WAIT_PV(n_tile);
                            tc_fence_after();
                        }
#if AT10_INTPACK
                        const bool redo = grow;           // P' = 2^7 P: anything written with a reference more than 2^8 low may exceed fp16
#else
                        const bool redo = grow && (mx - m_used) * c > 15.0f;
#endif
                        const float alpha = grow ? ex2_approx((m_used - mx) * c) : 1.0f;
                        if (grow) m_used = mx;
                        l_run *= alpha;
                        l_poly *= alpha;
                        if (__any_sync(0xffffffffu, redo)) {
                            // chunk 0 again for the whole warp, relative to each lane's (possibly unchanged) reference
                            attn_rescale(o_addr, lo, alpha, !FIRST, 0);
                            ls[0] = 0.f;
                            ls[1] = 0.f;
                            lp[0] = 0.f;
                            lp[1] = 0.f;
                            AT10_EXP(0, 16, c0, m_used * c);
                            tmem_st_32x32b_x16(lo, pk);
                        } else {
                            ls[0] *= alpha;
                            ls[1] *= alpha;
                            lp[0] *= alpha;
                            lp[1] *= alpha;
                            attn_rescale(o_addr, lo, alpha, !FIRST, 16);
                        }
                    }
                }
                {
                    const float mc = m_used * c;
#if AT10_SKIP_MASKED
                    if (kv_valid > 32) {
                        AT10_EXP(0, 16, c1, mc);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    }
#else
                    AT10_EXP(0, 16, c1, mc);
#endif
                    tmem_st_32x32b_x16(lo + 16, pk);
                    // The previous item's output: O_t stays untouched until this tile's P V, which is issued only after the
                    // p_full arrive below.  64 score registers (chunks 0 and 1) are free at this point.
                    if constexpr (FIRST) {
                        baton_pass();
                        if (pending) {
                            AT10_SEV(18);
                            AT10_PWAIT(10, AT10_EPILOGUE());
                            AT10_SEV(19);
                        }
                        baton_take();
                    }
#if AT10_SKIP_MASKED
                    if (kv_valid > 64) {
                        AT10_EXP(0, 16, c2, mc);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    }
#else
                    AT10_EXP(0, 16, c2, mc);
#endif
                    tmem_st_32x32b_x16(lo + 32, pk);
#if AT10_SKIP_MASKED
                    if (kv_valid > 96) {
                        AT10_EXP(0, 16, c3, mc);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) pk[i] = 0u;
                    }
#else
                    AT10_EXP(0, 16, c3, mc);
#endif
                    tmem_st_32x32b_x16(lo + 48, pk);
                }
                baton_pass();
                l_run += ls[0] + ls[1];
                l_poly += lp[0] + lp[1];
                AT10_SEV(15);
                // (Announcing P(n) only at the top of tile n+1, under the load of its first scores, measured 1.3 % slower: the
                // P V it delays is what the MMA warp issues before the S after next.)
                // Every warp observes EVERY completion of P V before it announces its next P: P(n-1) V(n-1) was issued ~2000 cycles
                // ago, so this wait returns at once — but without it a warp that takes the growth path above only now and then
                // waits on o_full after any number of unobserved phase flips, and that parity wait was seen to return BEFORE the
                // P V it is meant for had completed (the warp then rescaled an O_t the tensor core was still adding to: 32 rows of
                // one head off by ~1e-2, different from run to run; found through the 8-GPU all-gather check of bench.py,
                // reproduced with the rescale threshold set to 0 in tools/attn_determinism.py).  With the phases observed one by
                // one the growth path's wait can only ever be for the current phase.  An item's first tile has the deferred
                // epilogue (which waits for the previous item's last P V) in this role.  (Variant 1; the default, variant 2, moves
                // the regular observation to the MMA warp: AT10_PUBLISH_PV / AT10_WAIT_PV.)
#if AT10_OBSERVE_EVERY_PV == 1
                if constexpr (!FIRST) mbar_wait(&o_full[t], (n_tile - 1) & 1);
#endif
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[2 * t + (n_tile & 1)]);
                AT10_SEV(17);
                ++n_tile;
            };
            using yes_t = std::integral_constant<bool, true>;
            using no_t = std::integral_constant<bool, false>;
            if (n_kv == 1) {
                tile(yes_t{}, yes_t{});
            } else {
                tile(yes_t{}, no_t{});
#pragma unroll 1
                for (int j = 1; j + 1 < n_kv; ++j) tile(no_t{}, no_t{});
                tile(no_t{}, yes_t{});
            }
            pending = true;
#if AT10_INTPACK
            if (p.n_phantom > 0) l_poly += static_cast<float>(p.n_phantom) * ex2_approx(ATS_IP_SHIFT - m_used * c);   // the zero keys of the -fa path
            pend_l = fmaf(l_run, ATS_IP_UNBIAS, l_poly);
#else
            if (p.n_phantom > 0) l_run += static_cast<float>(p.n_phantom) * ex2_approx(-m_used * c);   // the zero keys of the -fa path
            pend_l = l_run;
#endif
            pend_c0 = it.head * 64;
            pend_c1 = it.qb * 256 + t * 128;
            pend_c2 = it.img;
        }
        if (pending) AT10_EPILOGUE();
        if (wg_leader) bulk_wait<0>();                    // the staging tile must outlive the last TMA store
#endif
    }

#ifdef AT10_PROF
    if (blockIdx.x == 0 && p.trace && lane == 0) {
        const long long tot = clock64() - prof_t0;
        if (warp == AT10_HW + 1) { p.trace[0] = prof_acc[0]; p.trace[1] = prof_acc[1]; }
        if (warp == AT10_HW + 3) { p.trace[2] = prof_acc[2]; p.trace[3] = prof_acc[3]; p.trace[4] = prof_acc[4]; p.trace[5] = prof_acc[5]; p.trace[9] = tot; }
        if (warp == 0) { p.trace[6] = prof_acc[6]; p.trace[7] = prof_acc[7]; p.trace[8] = tot; p.trace[10] = prof_acc[10]; }
    }
#endif
    griddep_launch_dependents_late();
    tc_fence_before();
    __syncthreads();
    if (warp == AT10_HW) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

#undef AT10_SEV
#undef AT10_ISSUE_S
#undef AT10_EPILOGUE
#undef AT10_EXP
#undef AT10_PUBLISH_PV
#undef AT10_WAIT_PV

}  // namespace dino
