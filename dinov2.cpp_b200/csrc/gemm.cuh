// Persistent warp-specialised tcgen05 GEMM for every weight contraction of the DINOv2 forward pass
// (reference: ggml_mul_mat with F16 src0, ggml-cpu.c:1266-1458 — fp16 x fp16 -> f32 accumulate):
//
//     C[M, N] = A[M, K] (fp16, K contiguous)  x  W[N, K]^T (fp16, K contiguous)        + fused epilogue
//
// Roles (384 threads, 1 CTA / SM, grid = min(tiles, #SM), static round-robin tile schedule).  The single-thread
// TMA and MMA roles sit in the HIGHEST warp ids: the sub-partition arbiter prefers the highest eligible warp id, and
// an MMA issuer that has to queue behind busy epilogue warps leaves the tensor pipe idle (measured with a cycle trace:
// ~100 cycles per tcgen05.mma issue when the issuer was warp 1).
//   warp 11     MMA issuer: one lane issues 4 x tcgen05.mma (128 x BN x 16) per stage into TMEM
//   warp 10     TMA producer: 128x64 A box + BNx64 W box per stage, 128B-swizzled, mbarrier tx-count
//   warp 8      TMEM allocator (512 columns = 2 accumulator stages x BN fp32 columns)
//   warps 0-7   epilogue: tcgen05.ld 32x32b.x32 -> registers -> fused op -> 128B-swizzled smem staging (4 KB per warp)
//               -> TMA store (fp16 outputs) or TMA reduce-add (fp32 residual stream): every global access is a full
//               coalesced 128-B row segment, and the residual read-modify-write happens in the memory system
// Three pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue), tile loop.
//
// Fused epilogues (each cites the reference graph ops it absorbs):
//   EPI_BIAS_F16    out16 = acc + b                               qkv   (dinov2.cpp:471-474)
//   EPI_GELU_F16    out16 = gelu_tanh(r16(acc + b))               fc1   (dinov2.cpp:561-567, vec.h:428-457)
//   EPI_RESID_F32   X    += lambda * (acc + b)   (TMA reduce-add)  o-proj / fc2 + LayerScale + residual
//                                                                 (dinov2.cpp:546-551,708-714 / 570-573,744-749)
//   EPI_SWIGLU_F16  out16 = silu(acc_g + b_g) * (acc_u + b_u)     weights_in (dinov2.cpp:582-605); W rows are
//                                                                 pre-interleaved 128 gate | 128 up per N tile
//   EPI_PATCH_F32   X[b, 1+R+p] = acc + b + pos[1+p]              patch-embed GEMM + bias + pos-embed + token
//                                                                 placement (dinov2.cpp:636-685)
//   EPI_RESID_LN_F32  as RESID, plus the LayerNorm that follows the residual in the graph (norm2 after the attention
//                   branch, the next block's norm1 after the MLP; dinov2.cpp:722-728, 694-700), done by the two warps that
//                   are otherwise idle (8: TMEM allocator, 9): the epilogue warps count the column tiles added to each
//                   128-row block of X; the worker warps of ALL CTAs draw 8-row slices from a global ticket, wait until
//                   the slice's block is complete, normalise its rows (read back from L2, where the reduce-adds just left
//                   them, four rows in flight per warp) and write the fp16 A operand of the next GEMM.  The GEMM pipeline
//                   never waits for them.  Saves the stand-alone LayerNorm kernel's pass over X in HBM (359 MB per
//                   LayerNorm at ViT-L, batch 64), its launch and its 86 us.  (Round 1's version let the epilogue warps of
//                   whichever CTA finished a block do the rows, one at a time: 3x slower than the two kernels apart.)
#pragma once
#include "ptx.cuh"
#include "ln_row.cuh"

namespace dino {

enum : int { EPI_BIAS_F16 = 0, EPI_GELU_F16 = 1, EPI_RESID_F32 = 2, EPI_SWIGLU_F16 = 3, EPI_PATCH_F32 = 4, EPI_RESID_LN_F32 = 5 };

struct GemmParams {
    int M, N, K;            // N counts weight rows (for SWIGLU: gate+up rows, output has N/2 columns)
    const float *bias;      // [N]
    const float *lscale;    // [N]            (RESID)
    void *out;              // fp16 or fp32, row stride ldo elements
    int ldo;
    const float *pos;       // [1 + np, N]    (PATCH)
    int np, ntok, tok_off;  // (PATCH) patches / image, tokens / image, 1 + registers
    // (RESID_LN) LayerNorm of the finished rows of `out`: fp16 ln_out[M, N] = LN(out row) * ln_gamma + ln_beta
    const float *ln_gamma, *ln_beta;
    int ln_slice;                       // rows per LayerNorm work item (a divisor of 128; 0 = GEMM_LN_SLICE)
    __half *ln_out;
    float ln_eps;
    int *ln_count;          // column tiles added per 128-row block; ln_count[nblk .. 2 nblk) = slices normalised per block;
                            // ln_count[2 nblk] = slice ticket, [2 nblk + 1] = workers finished.  All zero on entry and on exit.
    unsigned long long a_hint;   // L2 eviction-priority hints of the A (activation) / W (weight) loads; 0 = default (normal / evict-last)
    unsigned long long b_hint;
    // Walk the tiles from the LAST row block to the first.  Consecutive kernels of the forward pass alternate their direction
    // (engine.cu), so that the ~100 MB a kernel wrote last — still in the 126 MB L2 — are what its consumer reads first.
    int reverse;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;
// EPI_RESID_LN_F32 runs 16 warps: the same 8 epilogue warps, the TMA and MMA warps, and SIX LayerNorm worker warps (8, 9, 12-15).
// One worker warp retires a row in ~1.6 k cycles of dependent arithmetic and shuffles whatever the number of rows it keeps in
// flight, so the worker count — not memory-level parallelism — sets the LayerNorm throughput; the residual epilogue needs
// 126 registers, which fits the 128 a 512-thread CTA leaves per thread.
constexpr int GEMM_THREADS_LN = 512;
constexpr int GEMM_LN_WORKERS = 6;
__host__ __device__ constexpr int gemm_threads(int epi) { return epi == EPI_RESID_LN_F32 ? GEMM_THREADS_LN : GEMM_THREADS; }
constexpr int GEMM_EPI_WARPS = 8;
// accumulator column blocks issued round-robin per k-block (1 = one MMA per k-step; 2 and 4 measured no faster)
#ifndef GEMM_NSPLIT
#define GEMM_NSPLIT 1
#endif

// CG = CTAs per tile: 1 = every CTA computes its own 128 x BN tile; 2 = a CTA pair (cluster of 2, tcgen05 cta_group::2)
// computes a 256 x BN tile: each CTA loads its 128 A rows and HALF of the BN weight rows, the pair's tensor cores read
// both halves — per-CTA shared-memory fill and operand-read traffic drop by a third, which is what limits the 1-CTA
// kernel (48 KB of TMA writes + 48 KB of operand reads per 512 MMA cycles against a 128 B/clk shared-memory port).
// LN (EPI_RESID_LN_F32): 12 KB of the operand ring become a shared-memory copy of the LayerNorm gamma / beta for the workers.
template <int BN, int CG, bool LN = false> struct GemmCfg {
    static constexpr int kABytes = GEMM_BM * GEMM_BK * 2;          // 16 KB
    static constexpr int kBBytes = (BN / CG) * GEMM_BK * 2;        // this CTA's share of the weight tile
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kGbBytes = LN ? 2 * 1536 * 4 : 0;
    static constexpr int kStages = (192 * 1024 - kGbBytes) / kStageBytes;     // 4 (BN 256, CG 1) .. 8
    static constexpr int kBarBytes = 256;
    static constexpr int kEpiBytes = GEMM_EPI_WARPS * 4096;   // one 32-row x 128-B staging tile per epilogue warp
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + kGbBytes + 1024;   // +1024: manual alignment
    static constexpr int kTmemCols = 512;
};

// tanh-GELU of an fp16-rounded input, as the reference's fp16 lookup table computes it:
// 0.5 v (1 + tanh(c v (1 + a v^2))) == v * sigmoid(2 c v (1 + a v^2))
__device__ __forceinline__ float gelu_tanh_r16(float u) {
    const float v = __half2float(__float2half_rn(u));
    const float inner = 0.79788456080286535588f * v * fmaf(0.044715f * v, v, 1.0f);
    const float e = ex2_approx(-2.885390081777927f * inner);   // exp(-2 inner)
    return v * rcp_approx(1.0f + e);
}
__device__ __forceinline__ float silu_f32(float x) {
    return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}

// LayerNorm worker of the residual GEMM (EPI_RESID_LN_F32), one warp.  Kept out of line and handed plain values: its row buffers
// get their own register allocation instead of competing with the epilogue's inside one 168-register kernel body.
// Slices of GEMM_LN_SLICE rows are drawn from a global ticket in the order the row blocks complete; a slice waits until all column tiles
// of its 128-row block have been added (counter published by the epilogue warps, one tile late), then its rows are read back
// from L2, R at a time, and normalised.  Nothing in the GEMM pipeline waits for these warps.
// Widths: exactly 128 * NV4 columns (384, 768, 1024, 1536 = every DINOv2 hidden size; gemm_ln_width_ok).
#ifndef GEMM_LN_SLICE
#define GEMM_LN_SLICE 16
#endif
__host__ __device__ constexpr bool gemm_ln_width_ok(int N) { return N == 384 || N == 768 || N == 1024 || N == 1536; }

template <int NV4, int R>
__device__ __noinline__ void gemm_ln_worker(const float *X, size_t ldx, const float4 *gs, __half *ln_out,
                                            int *ln_count, int M, float eps, int num_n, int reverse, int slice_rows, int lane) {
    const int kSlicesPerBlk = GEMM_BM / slice_rows;
    constexpr int N = NV4 * 128;
    const int n_blk128 = (M + GEMM_BM - 1) / GEMM_BM;
    const int n_slices = (M + slice_rows - 1) / slice_rows;
    int *ln_done = ln_count + n_blk128;
    int *ticket = ln_count + 2 * n_blk128;
    int *finished = ticket + 1;
#ifdef GEMM_LN_PROF
    long long t_tick = 0, t_wait = 0, t_rows = 0, t_done = 0, n_sl = 0;
    long long tq = clock64();
#define LNP(ACC) do { const long long now__ = clock64(); ACC += now__ - tq; tq = now__; } while (0)
#else
#define LNP(ACC) do {} while (0)
#endif
    // The ticket of the NEXT slice and the completion count of the PREVIOUS one are in flight while a slice's rows are
    // processed: their L2 round trips (1-2 k cycles each under load) stay off the worker's critical path.
    int w_next = 0, done_prev = 0, blk_prev = -1;
    if (lane == 0) w_next = atomicAdd(ticket, 1);
    auto retire = [&]() {                                        // the last slice of a block puts the block's counters back to zero
        if (lane == 0 && blk_prev >= 0) {
            const int slices_in_blk = min(kSlicesPerBlk, n_slices - blk_prev * kSlicesPerBlk);
            if (done_prev == slices_in_blk - 1) {
                ln_count[blk_prev] = 0;
                ln_done[blk_prev] = 0;
            }
        }
    };
    while (true) {
        const int w = __shfl_sync(0xffffffffu, w_next, 0);
        LNP(t_tick);
        if (w >= n_slices) break;
        if (lane == 0) w_next = atomicAdd(ticket, 1);
        const int sl = reverse ? n_slices - 1 - w : w;
        const int blk = sl / kSlicesPerBlk;
        if (lane == 0) {
#ifdef GEMM_LN_SC_FENCE
            const volatile int *cnt = ln_count + blk;
            while (*cnt < num_n) __nanosleep(64);
            __threadfence();
#else
            while (ld_acquire_gpu(ln_count + blk) < num_n) __nanosleep(64);   // acquire: the rows added before the counter reached num_n
#endif
        }
        __syncwarp();
        LNP(t_wait);
        const int r0 = sl * slice_rows, nr = min(slice_rows, M - r0);
        const float *xs = X + static_cast<size_t>(r0) * ldx;
        __half *os = ln_out + static_cast<size_t>(r0) * N;
#pragma unroll 1
        for (int r = 0; r < nr; r += R)
            layernorm_rows_l2<NV4, R>(xs + static_cast<size_t>(r) * ldx, ldx, gs, os + static_cast<size_t>(r) * N, eps, lane, nr - r);
        LNP(t_rows);
        retire();
        if (lane == 0) done_prev = atomicAdd(ln_done + blk, 1);
        blk_prev = blk;
#ifdef GEMM_LN_PROF
        LNP(t_done);
        ++n_sl;
#endif
    }
    retire();
#ifdef GEMM_LN_PROF
    if (blockIdx.x == 0 && (threadIdx.x >> 5) == 8 && lane == 0) {
        int *o = ticket + 2;
        o[0] = static_cast<int>(t_tick); o[1] = static_cast<int>(t_wait); o[2] = static_cast<int>(t_rows); o[3] = static_cast<int>(t_done); o[4] = static_cast<int>(n_sl);
    }
#endif
    // ... and the last worker of the grid the ticket
    if (lane == 0) {
        if (atomicAdd(finished, 1) == static_cast<int>(gridDim.x) * GEMM_LN_WORKERS - 1) {
            *ticket = 0;
            *finished = 0;
            __threadfence();
        }
    }
}

#ifndef GEMM_LN_R8
#define GEMM_LN_R8 3     // rows in flight per worker warp at width 1024 (96 data registers of the 128 a 512-thread CTA allows)
#endif
__device__ __forceinline__ void gemm_ln_worker_dispatch(const GemmParams &p, const float4 *gs, int num_n, int lane) {
    const float *X = reinterpret_cast<const float *>(p.out);
    const int slice_rows = p.ln_slice > 0 ? p.ln_slice : GEMM_LN_SLICE;
    switch (p.N) {
    case 384:  gemm_ln_worker<3, 8>(X, p.ldo, gs, p.ln_out, p.ln_count, p.M, p.ln_eps, num_n, p.reverse, slice_rows, lane); break;
    case 768:  gemm_ln_worker<6, 4>(X, p.ldo, gs, p.ln_out, p.ln_count, p.M, p.ln_eps, num_n, p.reverse, slice_rows, lane); break;
    case 1024: gemm_ln_worker<8, GEMM_LN_R8>(X, p.ldo, gs, p.ln_out, p.ln_count, p.M, p.ln_eps, num_n, p.reverse, slice_rows, lane); break;
    case 1536: gemm_ln_worker<12, 2>(X, p.ldo, gs, p.ln_out, p.ln_count, p.M, p.ln_eps, num_n, p.reverse, slice_rows, lane); break;
    default: break;                                              // the host refuses other widths (gemm_ln_width_ok)
    }
}

// MC = 2 (only with CG = 2): clusters of FOUR CTAs = two CTA pairs that work on the same weight tile and on vertically adjacent
// 256-row blocks of A.  Every CTA fetches a quarter of the 256 x 64 weight tile per k-block and TMA-multicasts it to the
// CTA that holds the same half in the other pair, so the weight bytes cross the L2 -> SM fabric once per cluster instead of
// once per pair (-25 % operand traffic per FLOP).  A stage is refilled only after BOTH pairs' MMAs have released it.
template <int BN, int EPI, int CG, int MC = 1>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_f16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const GemmParams p) {
    using Cfg = GemmCfg<BN, CG, EPI == EPI_RESID_LN_F32>;
    static_assert(CG == 1 || CG == 2, "one CTA or a CTA pair per tile");
    static_assert(MC == 1 || (MC == 2 && CG == 2), "weight multicast couples two CTA pairs");
    static_assert(EPI != EPI_SWIGLU_F16 || BN == 256, "SwiGLU tiles pair 128 gate + 128 up columns");
    constexpr int kStages = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + kStages * Cfg::kABytes;
    uint8_t *smem_epi = smem + kStages * Cfg::kStageBytes;
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_epi + Cfg::kEpiBytes);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *tmem_full = empty_bar + kStages;
    uint64_t *tmem_empty = tmem_full + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(tmem_empty + 2);
    float4 *smem_gb = reinterpret_cast<float4 *>(smem_epi + Cfg::kEpiBytes + Cfg::kBarBytes);   // gamma[N] | beta[N] (EPI_RESID_LN_F32)
    constexpr bool kResid = EPI == EPI_RESID_F32 || EPI == EPI_RESID_LN_F32;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int num_m = (p.M + GEMM_BM * CG * MC - 1) / (GEMM_BM * CG * MC);   // row blocks of 128 * CG * MC rows (one per cluster)
    const int num_n = (p.N + BN - 1) / BN;
    const int num_k = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int num_tiles = num_m * num_n;
    const uint32_t cluster_rank = CG == 2 ? cluster_ctarank() : 0u;     // rank in the cluster of CG * MC CTAs
    const uint32_t cta_rank = cluster_rank & 1u;                        // rank in the CTA pair: 0 = leader (issues the MMAs)
    const uint32_t pair_idx = cluster_rank >> 1;                        // which pair of the cluster (MC = 2)
    const uint32_t leader_rank = cluster_rank & ~1u;                    // cluster rank of this pair's leader
    const int tile0 = static_cast<int>(blockIdx.x) / (CG * MC);         // first tile of this cluster
    const int tile_step = static_cast<int>(gridDim.x) / (CG * MC);

    if (warp == 10 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        if constexpr (EPI != EPI_PATCH_F32) prefetch_tmap(&tmC);
    }
    if (warp == 11 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], CG);                      // pair: the leader's own expect_tx arrive + the peer's arrive
            mbar_init(&empty_bar[s], MC);                      // multicast: both pairs' MMA warps release a stage
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], GEMM_EPI_WARPS * CG);   // pair: the epilogue warps of both CTAs free the accumulator
        }
        fence_mbar_init();
    }
    if (warp == 8) {
        if constexpr (CG == 2) tmem_alloc_2sm(tmem_ptr, Cfg::kTmemCols);
        else tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    }
    if constexpr (EPI == EPI_RESID_LN_F32) {                  // weights, not a product of the previous kernel: no dependency wait
        const int nv = p.N >> 2;
        for (int i = threadIdx.x; i < nv; i += gemm_threads(EPI)) {
            smem_gb[i] = __ldg(reinterpret_cast<const float4 *>(p.ln_gamma) + i);
            smem_gb[nv + i] = __ldg(reinterpret_cast<const float4 *>(p.ln_beta) + i);
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();                // barrier inits visible to the peer before any remote arrive
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    griddep_wait();                    // everything above overlapped the previous kernel's tail (programmatic dependent launch)
    griddep_launch_dependents();

    if (warp == 10) {
        // ------------------------------------------------------------ TMA producer
        // The whole warp walks the loop (so that addresses and coordinates stay in uniform registers); one elected
        // lane issues the copies.
        int s = 0;
        uint32_t ph = 0;
        // Activations: NO evict-first — an A block is re-read by the other column tiles of its row block over several waves, and
        // the streaming output stores would push evict-first lines out in between (measured: fc1 fetched its A operand ~4 times
        // from DRAM, every GEMM 3-4 % slower).  Weights are shared by every CTA for the whole launch: evict-last.
        const uint64_t a_hint = p.a_hint ? p.a_hint : kEvictNormal;
        const uint64_t b_hint = p.b_hint ? p.b_hint : kEvictLast;
        for (int t = tile0; t < num_tiles; t += tile_step) {
            const int tt = p.reverse ? num_tiles - 1 - t : t;
            const int m_blk = tt / num_n, n_blk = tt % num_n;
            const int row_a = ((m_blk * MC + static_cast<int>(pair_idx)) * CG + static_cast<int>(cta_rank)) * GEMM_BM;   // this CTA's 128 A rows
            const int row_b = n_blk * BN + static_cast<int>(cta_rank) * (BN / CG);          // this CTA's share of the weight rows
            for (int kb = 0; kb < num_k; ++kb) {
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (elect_one()) {
                    if constexpr (CG == 2) {
                        const uint32_t full0 = mapa_shared(&full_bar[s], leader_rank);       // the leader's barrier counts both CTAs' bytes
                        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
                        else mbar_arrive_cluster(full0);
                        tma_load_2d_2sm(smem_a + s * Cfg::kABytes, &tmA, full0, kb * GEMM_BK, row_a, a_hint);
                        if constexpr (MC == 2) {
                            // this CTA's quarter of the weight tile, delivered to itself and to its counterpart in the other pair
                            constexpr int kQRows = BN / CG / MC;
                            const uint16_t mask = static_cast<uint16_t>((1u << cluster_rank) | (1u << (cluster_rank ^ 2u)));
                            tma_load_2d_2sm_mc(smem_b + s * Cfg::kBBytes + pair_idx * (kQRows * GEMM_BK * 2), &tmB, full0, kb * GEMM_BK,
                                               row_b + static_cast<int>(pair_idx) * kQRows, mask, b_hint);
                        } else {
                            tma_load_2d_2sm(smem_b + s * Cfg::kBBytes, &tmB, full0, kb * GEMM_BK, row_b, b_hint);
                        }
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
                        tma_load_2d_hint(smem_a + s * Cfg::kABytes, &tmA, &full_bar[s], kb * GEMM_BK, row_a, a_hint);
                        tma_load_2d_hint(smem_b + s * Cfg::kBBytes, &tmB, &full_bar[s], kb * GEMM_BK, row_b, b_hint);
                    }
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 11) {
        // ------------------------------------------------------------ MMA issuer
        // All 32 lanes run the control flow and the descriptor arithmetic (warp-uniform -> uniform datapath, no
        // per-instruction R2UR traffic); a single elected lane issues tcgen05.mma / tcgen05.commit.
        // Optionally the tile's accumulator is split into NSPLIT independent column blocks whose MMAs are issued
        // round-robin (experiment: accumulate-dependent tcgen05.mma do pipeline — splitting bought nothing).
        constexpr int NSPLIT = GEMM_NSPLIT;
        constexpr int NSUB = BN / NSPLIT;                                     // accumulator columns per chain
        constexpr uint32_t idesc = make_idesc_f16(GEMM_BM * CG, NSUB, 0, 0);
        constexpr uint32_t b_sub = (NSUB / CG) * GEMM_BK * 2 / 16;              // descriptor units between the chains' B rows
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        if (cta_rank == 0) {                                   // pair: only the leader issues; its MMAs drive both tensor cores
            for (int t = tile0; t < num_tiles; t += tile_step, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t a_desc = make_smem_desc_sw128(smem_u32(smem_a + s * Cfg::kABytes), 16, 1024);
                    const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_b + s * Cfg::kBBytes), 16, 1024);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
#pragma unroll
                            for (int h = 0; h < NSPLIT; ++h) {
                                // +32 bytes (16 halves) along K inside the 128-B swizzle atom == +2 in the address field
                                if constexpr (CG == 2)
                                    umma_f16_ss_2sm(d_tmem + h * NSUB, a_desc + 2 * k, b_desc + h * b_sub + 2 * k, idesc, (kb | k) != 0);
                                else
                                    umma_f16_ss(d_tmem + h * NSUB, a_desc + 2 * k, b_desc + h * b_sub + 2 * k, idesc, (kb | k) != 0);
                            }
                        }
                        // frees the smem stage (in both CTAs of a pair) once these MMAs retire; after the last k-block the
                        // accumulator is complete -> epilogue warps (of both CTAs)
                        if constexpr (CG == 2) {
                            // the stage is released in every CTA that wrote into this pair's shared memory (multicast: all four)
                            umma_commit_2sm(&empty_bar[s], MC == 2 ? 0xF : 3);
                            if (kb == num_k - 1) umma_commit_2sm(&tmem_full[as], static_cast<uint16_t>(3u << leader_rank));
                        } else {
                            umma_commit(&empty_bar[s]);
                            if (kb == num_k - 1) umma_commit(&tmem_full[as]);
                        }
                    }
                    __syncwarp();
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (EPI == EPI_RESID_LN_F32 && (warp == 8 || warp == 9 || warp >= 12)) {
        // ------------------------------------------------------------ LayerNorm workers (the two warps with no other job)
        gemm_ln_worker_dispatch(p, smem_gb, num_n, lane);
    } else if (warp < 8) {
        // ------------------------------------------------------------ epilogue
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int half = warp >> 2;          // which half of the tile's columns
        uint8_t *stage = smem_epi + warp * 4096;                 // this warp's staging tile (1024-B aligned)
        uint8_t *stage_row = stage + lane * 128;
        const uint32_t sw = static_cast<uint32_t>(lane & 7);            // 128-B swizzle: chunk c of row r lives at c ^ (r & 7)
        // the previous TMA store of this warp must have finished reading the staging tile before it is rewritten
        auto stage_acquire = [&]() {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
        };
        auto stage_release = [&](int c0, int c1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if constexpr (kResid) tma_reduce_add_2d(&tmC, stage, c0, c1);
                else tma_store_2d(&tmC, stage, c0, c1);
                bulk_commit();
            }
        };
        // (RESID_LN) row block whose tile count is still to be published, and the publication itself: wait for the reduce-adds
        // of that tile (all but the newest kResidSteps bulk groups of this warp — every warp commits exactly that many per
        // tile, N being a multiple of BN), count the tile, and let the CTA that counted the last one normalise the rows.
        int ln_pending = -1;
        constexpr int kResidSteps = BN / 2 / 32;
        auto ln_publish = [&](int blk, bool final_flush) {
            if (lane == 0) {
                if (final_flush) bulk_wait<0>();
                else bulk_wait<kResidSteps>();
            }
            __syncwarp();
            named_bar_sync(1, GEMM_EPI_WARPS * 32);               // every epilogue warp's reduce-adds of that tile have been performed
            if (threadIdx.x == 0 && blk * GEMM_BM < p.M) {        // (the second CTA of the last pair may hold no rows at all)
                fence_proxy_async_all();
#ifdef GEMM_LN_SC_FENCE
                __threadfence();
                atomicAdd(p.ln_count + blk, 1);
#else
                red_add_release_gpu(p.ln_count + blk, 1);           // one more column tile of this 128-row block is in X
#endif
            }
        };
        int it = 0;
        for (int t = tile0; t < num_tiles; t += tile_step, ++it) {
            const int tt = p.reverse ? num_tiles - 1 - t : t;
            const int m_blk = tt / num_n, n_blk = tt % num_n;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            const int row0 = ((m_blk * MC + static_cast<int>(pair_idx)) * CG + static_cast<int>(cta_rank)) * GEMM_BM + q * 32;   // first row of this warp's 32-row slab
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

            if constexpr (EPI == EPI_SWIGLU_F16) {
                // 64 output columns per warp: gate columns [64 half, +64), up columns 128 + the same
                const int oc = n_blk * 128 + half * 64;
                if (oc < p.N / 2) {
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        const int lc = half * 64 + c * 32;
                        uint32_t g[32], u[32];
                        tmem_ld_32x32b_x32(t_row + lc, g);
                        tmem_ld_32x32b_x32(t_row + 128 + lc, u);
                        tmem_ld_wait();
                        if (c == 0) stage_acquire();
                        const int ng = n_blk * BN + lc;
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            uint32_t pk[4];
#pragma unroll
                            for (int e = 0; e < 8; e += 2) {
                                const int j = 8 * v + e;
                                const float g0 = __uint_as_float(g[j]) + __ldg(p.bias + ng + j);
                                const float g1 = __uint_as_float(g[j + 1]) + __ldg(p.bias + ng + j + 1);
                                const float u0 = __uint_as_float(u[j]) + __ldg(p.bias + ng + 128 + j);
                                const float u1 = __uint_as_float(u[j + 1]) + __ldg(p.bias + ng + 128 + j + 1);
                                pk[e >> 1] = pack_half2(silu_f32(g0) * u0, silu_f32(g1) * u1);
                            }
                            *reinterpret_cast<uint4 *>(stage_row + (((c * 4 + v) ^ sw) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    stage_release(oc, row0);
                }
            } else if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_GELU_F16) {
                constexpr int kSteps = BN / 2 / 64;                      // 64 fp16 columns (one 128-B row) per TMA store
#pragma unroll 1
                for (int c = 0; c < kSteps; ++c) {
                    const int lc = half * (BN / 2) + c * 64;
                    const int col = n_blk * BN + lc;
                    if (col >= p.N) break;                               // warp-uniform
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32b_x32(t_row + lc, r0);
                    tmem_ld_32x32b_x32(t_row + lc + 32, r1);
                    tmem_ld_wait();
                    stage_acquire();
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const uint32_t *r = v < 4 ? &r0[8 * v] : &r1[8 * (v - 4)];
                        float x[8];
                        if (col + 8 * v < p.N) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4 *>(p.bias + col + 8 * v));
                            const float4 b1 = __ldg(reinterpret_cast<const float4 *>(p.bias + col + 8 * v + 4));
                            x[0] = __uint_as_float(r[0]) + b0.x; x[1] = __uint_as_float(r[1]) + b0.y;
                            x[2] = __uint_as_float(r[2]) + b0.z; x[3] = __uint_as_float(r[3]) + b0.w;
                            x[4] = __uint_as_float(r[4]) + b1.x; x[5] = __uint_as_float(r[5]) + b1.y;
                            x[6] = __uint_as_float(r[6]) + b1.z; x[7] = __uint_as_float(r[7]) + b1.w;
                            if constexpr (EPI == EPI_GELU_F16) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) x[e] = gelu_tanh_r16(x[e]);
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) x[e] = 0.f;      // clipped by the TMA store anyway
                        }
                        *reinterpret_cast<uint4 *>(stage_row + ((v ^ sw) << 4)) =
                            make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
                    }
                    stage_release(col, row0);
                }
            } else if constexpr (kResid) {
                constexpr int kSteps = BN / 2 / 32;                      // 32 fp32 columns (one 128-B row) per TMA reduce
#pragma unroll 1
                for (int c = 0; c < kSteps; ++c) {
                    const int lc = half * (BN / 2) + c * 32;
                    const int col = n_blk * BN + lc;
                    if (col >= p.N) break;
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_row + lc, r);
                    tmem_ld_wait();
                    stage_acquire();
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (col + 4 * v < p.N) {
                            const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col + 4 * v));
                            const float4 ls = __ldg(reinterpret_cast<const float4 *>(p.lscale + col + 4 * v));
                            y = make_float4((__uint_as_float(r[4 * v + 0]) + b.x) * ls.x, (__uint_as_float(r[4 * v + 1]) + b.y) * ls.y,
                                            (__uint_as_float(r[4 * v + 2]) + b.z) * ls.z, (__uint_as_float(r[4 * v + 3]) + b.w) * ls.w);
                        }
                        *reinterpret_cast<float4 *>(stage_row + ((v ^ sw) << 4)) = y;
                    }
                    stage_release(col, row0);                            // X[row0.., col..] += staged tile
                }
            } else {   // EPI_PATCH_F32: rows scatter to (image, token) positions, written directly
                const int row = row0 + lane;
                const bool row_ok = row < p.M;
                const int img = row / p.np, pp = row - img * p.np;
                const size_t orow = static_cast<size_t>(img) * p.ntok + p.tok_off + pp;
                const float *pos_row = p.pos + static_cast<size_t>(1 + pp) * p.N;
                constexpr int kChunks = BN / 2 / 32;
#pragma unroll 1
                for (int c = 0; c < kChunks; ++c) {
                    const int lc = half * (BN / 2) + c * 32;
                    const int col = n_blk * BN + lc;
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(t_row + lc, r);
                    tmem_ld_wait();
                    if (row_ok && col < p.N) {
                        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p.out) + orow * p.ldo + col);
#pragma unroll
                        for (int v = 0; v < 8; ++v) {
                            if (col + 4 * v < p.N) {
                                const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col + 4 * v));
                                const float4 pe = __ldg(reinterpret_cast<const float4 *>(pos_row + col + 4 * v));
                                dst[v] = make_float4(__uint_as_float(r[4 * v + 0]) + b.x + pe.x, __uint_as_float(r[4 * v + 1]) + b.y + pe.y,
                                                     __uint_as_float(r[4 * v + 2]) + b.z + pe.z, __uint_as_float(r[4 * v + 3]) + b.w + pe.w);
                            }
                        }
                    }
                }
            }
            // all TMEM reads of this accumulator stage are complete (wait::ld above): hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(&tmem_empty[as], leader_rank));   // the leader's MMA warp waits on it
                else mbar_arrive(&tmem_empty[as]);
            }
            if constexpr (EPI == EPI_RESID_LN_F32) {
                // This CTA has added one more column tile to its 128 rows of X.  The row block is published (and, if this was
                // its last tile, normalised) one tile LATER, when its reduce-adds have certainly been performed: waiting for
                // them right here would put an L2 round trip into every tile's epilogue.
                if (ln_pending >= 0) ln_publish(ln_pending, false);
                ln_pending = (m_blk * MC + static_cast<int>(pair_idx)) * CG + static_cast<int>(cta_rank);
            }
        }
        if constexpr (EPI == EPI_RESID_LN_F32) {
            if (ln_pending >= 0) ln_publish(ln_pending, true);
        }
        if (lane == 0) bulk_wait<0>();       // staging smem must outlive the last TMA store; global writes complete
    }

    griddep_launch_dependents_late();
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all();    // the peer may still read this CTA's shared memory / arrive on its barriers
    else __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        if constexpr (CG == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace dino
