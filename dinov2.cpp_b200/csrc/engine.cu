// dinov2_b200 engine: weight residency, activation arena, kernel orchestration and the extern "C" ABI
// declared in include/dinov2_b200.h.  Replaces the reference's graph build + ggml backend execution
// (reference dinov2.cpp:616-838 forward_features/forward_head, :900-999 dino_predict) with a fixed fused
// pipeline of hand-written sm_100a kernels.  There is no CPU fallback: without an sm_100 device every
// entry point fails with DINO_B200_ERR_NO_DEVICE.
#include "../../include/dinov2_b200.h"

#include "attention_common.cuh"
// Superseded attention generations (v3, v5, v7), the weight-multicast GEMM (MC = 2) and the L2-read-back LayerNorm epilogue are
// measurement history: they are compiled only with -DDINO_B200_EXPERIMENTAL (tools/build_experimental.sh), never into the product .so.
#ifdef DINO_B200_EXPERIMENTAL
#include "attention3.cuh"
#include "attention5.cuh"
#include "attention7.cuh"
#include "attention8.cuh"
#endif
#include "attention10.cuh"
#include "elementwise.cuh"
#include "gemm.cuh"
#include "gguf_reader.hpp"
#include "pca.cuh"
#include "quantize.hpp"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace dino {

// ------------------------------------------------------------------------------------------------ errors
struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct StatusError : std::runtime_error {
    dino_b200_status st;
    StatusError(dino_b200_status s, const std::string &m) : std::runtime_error(m), st(s) {}
};

#define DINO_CUDA(expr)                                                                                        \
    do {                                                                                                       \
        cudaError_t err__ = (expr);                                                                            \
        if (err__ != cudaSuccess)                                                                              \
            throw dino::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(err__) + " (" __FILE__ \
                                  ":" + std::to_string(__LINE__) + ")");                                       \
    } while (0)

static thread_local std::string g_last_error;

// ------------------------------------------------------------------------------------------------ TMA maps
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
    });
    if (!fn) throw CudaError("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

// Row-major [rows, cols] tensor with row stride ld (elements); box = box_cols x box_rows with box_cols * elem = 128 B,
// 128-B swizzle; out-of-bounds elements read as zero / are clipped on store.
static CUtensorMap make_tmap_2d(const void *ptr, bool f32, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows,
                                CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
    CUtensorMap m;
    const size_t es = f32 ? sizeof(float) : sizeof(__half);
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {ld * es};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / es), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (gstride[0] & 15)) throw CudaError("TMA operand is not 16-byte aligned");
    const CUresult r = get_encode_fn()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                                       const_cast<void *>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_128B, promo,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
    return m;
}
// The attention kernel's view of QKV: 128-byte row segments (one head's 64 dims) 6*D bytes apart.  Promoting those to
// 256-byte L2 fetches (right for the GEMM operands, whose next k-block is the neighbouring 128 bytes) would drag in the
// NEXT head's data, which is evicted again before its turn: DINO_B200_ATTN_PROMO=256 restores that for A/B comparisons.
static CUtensorMapL2promotion attn_promotion() {
    static const CUtensorMapL2promotion v = [] {
        const char *e = getenv("DINO_B200_ATTN_PROMO");
        if (e && e[0] == '2') return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
        if (e && e[0] == '0') return CU_TENSOR_MAP_L2_PROMOTION_NONE;
        if (e && e[0] == '6') return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
        return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    return v;
}
// fp16 [images, rows, cols] (row stride ld elements, image stride rows * ld): box = 64 columns x box_rows rows of ONE image, 128-B
// swizzle.  Used for the attention output: a query tile that straddles the last token of an image is clipped per image.
static CUtensorMap make_tmap_3d_f16(const void *ptr, uint64_t cols, uint64_t rows, uint64_t images, uint64_t ld, uint32_t box_rows) {
    CUtensorMap m;
    const cuuint64_t gdim[3] = {cols, rows, images};
    const cuuint64_t gstride[2] = {ld * sizeof(__half), rows * ld * sizeof(__half)};
    const cuuint32_t box[3] = {64, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (gstride[0] & 15) || (gstride[1] & 15)) throw CudaError("TMA operand is not 16-byte aligned");
    const CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(ptr), gdim, gstride, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw CudaError("cuTensorMapEncodeTiled (3-D) failed with code " + std::to_string(static_cast<int>(r)));
    return m;
}
static CUtensorMap make_tmap_f16(const void *ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint32_t box_rows) {
    return make_tmap_2d(ptr, false, cols, rows, ld, box_rows);
}
// output map of a GEMM epilogue: 32-row x 128-B boxes (one per epilogue warp and step)
static CUtensorMap make_tmap_out(int epi, const void *out, uint64_t cols, uint64_t rows, uint64_t ld) {
    return make_tmap_2d(out, epi == EPI_RESID_F32 || epi == EPI_RESID_LN_F32, cols, rows, ld, 32);
}

// ------------------------------------------------------------------------------------------------ launches

// DINO_B200_GEMM_CG=1 selects the one-CTA-per-tile GEMM (A/B comparisons); default: CTA pairs (cta_group::2)
static int gemm_cg() {
    static int v = [] {
        const char *e = getenv("DINO_B200_GEMM_CG");
        return (e && e[0] == '1') ? 1 : 2;
    }();
    return v;
}

#ifdef DINO_B200_EXPERIMENTAL
template <int EPI> static void configure_gemm_mc() {
    DINO_CUDA(cudaFuncSetAttribute(gemm_f16_tcgen05<256, EPI, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, 2>::kSmemBytes));
}
#endif
template <int BN, int EPI> static void configure_gemm() {
    DINO_CUDA(cudaFuncSetAttribute(gemm_f16_tcgen05<BN, EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, 1, EPI == EPI_RESID_LN_F32>::kSmemBytes));
    DINO_CUDA(cudaFuncSetAttribute(gemm_f16_tcgen05<BN, EPI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN, 2, EPI == EPI_RESID_LN_F32>::kSmemBytes));
}
// Per-device state: cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the CURRENT device only, and SM counts may
// differ between devices, so both are tracked per ordinal (one process may hold engines on several GPUs).
constexpr int kMaxDevices = 64;
struct DeviceState {
    bool configured = false;
    int num_sms = 0;
};
static std::mutex g_dev_mutex;
static DeviceState g_dev[kMaxDevices];

static void configure_current_device_locked(DeviceState &d, int dev) {
    DINO_CUDA(cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev));
#ifdef DINO_B200_EXPERIMENTAL
    configure_gemm_mc<EPI_BIAS_F16>();
    configure_gemm_mc<EPI_GELU_F16>();
    configure_gemm_mc<EPI_RESID_F32>();
    configure_gemm_mc<EPI_SWIGLU_F16>();
    DINO_CUDA(cudaFuncSetAttribute(attention_fwd_v3, cudaFuncAttributeMaxDynamicSharedMemorySize, AT3_SMEM_BYTES));
    DINO_CUDA(cudaFuncSetAttribute(attention_fwd_v5, cudaFuncAttributeMaxDynamicSharedMemorySize, AT5_SMEM_BYTES));
    DINO_CUDA(cudaFuncSetAttribute(attention_fwd_v7, cudaFuncAttributeMaxDynamicSharedMemorySize, AT7_SMEM_BYTES));
    DINO_CUDA(cudaFuncSetAttribute(attention_fwd_v8, cudaFuncAttributeMaxDynamicSharedMemorySize, AT8_SMEM_BYTES));
#endif
    configure_gemm<256, EPI_BIAS_F16>();
    configure_gemm<128, EPI_BIAS_F16>();
    configure_gemm<256, EPI_GELU_F16>();
    configure_gemm<128, EPI_GELU_F16>();
    configure_gemm<256, EPI_RESID_F32>();
    configure_gemm<128, EPI_RESID_F32>();
    configure_gemm<256, EPI_RESID_LN_F32>();
    configure_gemm<128, EPI_RESID_LN_F32>();
    configure_gemm<256, EPI_SWIGLU_F16>();
    configure_gemm<256, EPI_PATCH_F32>();
    configure_gemm<128, EPI_PATCH_F32>();
    DINO_CUDA(cudaFuncSetAttribute(attention_fwd_v10, cudaFuncAttributeMaxDynamicSharedMemorySize, AT10_SMEM_BYTES));
    d.configured = true;
}
// Opt-in shared-memory sizes for the current device (idempotent); returns its SM count.
static int configure_current_device() {
    int dev = 0;
    DINO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) throw CudaError("device ordinal out of range");
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    DeviceState &d = g_dev[dev];
    if (!d.configured) configure_current_device_locked(d, dev);
    return d.num_sms;
}

// ------------------------------------------------------------------------------------------------ launch helper
// Programmatic dependent launch for the kernels of a forward pass (every one of them executes griddepcontrol.wait before its
// first global access): set by forward_enqueue for the duration of the pass, never by the stand-alone kernel hooks.
// DINO_B200_PDL=0 switches it off.
static thread_local bool t_pdl = false;
static bool pdl_enabled() {
    static const bool v = [] { const char *e = getenv("DINO_B200_PDL"); return !(e && e[0] == '0'); }();
    return v;
}
// largest workload (token rows) that is launched with programmatic dependent launch; DINO_B200_PDL_ROWS overrides (A/B runs)
static int pdl_max_rows() {
    static const int v = [] { const char *e = getenv("DINO_B200_PDL_ROWS"); return e ? atoi(e) : 12288; }();
    return v;
}
// largest workload (token rows) whose LayerNorms run inside the residual GEMMs (EPI_RESID_LN_F32) when DINO_B200_FUSE_LN is
// unset: 0 = never (at full batch the fused kernels cost what the two apart do)
static int fuse_ln_max_rows() {
    static const int v = [] { const char *e = getenv("DINO_B200_FUSE_LN_ROWS"); return e ? atoi(e) : 0; }();
    return v;
}
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = t_pdl ? 1 : 0;
    DINO_CUDA(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

static int pick_bn(int epi, int N) {
    if (epi == EPI_SWIGLU_F16) return 256;
    return (N % 256 == 0) ? 256 : 128;
}

// Tile shape for one GEMM call.  Large problems (the batched path) use the widest tile the epilogue allows on CTA pairs;
// when that leaves most SMs without a tile (single-image inference: M = 1370 rows is 6 row blocks of 256) smaller tiles /
// single CTAs are chosen so that more SMs take part.  DINO_B200_GEMM_CG=1 forces single-CTA tiles (A/B comparisons).
struct GemmPlan {
    int BN, CG;
    int MC = 1;     // 2: clusters of two CTA pairs that share (TMA-multicast) the weight tile
    int sms = 0;    // SM count of the device the plan was made for (persistent grid size)
};
// DINO_B200_GEMM_MC=1 enables the weight-multicast variant for the large (256-wide, CTA-pair) tiles
static int gemm_mc() {
#ifdef DINO_B200_EXPERIMENTAL
    static int v = [] {
        const char *e = getenv("DINO_B200_GEMM_MC");
        return (e && e[0] == '1') ? 2 : 1;
    }();
    return v;
#else
    return 1;
#endif
}
static GemmPlan plan_gemm(int epi, int M, int N, int num_sms) {
    const int bn_big = pick_bn(epi, N);
    const GemmPlan cand[3] = {{bn_big, gemm_cg()}, {bn_big, 1}, {128, 1}};
    const int n_cand = (epi == EPI_SWIGLU_F16) ? 2 : 3;
    GemmPlan best = cand[0];
    best.sms = num_sms;
    int best_busy = -1;
    for (int i = 0; i < n_cand; ++i) {
        const GemmPlan &c = cand[i];
        const int tiles = ((M + GEMM_BM * c.CG - 1) / (GEMM_BM * c.CG)) * ((N + c.BN - 1) / c.BN);
        const int busy = std::min(tiles * c.CG, num_sms);
        if (busy * 10 >= num_sms * 9) {           // enough work for (nearly) every SM: the widest such tile
            GemmPlan r = c;
            r.sms = num_sms;
            if (i == 0 && c.CG == 2 && c.BN == 256 && gemm_mc() == 2 && num_sms % 4 == 0 && epi != EPI_PATCH_F32 &&
                epi != EPI_RESID_LN_F32 && tiles >= num_sms)
                r.MC = 2;
            return r;
        }
        if (busy > best_busy) {
            best = c;
            best.sms = num_sms;
            best_busy = busy;
        }
    }
    return best;
}

template <int BN, int EPI, int CG, int MC = 1>
static void launch_gemm_t(const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const GemmParams &p, int num_sms, cudaStream_t st) {
    const int tiles = ((p.M + GEMM_BM * CG * MC - 1) / (GEMM_BM * CG * MC)) * ((p.N + BN - 1) / BN);   // per cluster
    // DINO_B200_GEMM_SMS=n restricts the persistent grid to n SMs (experiments: per-SM throughput vs L2 bandwidth share)
    static const int sm_cap = [] { const char *e = getenv("DINO_B200_GEMM_SMS"); return e ? atoi(e) : 0; }();
    const int sms = sm_cap > 0 ? std::min(sm_cap, num_sms) : num_sms;
    const int grid = std::max(1, std::min(tiles, sms / (CG * MC))) * CG * MC;      // persistent: one cluster per CG * MC SMs
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(dino::gemm_threads(EPI));
    cfg.dynamicSmemBytes = GemmCfg<BN, CG, EPI == EPI_RESID_LN_F32>::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG * MC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = t_pdl ? 2 : 1;
    DINO_CUDA(cudaLaunchKernelEx(&cfg, gemm_f16_tcgen05<BN, EPI, CG, MC>, tmA, tmB, tmC, p));
}

// tmC: output map (make_tmap_out) for every epilogue except PATCH, which scatters rows and ignores it
static void launch_gemm(int epi, GemmPlan plan, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmC, const GemmParams &p_in,
                        cudaStream_t st) {
    // DINO_B200_GEMM_AHINT / DINO_B200_GEMM_BHINT = first|normal|last override the L2 eviction hints of the activation / weight
    // loads (defaults: normal / last; measured in tools/gemm_bench.py)
    auto hint_of = [](const char *name) -> unsigned long long {
        const char *e = getenv(name);
        if (e && e[0] == 'f') return kEvictFirst;
        if (e && e[0] == 'n') return kEvictNormal;
        if (e && e[0] == 'l') return kEvictLast;
        return 0ull;
    };
    static const unsigned long long a_hint = hint_of("DINO_B200_GEMM_AHINT"), b_hint = hint_of("DINO_B200_GEMM_BHINT");
    GemmParams p = p_in;
    if (!p.a_hint) p.a_hint = a_hint;
    if (!p.b_hint) p.b_hint = b_hint;
    const int BN = plan.BN;
#ifdef DINO_B200_EXPERIMENTAL
    if (plan.MC == 2) {
        if (epi == EPI_BIAS_F16) return launch_gemm_t<256, EPI_BIAS_F16, 2, 2>(tmA, tmB, tmC, p, plan.sms, st);
        if (epi == EPI_GELU_F16) return launch_gemm_t<256, EPI_GELU_F16, 2, 2>(tmA, tmB, tmC, p, plan.sms, st);
        if (epi == EPI_RESID_F32) return launch_gemm_t<256, EPI_RESID_F32, 2, 2>(tmA, tmB, tmC, p, plan.sms, st);
        if (epi == EPI_SWIGLU_F16) return launch_gemm_t<256, EPI_SWIGLU_F16, 2, 2>(tmA, tmB, tmC, p, plan.sms, st);
        throw StatusError(DINO_B200_ERR_UNSUPPORTED, "gemm: no multicast kernel for this epilogue");
    }
#endif
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) throw StatusError(DINO_B200_ERR_INVALID, "gemm: empty problem");
    if (p.N % 8) throw StatusError(DINO_B200_ERR_UNSUPPORTED, "gemm: N must be a multiple of 8");
#define DINO_GEMM_CASE(bn, e) \
    if (BN == bn && epi == e) return plan.CG == 2 ? launch_gemm_t<bn, e, 2>(tmA, tmB, tmC, p, plan.sms, st) : launch_gemm_t<bn, e, 1>(tmA, tmB, tmC, p, plan.sms, st)
    DINO_GEMM_CASE(256, EPI_BIAS_F16);
    DINO_GEMM_CASE(128, EPI_BIAS_F16);
    DINO_GEMM_CASE(256, EPI_GELU_F16);
    DINO_GEMM_CASE(128, EPI_GELU_F16);
    DINO_GEMM_CASE(256, EPI_RESID_F32);
    DINO_GEMM_CASE(128, EPI_RESID_F32);
    DINO_GEMM_CASE(256, EPI_RESID_LN_F32);
    DINO_GEMM_CASE(128, EPI_RESID_LN_F32);
    DINO_GEMM_CASE(256, EPI_SWIGLU_F16);
    DINO_GEMM_CASE(256, EPI_PATCH_F32);
    DINO_GEMM_CASE(128, EPI_PATCH_F32);
#undef DINO_GEMM_CASE
    throw StatusError(DINO_B200_ERR_UNSUPPORTED, "gemm: no kernel for this (tile, epilogue) pair");
}

// The product library carries ONE attention kernel (v10: v8's tile loop run as one continuous stream across work items, one MMA
// warp per query tile, deferred TMA-store epilogue, a third of the exponentials on the FMA pipe).  Per ViT-L layer at batch 64 on
// B200: v3 932 us, v4 885, v5 780, v6 (16 softmax warps) 872, v7 878, v8 724, v9 (loop rotated by half a chunk) 817, v10 666-680.
// v3 / v5 / v7 / v8 are only in -DDINO_B200_EXPERIMENTAL builds (DINO_B200_ATTN=3|5|7|8, tools/build_variant.sh); v1, v2, v4, v6
// and v9 were removed from the tree after measurement (git history).
static int attention_variant() {
    static int v = [] {
#ifdef DINO_B200_EXPERIMENTAL
        const char *e = getenv("DINO_B200_ATTN");
        if (e && (e[0] == '3' || e[0] == '5' || e[0] == '7' || e[0] == '8')) return e[0] - '0';
#endif
        return 10;
    }();
    return v;
}

static unsigned long long *attention_trace_buffer(cudaStream_t st) {
    static unsigned long long *tr = nullptr;
    if (!tr) cudaMalloc(&tr, 3 * 512 * 2 * 8);
    cudaMemsetAsync(tr, 0, 3 * 512 * 2 * 8, st);
    if (const char *f = getenv("DINO_B200_TRACE_PTR")) {
        FILE *fp = fopen(f, "w");
        if (fp) {
            fprintf(fp, "%llu\n", (unsigned long long) tr);
            fclose(fp);
        }
    }
    return tr;
}

template <typename Params> static Params attention_params(__half *out, int B, int n_tok, int D) {
    // (n_phantom, where a kernel has it, is value-initialised to 0 = exact attention)
    Params ap{};
    ap.n_tok = n_tok;
    ap.hidden = D;
    ap.n_heads = D / ATT_HD;
    ap.n_qblk = (n_tok + 255) / 256;
    ap.num_items = B * ap.n_heads * ap.n_qblk;
    ap.out = out;
    ap.scale_log2 = (1.0f / sqrtf(static_cast<float>(ATT_HD))) * 1.4426950408889634f;
    ap.trace = nullptr;
    return ap;
}

static void launch_attention(const CUtensorMap &tmQKV, __half *out, int B, int n_tok, int D, int num_sms, cudaStream_t st, bool fa_compat = false,
                             int reverse = 0) {
    const int variant = attention_variant();
    if (variant == 10) {
        Attn10Params ap = attention_params<Attn10Params>(out, B, n_tok, D);
        ap.n_phantom = fa_compat ? (32 - n_tok % 32) % 32 : 0;        // GGML_PAD(tokens, 32) - tokens (dinov2.cpp:499-500)
        ap.reverse = reverse;
#if defined(AT10_TRACE) || defined(AT10_PROF)
        ap.trace = attention_trace_buffer(st);
#endif
        const CUtensorMap tmOut = make_tmap_3d_f16(out, D, n_tok, B, D, ATT_BKV);
        const int grid = std::max(1, std::min(ap.num_items, num_sms));
        launch_k(attention_fwd_v10, dim3(grid), dim3(AT10_THREADS), AT10_SMEM_BYTES, st, tmQKV, tmOut, ap);
    }
#ifdef DINO_B200_EXPERIMENTAL
    else if (variant == 8) {
        Attn8Params ap = attention_params<Attn8Params>(out, B, n_tok, D);
#ifdef AT8_TRACE
        ap.trace = attention_trace_buffer(st);
#endif
        const int grid = std::max(1, std::min(ap.num_items, num_sms));
        attention_fwd_v8<<<grid, AT8_THREADS, AT8_SMEM_BYTES, st>>>(tmQKV, ap);
    } else if (variant == 7) {
        Attn7Params ap = attention_params<Attn7Params>(out, B, n_tok, D);
#ifdef AT7_TRACE
        ap.trace = attention_trace_buffer(st);
#endif
        const int grid = std::max(1, std::min(ap.num_items, num_sms));
        attention_fwd_v7<<<grid, AT7_THREADS, AT7_SMEM_BYTES, st>>>(tmQKV, ap);
    } else if (variant == 5) {
        Attn5Params ap = attention_params<Attn5Params>(out, B, n_tok, D);
#ifdef AT5_TRACE
        ap.trace = attention_trace_buffer(st);
#endif
        const int grid = std::max(1, std::min(ap.num_items, num_sms));
        attention_fwd_v5<<<grid, AT5_THREADS, AT5_SMEM_BYTES, st>>>(tmQKV, ap);
    } else {
        Attn3Params ap = attention_params<Attn3Params>(out, B, n_tok, D);
        static const int pingpong = [] { const char *e = getenv("DINO_B200_ATTN_PINGPONG"); return (e && e[0] == '1') ? 1 : 0; }();
        ap.pingpong = pingpong;
#ifdef AT3_TRACE
        ap.trace = attention_trace_buffer(st);
#endif
        static const int grid_mult = [] { const char *e = getenv("DINO_B200_ATTN_GRID"); return e ? atoi(e) : 1; }();
        const int grid = grid_mult <= 0 ? ap.num_items : std::max(1, std::min(ap.num_items, num_sms * grid_mult));
        attention_fwd_v3<<<grid, AT3_THREADS, AT3_SMEM_BYTES, st>>>(tmQKV, ap);
    }
#endif
    (void) attention_trace_buffer;
    DINO_CUDA(cudaGetLastError());
}

static void launch_layernorm(const float *X, const float *g, const float *b, void *out, int rows, int D, float eps, bool half_out,
                             cudaStream_t st, int reverse = 0) {
    if (D % 4 || D > 128 * LN_MAX_V4) throw StatusError(DINO_B200_ERR_UNSUPPORTED, "layernorm: hidden size not supported");
    const int grid = (rows + 7) / 8;
    const int nv4 = (D + 127) / 128;          // float4 per lane; the row buffer is sized for the model width
#define DINO_LN_CASE(n)                                                                                       \
    if (nv4 <= n) {                                                                                            \
        if (half_out) launch_k(layernorm_kernel<true, n>, dim3(grid), dim3(256), 0, st, X, g, b, out, rows, D, eps, reverse);   \
        else launch_k(layernorm_kernel<false, n>, dim3(grid), dim3(256), 0, st, X, g, b, out, rows, D, eps, reverse);           \
        DINO_CUDA(cudaGetLastError());                                                                         \
        return;                                                                                                \
    }
    DINO_LN_CASE(3)      // D <= 384  (ViT-S)
    DINO_LN_CASE(6)      // D <= 768  (ViT-B)
    DINO_LN_CASE(8)      // D <= 1024 (ViT-L)
    DINO_LN_CASE(12)     // D <= 1536 (ViT-g)
#undef DINO_LN_CASE
    DINO_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ engine
struct Linear {
    __half *w = nullptr;   // [N, ldw]
    float *bias = nullptr; // [N]
    int N = 0, K = 0, ldw = 0, BN = 0;
    CUtensorMap tm64, tm128, tm256;    // weight tiles of 64 / 128 / 256 rows (a CTA loads BN / CG rows per k-block)
    const CUtensorMap &tm(GemmPlan p) const {
        const int rows = p.BN / p.CG / p.MC;
        return rows == 64 ? tm64 : rows == 128 ? tm128 : tm256;
    }
};

struct Layer {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ls1, *ls2;
    Linear qkv, proj, fc1, fc2;
};

struct ProfileEvent {
    cudaEvent_t a, b;
    int kind;   // 0 gemm, 1 attention, 2 other
};

}  // namespace dino

using namespace dino;

// One captured forward pass: every pointer and shape a replay depends on is part of the key
struct GraphKey {
    const void *images;
    int layout, B, H, W, flags;
    const void *cls, *patch, *logits, *probs;
    bool operator<(const GraphKey &o) const {
        return std::tie(images, layout, B, H, W, flags, cls, patch, logits, probs) <
               std::tie(o.images, o.layout, o.B, o.H, o.W, o.flags, o.cls, o.patch, o.logits, o.probs);
    }
};
struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    uint64_t launches = 0;     // kernels per replay
};

struct dino_b200_engine {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    // CUDA-graph cache of whole forward passes (reference: dino_predict rebuilds and re-plans its ggml graph on every call,
    // dinov2.cpp:907-942).  Dropped whenever a baked pointer may change (arena growth, new pos-embed table).
    bool use_graphs = true;
    std::map<GraphKey, GraphEntry> graphs;
    uint64_t graph_replays = 0, graph_captures = 0;
    dino_b200_hparams hp{};
    bool swiglu = false;
    int mlp_in = 0, mlp_hidden = 0;
    std::vector<std::string> labels;

    std::vector<void *> allocs;
    Linear patch;
    float *cls = nullptr, *pos = nullptr, *reg = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
    __half *wc = nullptr;
    float *bc = nullptr;
    std::vector<Layer> layers;
    std::map<std::pair<int, int>, float *> pos_cache;

    // activation arena (grown on demand)
    size_t cap_tok = 0, cap_patch = 0, cap_img = 0, cap_batch = 0;
    float *d_img = nullptr, *X = nullptr, *Y = nullptr, *feat = nullptr, *logits = nullptr, *probs = nullptr;
    __half *Ape = nullptr, *Xn = nullptr, *QKV = nullptr, *AO = nullptr, *H1 = nullptr;
    // host-API output staging in device memory
    float *o_cls = nullptr, *o_patch = nullptr;
    size_t cap_o_patch = 0;
    // pipelined host interface (dino_b200_submit / dino_b200_wait): two input slots, uploads on their own stream
    cudaStream_t copy_stream = nullptr;
    float *d_in[2] = {nullptr, nullptr};
    size_t cap_in[2] = {0, 0};
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_fwd[2] = {nullptr, nullptr};
    cudaStream_t d2h_stream = nullptr;                  // read-backs run under the NEXT batch's forward pass
    float *r_buf[2] = {nullptr, nullptr};               // per-slot device results: cls | logits | probs | patch tokens
    size_t cap_r[2] = {0, 0};
    uint64_t n_submitted = 0, n_waited = 0;
    // PCA colouring scratch (dino_b200_pca_rgb*): mean [B][D], V / W [B][D][3], Y [B][NP][3], staging for the host variant
    float *pca_mean = nullptr, *pca_v = nullptr, *pca_w = nullptr, *pca_y = nullptr, *pca_x = nullptr;
    uint8_t *pca_rgb = nullptr;
    size_t pca_cap_b = 0, pca_cap_np = 0, pca_cap_x = 0;
    int *ln_count = nullptr;          // per-128-row-block tile counters of the fused residual-GEMM + LayerNorm epilogue
    uint8_t *d_u8 = nullptr;          // raw frames for on-device preprocessing
    size_t cap_u8 = 0;

    // feature all-gather (dino_b200_gather_*): this engine is rank `g_rank` of `g_world`; g_peer[r] = rank r's gather buffer
    // ([g_world * g_max_batch][g_rpi][D] fp32) as addressable from this device (own allocation, same-process peer, or CUDA IPC)
    int g_rank = 0, g_world = 0, g_what = 0, g_max_batch = 0, g_rpi = 0, g_H = 0, g_W = 0;
    float *g_buf = nullptr;
    float *g_peer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool g_ipc[8] = {false, false, false, false, false, false, false, false};
    // raw-frame slots of the pipelined uint8 interface (dino_b200_submit_u8) and their PCA colour results
    uint8_t *d_u8s[2] = {nullptr, nullptr};
    size_t cap_u8s[2] = {0, 0};
    uint8_t *r_rgb[2] = {nullptr, nullptr};
    size_t cap_rgb[2] = {0, 0};

    uint64_t launches = 0;
    bool profiling = false;
    std::vector<ProfileEvent> prof;
    std::vector<ProfileEvent> prof_pool;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    std::string err;

    void *dmalloc(size_t bytes) {
        void *p = nullptr;
        DINO_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
        allocs.push_back(p);
        return p;
    }
};

namespace dino {

struct TensorTable {
    std::map<std::string, const dino_b200_tensor *> by_name;
    const dino_b200_tensor &at(const std::string &n) const {
        auto it = by_name.find(n);
        if (it == by_name.end()) throw StatusError(DINO_B200_ERR_FORMAT, "checkpoint is missing tensor '" + n + "'");
        return *it->second;
    }
    bool has(const std::string &n) const { return by_name.count(n) != 0; }
};

// Every host->device upload is ordered on the engine's (non-blocking) stream: a plain cudaMemcpy from pageable
// memory returns once the data is staged, possibly before the DMA lands, and the legacy default stream does not
// synchronise with non-blocking streams — a conversion kernel launched right after could read stale bytes.
static void h2d(dino_b200_engine *e, void *dst, const void *src, size_t bytes) {
    DINO_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e->stream));
}

// Every dimension of a checkpoint tensor must be positive and the element count must fit comfortably in 63 bits;
// callers divide by ne[0] and size buffers from these numbers.
static int64_t numel(const dino_b200_tensor &t) {
    if (t.n_dims < 1 || t.n_dims > 4) throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + (t.name ? t.name : "?") + "' has a bad rank");
    int64_t n = 1;
    for (int d = 0; d < t.n_dims; ++d) {
        if (t.ne[d] <= 0 || t.ne[d] > (int64_t(1) << 40) || n > (int64_t(1) << 62) / t.ne[d])
            throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + (t.name ? t.name : "?") + "' has a bad dimension");
        n *= t.ne[d];
    }
    return n;
}

// Scratch device allocation that is released on every exit path (conversion temporaries of the loader)
struct DeviceTemp {
    void *p = nullptr;
    explicit DeviceTemp(size_t bytes) { DINO_CUDA(cudaMalloc(&p, std::max<size_t>(bytes, 16))); }
    ~DeviceTemp() { if (p) cudaFree(p); }
    DeviceTemp(const DeviceTemp &) = delete;
    DeviceTemp &operator=(const DeviceTemp &) = delete;
    template <typename T> T *as() const { return static_cast<T *>(p); }
};

static float *upload_f32(dino_b200_engine *e, const dino_b200_tensor &t, int64_t expect, const std::vector<int> *perm = nullptr) {
    if (t.type != DINO_B200_TYPE_F32) throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + t.name + "' must be F32");
    if (numel(t) != expect)
        throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + t.name + "' has " + std::to_string(numel(t)) +
                                                    " elements, expected " + std::to_string(expect));
    if (!t.data) throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + t.name + "' has no data");
    float *d = static_cast<float *>(e->dmalloc(expect * sizeof(float)));
    if (perm) {
        std::vector<float> tmp(expect);
        const float *src = static_cast<const float *>(t.data);
        for (int64_t i = 0; i < expect; ++i) tmp[(*perm)[i]] = src[i];
        h2d(e, d, tmp.data(), expect * sizeof(float));
    } else {
        h2d(e, d, t.data, expect * sizeof(float));
    }
    return d;
}

// Weight matrix [N, K] (ggml ne = [K, N], or [14,14,3,N] for the patch projection) -> device fp16 [N, ldw].
static void upload_linear(dino_b200_engine *e, Linear &L, const dino_b200_tensor &w, const dino_b200_tensor &b, int N, int K,
                          int epi, const std::vector<int> *perm, bool want_tmap = true) {
    const int64_t n_w = numel(w);                    // validates the dimensions (all positive)
    int64_t k_file = w.ne[0];
    if (w.n_dims == 4) k_file = w.ne[0] * w.ne[1] * w.ne[2];
    const int64_t n_file = n_w / k_file;
    if (k_file != K || n_file != N)
        throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + w.name + "' is [" + std::to_string(n_file) + ", " +
                                                    std::to_string(k_file) + "], expected [" + std::to_string(N) + ", " +
                                                    std::to_string(K) + "]");
    if (!w.data) throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + w.name + "' has no data");
    L.N = N;
    L.K = K;
    L.ldw = (K + 63) / 64 * 64;
    L.BN = pick_bn(epi, N);
    L.w = static_cast<__half *>(e->dmalloc(static_cast<size_t>(N) * L.ldw * sizeof(__half)));
    std::unique_ptr<DeviceTemp> d_perm_buf;
    int *d_perm = nullptr;
    if (perm) {
        d_perm_buf.reset(new DeviceTemp(N * sizeof(int)));
        d_perm = d_perm_buf->as<int>();
        h2d(e, d_perm, perm->data(), N * sizeof(int));
    }
    const int grid = e->num_sms * 8;
    if (w.type == DINO_B200_TYPE_F16) {
        if (w.nbytes && w.nbytes < static_cast<uint64_t>(N) * K * sizeof(__half)) throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + w.name + "' is shorter than its shape");
        if (!perm && L.ldw == K) {
            h2d(e, L.w, w.data, static_cast<size_t>(N) * K * sizeof(__half));
        } else {
            DeviceTemp tmp(static_cast<size_t>(N) * K * sizeof(__half));
            h2d(e, tmp.p, w.data, static_cast<size_t>(N) * K * sizeof(__half));
            copy_rows_f16_kernel<<<grid, 256, 0, e->stream>>>(tmp.as<__half>(), L.w, N, K, L.ldw, d_perm);
            DINO_CUDA(cudaGetLastError());
            DINO_CUDA(cudaStreamSynchronize(e->stream));
        }
    } else if (w.type == DINO_B200_TYPE_Q8_0 || w.type == DINO_B200_TYPE_Q4_0 || w.type == DINO_B200_TYPE_Q4_1 ||
               w.type == DINO_B200_TYPE_Q5_0 || w.type == DINO_B200_TYPE_Q5_1) {
        if (K % 32 || L.ldw != K) throw StatusError(DINO_B200_ERR_FORMAT, std::string("quantised tensor '") + w.name + "' has an unsupported row length");
        static const int kBlockBytes[9] = {0, 0, 18, 20, 0, 0, 22, 24, 34};
        const long long nblk = static_cast<long long>(N) * (K / 32);
        const uint64_t need = static_cast<uint64_t>(nblk) * kBlockBytes[w.type];
        if (w.nbytes < need) throw StatusError(DINO_B200_ERR_FORMAT, std::string("quantised tensor '") + w.name + "' is shorter than its shape");
        DeviceTemp raw(need);
        h2d(e, raw.p, w.data, need);
        const uint8_t *rp = raw.as<uint8_t>();
        switch (w.type) {
            case DINO_B200_TYPE_Q4_0: dequant_kernel<2><<<grid, 256, 0, e->stream>>>(rp, L.w, nblk, K / 32, L.ldw, d_perm); break;
            case DINO_B200_TYPE_Q4_1: dequant_kernel<3><<<grid, 256, 0, e->stream>>>(rp, L.w, nblk, K / 32, L.ldw, d_perm); break;
            case DINO_B200_TYPE_Q5_0: dequant_kernel<6><<<grid, 256, 0, e->stream>>>(rp, L.w, nblk, K / 32, L.ldw, d_perm); break;
            case DINO_B200_TYPE_Q5_1: dequant_kernel<7><<<grid, 256, 0, e->stream>>>(rp, L.w, nblk, K / 32, L.ldw, d_perm); break;
            default: dequant_kernel<8><<<grid, 256, 0, e->stream>>>(rp, L.w, nblk, K / 32, L.ldw, d_perm); break;
        }
        DINO_CUDA(cudaGetLastError());
        DINO_CUDA(cudaStreamSynchronize(e->stream));
    } else if (w.type == DINO_B200_TYPE_F32) {
        std::vector<__half> h(static_cast<size_t>(N) * L.ldw, __float2half(0.f));
        const float *src = static_cast<const float *>(w.data);
        for (int n = 0; n < N; ++n) {
            const int on = perm ? (*perm)[n] : n;
            for (int k = 0; k < K; ++k) h[static_cast<size_t>(on) * L.ldw + k] = __float2half_rn(src[static_cast<size_t>(n) * K + k]);
        }
        h2d(e, L.w, h.data(), h.size() * sizeof(__half));
        DINO_CUDA(cudaStreamSynchronize(e->stream));
    } else {
        throw StatusError(DINO_B200_ERR_FORMAT, std::string("tensor '") + w.name + "' has unsupported type " + std::to_string(w.type));
    }
    DINO_CUDA(cudaStreamSynchronize(e->stream));     // the temporaries above are released only after their last use
    d_perm_buf.reset();
    L.bias = upload_f32(e, b, N, perm);
    // logical K columns = ldw: the pad columns are real zeros, so the K loop needs no tail case
    if (want_tmap) {
        L.tm64 = make_tmap_f16(L.w, L.ldw, N, L.ldw, 64);
        L.tm128 = make_tmap_f16(L.w, L.ldw, N, L.ldw, 128);
        L.tm256 = make_tmap_f16(L.w, L.ldw, N, L.ldw, N >= 256 ? 256 : 128);   // 256-row tiles only exist for N % 256 == 0
    }
}

static void build_engine(dino_b200_engine *e, const dino_b200_model_desc *desc) {
    const dino_b200_hparams &hp = desc->hparams;
    e->hp = hp;
    if (e->hp.eps <= 0.f) e->hp.eps = 1e-6f;   // dino_hparams::eps default (dinov2.h:33)
    const int D = hp.hidden_size, Lyr = hp.num_hidden_layers, H = hp.num_attention_heads, R = hp.num_register_tokens;
    if (D <= 0 || Lyr <= 0 || H <= 0 || hp.patch_size == 0 || hp.img_size < hp.patch_size)
        throw StatusError(DINO_B200_ERR_INVALID, "invalid hyper-parameters");
    if (D % H || D / H != ATT_HD)
        throw StatusError(DINO_B200_ERR_UNSUPPORTED, "attention kernel requires head_dim == 64 (every DINOv2 size has it)");
    if (D % 64 || D > 128 * LN_MAX_V4) throw StatusError(DINO_B200_ERR_UNSUPPORTED, "hidden size must be a multiple of 64 and <= 1536");
    if (hp.patch_size != 14) throw StatusError(DINO_B200_ERR_UNSUPPORTED, "patch size must be 14");

    TensorTable tt;
    for (int i = 0; i < desc->n_tensors; ++i) tt.by_name[desc->tensors[i].name] = &desc->tensors[i];

    const int grid_m = hp.img_size / hp.patch_size;
    e->cls = upload_f32(e, tt.at("embeddings.cls_token"), D);
    e->pos = upload_f32(e, tt.at("embeddings.position_embeddings"), static_cast<int64_t>(1 + grid_m * grid_m) * D);
    if (R > 0) e->reg = upload_f32(e, tt.at("embeddings.register_tokens"), static_cast<int64_t>(R) * D);
    upload_linear(e, e->patch, tt.at("embeddings.patch_embeddings.projection.weight"),
                  tt.at("embeddings.patch_embeddings.projection.bias"), D, 3 * hp.patch_size * hp.patch_size, EPI_PATCH_F32, nullptr);

    e->swiglu = (Lyr == 40);   // the reference's own switch (dinov2.cpp:740)
    e->layers.resize(Lyr);
    for (int l = 0; l < Lyr; ++l) {
        const std::string b = "encoder.layer." + std::to_string(l) + ".";
        Layer &ly = e->layers[l];
        ly.ln1_g = upload_f32(e, tt.at(b + "norm1.weight"), D);
        ly.ln1_b = upload_f32(e, tt.at(b + "norm1.bias"), D);
        ly.ln2_g = upload_f32(e, tt.at(b + "norm2.weight"), D);
        ly.ln2_b = upload_f32(e, tt.at(b + "norm2.bias"), D);
        ly.ls1 = upload_f32(e, tt.at(b + "layer_scale1.lambda1"), D);
        ly.ls2 = upload_f32(e, tt.at(b + "layer_scale2.lambda1"), D);
        upload_linear(e, ly.qkv, tt.at(b + "attention.attention.qkv.weight"), tt.at(b + "attention.attention.qkv.bias"), 3 * D, D,
                      EPI_BIAS_F16, nullptr);
        upload_linear(e, ly.proj, tt.at(b + "attention.output.dense.weight"), tt.at(b + "attention.output.dense.bias"), D, D,
                      EPI_RESID_F32, nullptr);
        if (e->swiglu) {
            const dino_b200_tensor &win = tt.at(b + "mlp.weights_in.weight");
            if (win.n_dims < 2 || win.ne[1] <= 0 || win.ne[1] > (1 << 20)) throw StatusError(DINO_B200_ERR_FORMAT, "mlp.weights_in.weight has a bad shape");
            const int n_in = static_cast<int>(win.ne[1]);
            if (n_in % 256) throw StatusError(DINO_B200_ERR_UNSUPPORTED, "SwiGLU hidden size must be a multiple of 128");
            const int hid = n_in / 2;
            e->mlp_in = n_in;
            e->mlp_hidden = hid;
            // interleave per 256-row tile: rows [256t, 256t+128) = gate[128t..], rows [256t+128, 256t+256) = up[128t..]
            std::vector<int> perm(n_in);
            for (int j = 0; j < hid; ++j) {
                perm[j] = (j / 128) * 256 + (j % 128);
                perm[hid + j] = (j / 128) * 256 + 128 + (j % 128);
            }
            upload_linear(e, ly.fc1, win, tt.at(b + "mlp.weights_in.bias"), n_in, D, EPI_SWIGLU_F16, &perm);
            upload_linear(e, ly.fc2, tt.at(b + "mlp.weights_out.weight"), tt.at(b + "mlp.weights_out.bias"), D, hid, EPI_RESID_F32, nullptr);
        } else {
            const dino_b200_tensor &w1 = tt.at(b + "mlp.fc1.weight");
            if (w1.n_dims < 2 || w1.ne[1] <= 0 || w1.ne[1] > (1 << 20)) throw StatusError(DINO_B200_ERR_FORMAT, "mlp.fc1.weight has a bad shape");
            const int n_in = static_cast<int>(w1.ne[1]);
            e->mlp_in = n_in;
            e->mlp_hidden = n_in;
            upload_linear(e, ly.fc1, w1, tt.at(b + "mlp.fc1.bias"), n_in, D, EPI_GELU_F16, nullptr);
            upload_linear(e, ly.fc2, tt.at(b + "mlp.fc2.weight"), tt.at(b + "mlp.fc2.bias"), D, n_in, EPI_RESID_F32, nullptr);
        }
    }
    e->lnf_g = upload_f32(e, tt.at("layernorm.weight"), D);
    e->lnf_b = upload_f32(e, tt.at("layernorm.bias"), D);
    if (hp.num_classes > 0 && tt.has("classifier.weight")) {
        Linear c;   // reuse the conversion path (q8_0 / f16 -> fp16 [C, 2D]); the head uses a GEMV kernel, not TMA tiles
        upload_linear(e, c, tt.at("classifier.weight"), tt.at("classifier.bias"), hp.num_classes, 2 * D, EPI_BIAS_F16, nullptr, false);
        e->wc = c.w;
        e->bc = c.bias;
    }
    DINO_CUDA(cudaStreamSynchronize(e->stream));
}

// ------------------------------------------------------------------------------------------------ arena
static void drop_graphs(dino_b200_engine *e);
static void gather_release(dino_b200_engine *e);
static void free_arena(dino_b200_engine *e) {
    drop_graphs(e);                                         // captured forwards hold arena pointers
    void *ptrs[] = {e->d_img, e->X, e->Y, e->feat, e->logits, e->probs, e->Ape, e->Xn, e->QKV, e->AO, e->H1, e->o_cls, e->o_patch, e->ln_count};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    e->d_img = e->X = e->Y = e->feat = e->logits = e->probs = e->o_cls = e->o_patch = nullptr;
    e->Ape = e->Xn = e->QKV = e->AO = e->H1 = nullptr;
    e->ln_count = nullptr;
    e->cap_tok = e->cap_patch = e->cap_img = e->cap_batch = e->cap_o_patch = 0;
}

template <typename T> static void arena_alloc(T *&p, size_t elems, cudaStream_t st) {
    DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&p), std::max<size_t>(elems * sizeof(T), 16)));
    DINO_CUDA(cudaMemsetAsync(p, 0, std::max<size_t>(elems * sizeof(T), 16), st));
}

static void ensure_arena(dino_b200_engine *e, int B, int H, int W) {
    const int ps = e->hp.patch_size, D = e->hp.hidden_size, R = e->hp.num_register_tokens;
    const size_t np = static_cast<size_t>(H / ps) * (W / ps);
    const size_t tok = static_cast<size_t>(B) * (1 + R + np), patch = static_cast<size_t>(B) * np;
    const size_t img = static_cast<size_t>(B) * 3 * H * W;
    if (tok <= e->cap_tok && patch <= e->cap_patch && img <= e->cap_img && static_cast<size_t>(B) <= e->cap_batch) return;
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    const size_t ntok = std::max(tok, e->cap_tok), npatch = std::max(patch, e->cap_patch), nimg = std::max(img, e->cap_img);
    const size_t nb = std::max<size_t>(B, e->cap_batch);
    free_arena(e);
    // +128 rows of slack: attention / GEMM boxes may read (never write) past the last token row
    const size_t rows = ntok + 128;
    arena_alloc(e->d_img, nimg, e->stream);
    arena_alloc(e->Ape, (npatch + 128) * e->patch.ldw, e->stream);
    arena_alloc(e->X, rows * D, e->stream);
    arena_alloc(e->Y, rows * D, e->stream);
    arena_alloc(e->Xn, rows * D, e->stream);
    arena_alloc(e->QKV, rows * 3 * D, e->stream);
    arena_alloc(e->AO, rows * D, e->stream);
    arena_alloc(e->H1, rows * e->mlp_hidden, e->stream);
    arena_alloc(e->feat, nb * 2 * D, e->stream);
    arena_alloc(e->logits, nb * std::max<size_t>(e->hp.num_classes, 1), e->stream);
    arena_alloc(e->probs, nb * std::max<size_t>(e->hp.num_classes, 1), e->stream);
    arena_alloc(e->o_cls, nb * D, e->stream);
    arena_alloc(e->ln_count, 2 * (rows / GEMM_BM + 2) + 4, e->stream);   // block counters, slice counters, ticket: zeroed here; every launch leaves them zero again
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    e->cap_tok = ntok;
    e->cap_patch = npatch;
    e->cap_img = nimg;
    e->cap_batch = nb;
}

static float *pos_for_grid(dino_b200_engine *e, int gh, int gw) {
    const int M = e->hp.img_size / e->hp.patch_size, D = e->hp.hidden_size;
    auto it = e->pos_cache.find({gh, gw});
    if (it != e->pos_cache.end()) return it->second;
    if (gh * gw == M * M) return e->pos;   // the reference's early return keys on the patch COUNT (dinov2.cpp:176-179)
    float *buf = static_cast<float *>(e->dmalloc(static_cast<size_t>(1 + gh * gw) * D * sizeof(float)));
    pos_embed_bicubic_kernel<<<1 + gh * gw, 128, 0, e->stream>>>(e->pos, buf, M, gh, gw, D);
    DINO_CUDA(cudaGetLastError());
    e->launches++;
    e->pos_cache[{gh, gw}] = buf;
    return buf;
}

// ------------------------------------------------------------------------------------------------ forward
struct Prof {
    dino_b200_engine *e;
    cudaStream_t st;
    void begin(int kind) {
        if (!e->profiling) return;
        ProfileEvent pe;
        if (!e->prof_pool.empty()) {
            pe = e->prof_pool.back();
            e->prof_pool.pop_back();
        } else {
            DINO_CUDA(cudaEventCreate(&pe.a));
            DINO_CUDA(cudaEventCreate(&pe.b));
        }
        pe.kind = kind;
        DINO_CUDA(cudaEventRecord(pe.a, st));
        e->prof.push_back(pe);
    }
    void end() {
        if (!e->profiling) return;
        DINO_CUDA(cudaEventRecord(e->prof.back().b, st));
    }
};

constexpr int kFlagGather = 0x100;   // internal forward flag: run the fused final-LayerNorm + all-gather kernel (engine gather state)

static void drop_graphs(dino_b200_engine *e) {
    for (auto &kv : e->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    e->graphs.clear();
}

// Enqueues one forward pass on `st` (kernel launches and device-to-device copies only: capturable).  Returns the number of
// kernels launched.  `pos` is the position-embedding table for this grid (pos_for_grid, resolved by the caller).
static uint64_t forward_enqueue(dino_b200_engine *e, const float *images, int layout, int B, int H, int W, int flags, float *cls,
                                float *patch, float *logits, float *probs, const float *pos, cudaStream_t st) {
    const dino_b200_hparams &hp = e->hp;
    const int ps = hp.patch_size, D = hp.hidden_size, R = hp.num_register_tokens;
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    const int g_num_sms = e->num_sms;
    const int gh = H / ps, gw = W / ps, np = gh * gw, ntok = 1 + R + np;
    const int M = B * ntok, Mp = B * np;
    uint64_t nl = 0;
    Prof prof{e, st};
    // Programmatic dependent launch for every kernel of this pass — only for SMALL workloads, where kernels last ~10 us and the
    // launch / prologue latency is a tenth of the pass (ViT-L batch 1: 2.48 -> 2.36 ms).  At batch 64 it measured 1.7 % SLOWER
    // (903 vs 918 images/s, interleaved A/B x3): dependents that are scheduled early hold SM slots next to the stragglers of
    // the primary.  Never while per-kernel events are being recorded.
    struct PdlScope {
        explicit PdlScope(bool on) { t_pdl = on; }
        ~PdlScope() { t_pdl = false; }
    } pdl_scope(pdl_enabled() && !e->profiling && M <= pdl_max_rows());
    // Consecutive kernels walk their row blocks in opposite directions: what a kernel wrote last (up to ~100 MB still in
    // the 126 MB L2) is what its consumer reads first.  DINO_B200_ZIGZAG=0 switches the alternation off (A/B measurements).
    static const bool zigzag = [] { const char *v = getenv("DINO_B200_ZIGZAG"); return !(v && v[0] == '0'); }();
    int dir_state = 0;
    auto next_dir = [&]() -> int {
        const int r = dir_state;
        if (zigzag) dir_state ^= 1;
        return r;
    };

    const CUtensorMap tm_ape = make_tmap_f16(e->Ape, e->patch.ldw, Mp, e->patch.ldw, GEMM_BM);
    const CUtensorMap tm_xn = make_tmap_f16(e->Xn, D, M, D, GEMM_BM);
    const CUtensorMap tm_ao = make_tmap_f16(e->AO, D, M, D, GEMM_BM);
    const CUtensorMap tm_h1 = make_tmap_f16(e->H1, e->mlp_hidden, M, e->mlp_hidden, GEMM_BM);
    const CUtensorMap tm_qkv = make_tmap_2d(e->QKV, false, 3 * D, M, 3 * D, ATT_BKV, attn_promotion());
    const CUtensorMap tmo_qkv = make_tmap_out(EPI_BIAS_F16, e->QKV, 3 * D, M, 3 * D);
    const CUtensorMap tmo_h1 = make_tmap_out(EPI_GELU_F16, e->H1, e->mlp_hidden, M, e->mlp_hidden);
    const CUtensorMap tmo_x = make_tmap_out(EPI_RESID_F32, e->X, D, M, D);

    // 1. patch embedding: im2col -> GEMM (+bias +pos, scattered to token rows) ; cls/register rows
    prof.begin(2);
    {
        const long long total = static_cast<long long>(B) * (gh * ps) * (gw * ps);   // one thread per covered pixel
        const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(g_num_sms) * 16));
        launch_k(im2col_patch14_kernel, dim3(grid), dim3(256), 0, st, images, e->Ape, B, H, W, ps, gh, gw, e->patch.ldw, layout);
        DINO_CUDA(cudaGetLastError());
        nl++;
    }
    prof.end();
    {
        GemmParams gp{};
        gp.M = Mp; gp.N = D; gp.K = e->patch.ldw;
        gp.bias = e->patch.bias; gp.out = e->X; gp.ldo = D;
        gp.pos = pos; gp.np = np; gp.ntok = ntok; gp.tok_off = 1 + R;
        gp.reverse = next_dir();
        prof.begin(0);
        const GemmPlan pl = plan_gemm(EPI_PATCH_F32, gp.M, gp.N, g_num_sms);
        launch_gemm(EPI_PATCH_F32, pl, tm_ape, e->patch.tm(pl), tmo_x, gp, st);
        prof.end();
        nl++;
    }
    prof.begin(2);
    launch_k(prefix_tokens_kernel, dim3(B), dim3(128), 0, st, e->X, e->cls, pos, e->reg, ntok, D, R);
    DINO_CUDA(cudaGetLastError());
    nl++;
    prof.end();

    // 2. encoder blocks
    // With DINO_B200_FUSE_LN=1 norm2 runs inside the o-proj GEMM and the next block's norm1 inside the fc2 GEMM
    // (EPI_RESID_LN_F32: LayerNorm worker warps normalise each row block out of L2 as soon as its last column tile has been
    // added; bit-identical to the stand-alone kernel, tests/test_gpu_kernels.py).  Off by default: measured equal at full
    // batch (the step is power-limited) and slower at batch 1 — DESIGN.md section 5.
    static const int fuse_ln_env = [] { const char *e = getenv("DINO_B200_FUSE_LN"); return e ? atoi(e) : -1; }();   // -1 = by size
    static const int ln_slice_env = [] { const char *e = getenv("DINO_B200_LN_SLICE"); return e ? atoi(e) : 0; }();
    const bool fuse_ln = dino::gemm_ln_width_ok(D) && (fuse_ln_env < 0 ? M <= fuse_ln_max_rows() : fuse_ln_env != 0);
    const int ln_slice = ln_slice_env > 0 ? ln_slice_env : (M <= fuse_ln_max_rows() ? 4 : 16);
    const size_t n_layers = e->layers.size();
    for (size_t li = 0; li < n_layers; ++li) {
        const Layer &ly = e->layers[li];
        if (li == 0 || !fuse_ln) {
            prof.begin(2);
            launch_layernorm(e->X, ly.ln1_g, ly.ln1_b, e->Xn, M, D, hp.eps, true, st, next_dir());
            prof.end();
            nl++;
        }
        {
            GemmParams gp{};
            gp.M = M; gp.N = 3 * D; gp.K = D; gp.bias = ly.qkv.bias; gp.out = e->QKV; gp.ldo = 3 * D;
            gp.reverse = next_dir();
            prof.begin(0);
            const GemmPlan pl = plan_gemm(EPI_BIAS_F16, gp.M, gp.N, g_num_sms);
            launch_gemm(EPI_BIAS_F16, pl, tm_xn, ly.qkv.tm(pl), tmo_qkv, gp, st);
            prof.end();
        }
        prof.begin(1);
        launch_attention(tm_qkv, e->AO, B, ntok, D, g_num_sms, st, (flags & DINO_B200_FLASH_ATTN_COMPAT) != 0, next_dir());
        prof.end();
        {
            GemmParams gp{};
            gp.M = M; gp.N = D; gp.K = D; gp.bias = ly.proj.bias; gp.lscale = ly.ls1; gp.out = e->X; gp.ldo = D;
            gp.reverse = next_dir();
            gp.ln_gamma = ly.ln2_g; gp.ln_beta = ly.ln2_b; gp.ln_out = e->Xn; gp.ln_eps = hp.eps; gp.ln_count = e->ln_count; gp.ln_slice = ln_slice;
            prof.begin(0);
            const int epi = fuse_ln ? EPI_RESID_LN_F32 : EPI_RESID_F32;
            const GemmPlan pl = plan_gemm(epi, gp.M, gp.N, g_num_sms);
            launch_gemm(epi, pl, tm_ao, ly.proj.tm(pl), tmo_x, gp, st);
            prof.end();
        }
        if (!fuse_ln) {
            prof.begin(2);
            launch_layernorm(e->X, ly.ln2_g, ly.ln2_b, e->Xn, M, D, hp.eps, true, st, next_dir());
            prof.end();
            nl++;
        }
        {
            GemmParams gp{};
            gp.M = M; gp.N = e->mlp_in; gp.K = D; gp.bias = ly.fc1.bias; gp.out = e->H1; gp.ldo = e->mlp_hidden;
            gp.reverse = next_dir();
            prof.begin(0);
            const int epi = e->swiglu ? EPI_SWIGLU_F16 : EPI_GELU_F16;
            const GemmPlan pl = plan_gemm(epi, gp.M, gp.N, g_num_sms);
            launch_gemm(epi, pl, tm_xn, ly.fc1.tm(pl), tmo_h1, gp, st);
            prof.end();
        }
        {
            GemmParams gp{};
            gp.M = M; gp.N = D; gp.K = e->mlp_hidden; gp.bias = ly.fc2.bias; gp.lscale = ly.ls2; gp.out = e->X; gp.ldo = D;
            gp.reverse = next_dir();
            const bool fuse_next = fuse_ln && li + 1 < n_layers;     // the next block's norm1; the last block feeds the final norm
            if (fuse_next) {
                const Layer &nx = e->layers[li + 1];
                gp.ln_gamma = nx.ln1_g; gp.ln_beta = nx.ln1_b; gp.ln_out = e->Xn; gp.ln_eps = hp.eps; gp.ln_count = e->ln_count; gp.ln_slice = ln_slice;
            }
            prof.begin(0);
            const int epi = fuse_next ? EPI_RESID_LN_F32 : EPI_RESID_F32;
            const GemmPlan pl = plan_gemm(epi, gp.M, gp.N, g_num_sms);
            launch_gemm(epi, pl, tm_h1, ly.fc2.tm(pl), tmo_x, gp, st);
            prof.end();
        }
        nl += 5;
    }

    // 3a. feature all-gather fused with the final LayerNorm: exported rows go straight into every rank's gather buffer
    if (flags & kFlagGather) {
        prof.begin(2);
        GatherDst dst{};
        dst.n = e->g_world;
        for (int r = 0; r < e->g_world; ++r) dst.p[r] = e->g_peer[r];
        const int rpi = e->g_rpi, tok0 = e->g_what == DINO_B200_GATHER_CLS ? 0 : 1 + R;
        const long long warps = static_cast<long long>(B) * rpi;
        const int grid = static_cast<int>((warps * 32 + 255) / 256);
        const size_t slot_row0 = static_cast<size_t>(e->g_rank) * e->g_max_batch * rpi;
        const int nv4 = (D + 127) / 128;
        const float *Xc = e->X, *gc = e->lnf_g, *bc = e->lnf_b;
        if (nv4 <= 3) launch_k(layernorm_gather_kernel<3>, dim3(grid), dim3(256), 0, st, Xc, gc, bc, dst, B, ntok, tok0, rpi, slot_row0, D, hp.eps);
        else if (nv4 <= 6) launch_k(layernorm_gather_kernel<6>, dim3(grid), dim3(256), 0, st, Xc, gc, bc, dst, B, ntok, tok0, rpi, slot_row0, D, hp.eps);
        else if (nv4 <= 8) launch_k(layernorm_gather_kernel<8>, dim3(grid), dim3(256), 0, st, Xc, gc, bc, dst, B, ntok, tok0, rpi, slot_row0, D, hp.eps);
        else launch_k(layernorm_gather_kernel<12>, dim3(grid), dim3(256), 0, st, Xc, gc, bc, dst, B, ntok, tok0, rpi, slot_row0, D, hp.eps);
        DINO_CUDA(cudaGetLastError());
        nl++;
        prof.end();
        if (!cls && !patch && !classify) return nl;        // nothing wanted locally: the stand-alone final LayerNorm is skipped
    }

    // 3. final LayerNorm (all tokens: the head pools over registers too) + outputs
    prof.begin(2);
    launch_layernorm(e->X, e->lnf_g, e->lnf_b, e->Y, M, D, hp.eps, false, st, next_dir());
    nl++;
    const size_t row_bytes = static_cast<size_t>(D) * sizeof(float);
    if (cls) DINO_CUDA(cudaMemcpy2DAsync(cls, row_bytes, e->Y, ntok * row_bytes, row_bytes, B, cudaMemcpyDeviceToDevice, st));
    if (patch)
        DINO_CUDA(cudaMemcpy2DAsync(patch, np * row_bytes, e->Y + static_cast<size_t>(1 + R) * D, ntok * row_bytes, np * row_bytes, B,
                                    cudaMemcpyDeviceToDevice, st));
    if (classify) {
        const int n_embd = hp.img_size / ps;
        launch_k(pool_tokens_kernel, dim3((D + 127) / 128, B), dim3(POOL_THREADS), 0, st, e->Y, e->feat, ntok, D, 1.0f / static_cast<float>(n_embd * n_embd));
        DINO_CUDA(cudaGetLastError());
        const int C = hp.num_classes;
        const long long warps = static_cast<long long>(B) * C;
        launch_k(classifier_kernel, dim3(static_cast<unsigned>((warps * 32 + 255) / 256)), dim3(256), 0, st, e->feat, e->wc, e->bc, e->logits, B, 2 * D, C);
        DINO_CUDA(cudaGetLastError());
        nl += 2;
        if (logits && logits != e->logits) DINO_CUDA(cudaMemcpyAsync(logits, e->logits, static_cast<size_t>(B) * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
        if (probs) {
            launch_k(softmax_rows_kernel, dim3(B), dim3(256), 0, st, e->logits, probs, C);
            DINO_CUDA(cudaGetLastError());
            nl++;
        }
    }
    prof.end();
    return nl;
}

// DINO_B200_GRAPH=0 disables CUDA-graph replay (every forward then enqueues its ~7 L + 6 kernels one by one)
static bool graphs_enabled() {
    static const bool v = [] { const char *e = getenv("DINO_B200_GRAPH"); return !(e && e[0] == '0'); }();
    return v;
}

static void forward_device(dino_b200_engine *e, const float *images, int layout, int B, int H, int W, int flags, float *cls,
                           float *patch, float *logits, float *probs, cudaStream_t st) {
    const dino_b200_hparams &hp = e->hp;
    const int ps = hp.patch_size;
    if (!images || B <= 0 || H < ps || W < ps) throw StatusError(DINO_B200_ERR_INVALID, "forward: bad batch or image size");
    if (H % ps || W % ps) throw StatusError(DINO_B200_ERR_INVALID, "forward: image size must be a multiple of the patch size");
    if (layout != DINO_B200_LAYOUT_RGB_PLANAR && layout != DINO_B200_LAYOUT_BGR_HWC) throw StatusError(DINO_B200_ERR_INVALID, "forward: unknown image layout");
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    if ((logits || probs) && !classify) throw StatusError(DINO_B200_ERR_INVALID, "forward: logits/probs require DINO_B200_CLASSIFY");
    if (classify && !e->wc) throw StatusError(DINO_B200_ERR_INVALID, "forward: checkpoint has no classifier head");

    const size_t n_pos_tables = e->pos_cache.size();
    const float *pos = pos_for_grid(e, H / ps, W / ps);        // may launch the bicubic resample on the engine stream (first use of a grid)
    if (e->pos_cache.size() != n_pos_tables && st != e->stream) DINO_CUDA(cudaStreamSynchronize(e->stream));

    if (e->profiling) {
        // per-kernel event pairs (bench roofline / tools): direct launches, never a graph
        for (auto &pe : e->prof) e->prof_pool.push_back(pe);
        e->prof.clear();
        if (!e->ev_t0) {
            DINO_CUDA(cudaEventCreate(&e->ev_t0));
            DINO_CUDA(cudaEventCreate(&e->ev_t1));
        }
        DINO_CUDA(cudaEventRecord(e->ev_t0, st));
        e->launches += forward_enqueue(e, images, layout, B, H, W, flags, cls, patch, logits, probs, pos, st);
        DINO_CUDA(cudaEventRecord(e->ev_t1, st));
        return;
    }
    if (!e->use_graphs || !graphs_enabled()) {
        e->launches += forward_enqueue(e, images, layout, B, H, W, flags, cls, patch, logits, probs, pos, st);
        return;
    }
    const GraphKey key{images, layout, B, H, W, flags, cls, patch, logits, probs};
    auto it = e->graphs.find(key);
    if (it == e->graphs.end()) {
        if (e->graphs.size() >= 16) drop_graphs(e);            // callers that cycle through many buffers: start over
        GraphEntry ge;
        cudaGraph_t graph = nullptr;
        DINO_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        try {
            ge.launches = forward_enqueue(e, images, layout, B, H, W, flags, cls, patch, logits, probs, pos, st);
        } catch (...) {
            cudaStreamEndCapture(st, &graph);                  // leave capture mode before reporting
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            throw;
        }
        DINO_CUDA(cudaStreamEndCapture(st, &graph));
        const cudaError_t ierr = cudaGraphInstantiate(&ge.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ierr != cudaSuccess) throw CudaError(std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ierr));
        e->graph_captures++;
        it = e->graphs.emplace(key, ge).first;
    }
    DINO_CUDA(cudaGraphLaunch(it->second.exec, st));
    e->graph_replays++;
    e->launches += it->second.launches;
}

}  // namespace dino

// ================================================================================================ C ABI
#define DINO_API_BEGIN try {
#define DINO_API_END(eng)                                                   \
    }                                                                       \
    catch (const dino::StatusError &ex) {                                   \
        dino::g_last_error = ex.what();                                     \
        if (eng) (eng)->err = ex.what();                                    \
        return ex.st;                                                       \
    }                                                                       \
    catch (const dino::CudaError &ex) {                                     \
        dino::g_last_error = ex.what();                                     \
        if (eng) (eng)->err = ex.what();                                    \
        return DINO_B200_ERR_CUDA;                                          \
    }                                                                       \
    catch (const std::exception &ex) {                                      \
        dino::g_last_error = ex.what();                                     \
        if (eng) (eng)->err = ex.what();                                    \
        return DINO_B200_ERR_INVALID;                                       \
    }

static bool device_is_sm100(int dev) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
    return major == 10;
}

extern "C" {

int dino_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; ++i) ok += device_is_sm100(i) ? 1 : 0;
    return ok;
}

static dino_b200_status require_device(int device) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        dino::g_last_error = "no CUDA device visible: dinov2_b200 has no CPU fallback";
        return DINO_B200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        dino::g_last_error = "device index out of range";
        return DINO_B200_ERR_INVALID;
    }
    if (!device_is_sm100(device)) {
        dino::g_last_error = "device is not compute capability 10.x (kernels are built for sm_100a only)";
        return DINO_B200_ERR_NO_DEVICE;
    }
    return DINO_B200_OK;
}

dino_b200_status dino_b200_create(const dino_b200_model_desc *desc, int device, dino_b200_engine **out) {
    if (!desc || !out || (desc->n_tensors > 0 && !desc->tensors)) {
        dino::g_last_error = "create: NULL argument";
        return DINO_B200_ERR_INVALID;
    }
    *out = nullptr;
    const dino_b200_status ds = require_device(device);
    if (ds != DINO_B200_OK) return ds;
    dino_b200_engine *e = nullptr;
    DINO_API_BEGIN
    DINO_CUDA(cudaSetDevice(device));
    const int num_sms = configure_current_device();
    e = new dino_b200_engine();
    e->device = device;
    e->num_sms = num_sms;
    DINO_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    try {
        build_engine(e, desc);
    } catch (...) {
        dino_b200_destroy(e);
        e = nullptr;
        throw;
    }
    *out = e;
    return DINO_B200_OK;
    DINO_API_END(static_cast<dino_b200_engine *>(nullptr))
}

dino_b200_status dino_b200_create_from_gguf(const char *path, int device, dino_b200_engine **out) {
    if (!path || !out) {
        dino::g_last_error = "create_from_gguf: NULL argument";
        return DINO_B200_ERR_INVALID;
    }
    *out = nullptr;
    dino::GGUFFile gg;
    try {
        dino::gguf_read(path, gg);
    } catch (const std::exception &ex) {
        dino::g_last_error = ex.what();
        return std::string(ex.what()).rfind("cannot open", 0) == 0 || std::string(ex.what()).rfind("short read", 0) == 0
                   ? DINO_B200_ERR_IO
                   : DINO_B200_ERR_FORMAT;
    }
    auto need = [&](const char *k, uint32_t &dst) -> bool {
        auto it = gg.kv_u.find(k);
        if (it == gg.kv_u.end()) {
            dino::g_last_error = std::string("gguf is missing key '") + k + "'";
            return false;
        }
        dst = static_cast<uint32_t>(it->second);
        return true;
    };
    dino_b200_model_desc desc{};
    dino_b200_hparams &hp = desc.hparams;
    if (!need("hidden_size", hp.hidden_size) || !need("num_hidden_layers", hp.num_hidden_layers) ||
        !need("num_attention_heads", hp.num_attention_heads) || !need("patch_size", hp.patch_size) ||
        !need("img_size", hp.img_size) || !need("ftype", hp.ftype) || !need("num_register_tokens", hp.num_register_tokens))
        return DINO_B200_ERR_FORMAT;
    hp.num_classes = gg.kv_u.count("num_classes") ? static_cast<uint32_t>(gg.kv_u["num_classes"]) : 0;
    hp.ftype %= 1000;   // GGML_QNT_VERSION_FACTOR (dinov2.cpp:307)
    hp.eps = 1e-6f;
    std::vector<dino_b200_tensor> ts(gg.tensors.size());
    for (size_t i = 0; i < ts.size(); ++i) {
        const auto &t = gg.tensors[i];
        ts[i].name = t.name.c_str();
        ts[i].type = t.type;
        ts[i].n_dims = t.n_dims;
        for (int d = 0; d < 4; ++d) ts[i].ne[d] = t.ne[d];
        ts[i].data = t.data;
        ts[i].nbytes = t.nbytes;
    }
    desc.n_tensors = static_cast<int32_t>(ts.size());
    desc.tensors = ts.data();
    // a class count without a matching classifier head is a malformed checkpoint, not an allocation request
    if (hp.num_classes > 0) {
        const dino::GGUFTensorInfo *cw = gg.find("classifier.weight");
        if (!cw || cw->n_dims < 2 || static_cast<uint64_t>(cw->ne[1]) != hp.num_classes) {
            dino::g_last_error = "gguf: num_classes does not match the classifier.weight tensor";
            return DINO_B200_ERR_FORMAT;
        }
    }
    const dino_b200_status st = dino_b200_create(&desc, device, out);
    if (st != DINO_B200_OK) return st;
    dino_b200_engine *eng = *out;
    DINO_API_BEGIN
    eng->labels.resize(hp.num_classes);
    for (uint32_t i = 0; i < hp.num_classes; ++i) {
        auto it = gg.kv_s.find(std::to_string(i));
        if (it != gg.kv_s.end()) eng->labels[i] = it->second;
    }
    return DINO_B200_OK;
    DINO_API_END(eng)
}

dino_b200_status dino_b200_quantize_gguf(const char *fname_inp, const char *fname_out, int ggml_type) {
    if (!fname_inp || !fname_out) {
        dino::g_last_error = "quantize_gguf: NULL argument";
        return DINO_B200_ERR_INVALID;
    }
    try {
        dino::quantize_gguf(fname_inp, fname_out, ggml_type);
        return DINO_B200_OK;
    } catch (const std::exception &ex) {
        dino::g_last_error = ex.what();
        const std::string m = ex.what();
        if (m.rfind("cannot open", 0) == 0 || m.rfind("short", 0) == 0) return DINO_B200_ERR_IO;
        if (m.rfind("quantize: target type", 0) == 0) return DINO_B200_ERR_INVALID;
        return DINO_B200_ERR_FORMAT;
    }
}

void dino_b200_destroy(dino_b200_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    dino::free_arena(e);
    for (void *p : e->allocs) cudaFree(p);
    if (e->d_u8) cudaFree(e->d_u8);
    dino::gather_release(e);
    for (int i = 0; i < 2; ++i) {
        if (e->d_u8s[i]) cudaFree(e->d_u8s[i]);
        if (e->r_rgb[i]) cudaFree(e->r_rgb[i]);
    }
    for (void *q : {static_cast<void *>(e->pca_mean), static_cast<void *>(e->pca_v), static_cast<void *>(e->pca_w), static_cast<void *>(e->pca_y),
                    static_cast<void *>(e->pca_x), static_cast<void *>(e->pca_rgb)})
        if (q) cudaFree(q);
    for (int i = 0; i < 2; ++i) {
        if (e->d_in[i]) cudaFree(e->d_in[i]);
        if (e->ev_up[i]) cudaEventDestroy(e->ev_up[i]);
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
        if (e->ev_fwd[i]) cudaEventDestroy(e->ev_fwd[i]);
        if (e->r_buf[i]) cudaFree(e->r_buf[i]);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->d2h_stream) cudaStreamDestroy(e->d2h_stream);
    for (auto &pe : e->prof) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    for (auto &pe : e->prof_pool) { cudaEventDestroy(pe.a); cudaEventDestroy(pe.b); }
    if (e->ev_t0) cudaEventDestroy(e->ev_t0);
    if (e->ev_t1) cudaEventDestroy(e->ev_t1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

dino_b200_status dino_b200_get_hparams(const dino_b200_engine *e, dino_b200_hparams *out) {
    if (!e || !out) return DINO_B200_ERR_INVALID;
    *out = e->hp;
    return DINO_B200_OK;
}

const char *dino_b200_label(const dino_b200_engine *e, int class_id) {
    if (!e || class_id < 0 || static_cast<size_t>(class_id) >= e->labels.size() || e->labels[class_id].empty()) return nullptr;
    return e->labels[class_id].c_str();
}

dino_b200_status dino_b200_reserve(dino_b200_engine *e, int max_batch, int H, int W) {
    if (!e || max_batch <= 0 || H <= 0 || W <= 0) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    DINO_CUDA(cudaSetDevice(e->device));
    dino::ensure_arena(e, max_batch, H, W);
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_set_pos_embed(dino_b200_engine *e, int gh, int gw, const float *pos) {
    if (!e || !pos || gh <= 0 || gw <= 0) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    DINO_CUDA(cudaSetDevice(e->device));
    const size_t n = static_cast<size_t>(1 + gh * gw) * e->hp.hidden_size;
    float *&slot = e->pos_cache[{gh, gw}];
    if (!slot) slot = static_cast<float *>(e->dmalloc(n * sizeof(float)));
    h2d(e, slot, pos, n * sizeof(float));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_get_pos_embed(dino_b200_engine *e, int H, int W, float *out) {
    if (!e || !out || H <= 0 || W <= 0) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    DINO_CUDA(cudaSetDevice(e->device));
    const int gh = H / e->hp.patch_size, gw = W / e->hp.patch_size;
    if (gh <= 0 || gw <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "image smaller than one patch");
    const float *p = dino::pos_for_grid(e, gh, gw);
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    DINO_CUDA(cudaMemcpy(out, p, static_cast<size_t>(1 + gh * gw) * e->hp.hidden_size * sizeof(float), cudaMemcpyDeviceToHost));
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_forward_device(dino_b200_engine *e, const float *images, int layout, int B, int H, int W, int flags,
                                          float *cls, float *patch, float *logits, float *probs, void *stream) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    DINO_CUDA(cudaSetDevice(e->device));
    dino::ensure_arena(e, B > 0 ? B : 1, H, W);
    dino::forward_device(e, images, layout, B, H, W, flags, cls, patch, logits, probs,
                         stream ? static_cast<cudaStream_t>(stream) : e->stream);
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_forward(dino_b200_engine *e, const float *images, int layout, int B, int H, int W, int flags, float *cls,
                                   float *patch, float *logits, float *probs) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!images || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "forward: bad batch or image size");
    DINO_CUDA(cudaSetDevice(e->device));
    dino::ensure_arena(e, B, H, W);
    const int ps = e->hp.patch_size, D = e->hp.hidden_size, C = e->hp.num_classes;
    const size_t np = static_cast<size_t>(H / ps) * (W / ps);
    if (patch && static_cast<size_t>(B) * np * D > e->cap_o_patch) {
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        if (e->o_patch) DINO_CUDA(cudaFree(e->o_patch));
        e->o_patch = nullptr;
        e->cap_o_patch = 0;
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->o_patch), static_cast<size_t>(B) * np * D * sizeof(float)));
        e->cap_o_patch = static_cast<size_t>(B) * np * D;
    }
    cudaStream_t st = e->stream;
    DINO_CUDA(cudaMemcpyAsync(e->d_img, images, static_cast<size_t>(B) * 3 * H * W * sizeof(float), cudaMemcpyHostToDevice, st));
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    dino::forward_device(e, e->d_img, layout, B, H, W, flags, cls ? e->o_cls : nullptr, patch ? e->o_patch : nullptr,
                         (classify && logits) ? e->logits : nullptr, (classify && probs) ? e->probs : nullptr, st);
    if (cls) DINO_CUDA(cudaMemcpyAsync(cls, e->o_cls, static_cast<size_t>(B) * D * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (patch) DINO_CUDA(cudaMemcpyAsync(patch, e->o_patch, static_cast<size_t>(B) * np * D * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (classify && logits) DINO_CUDA(cudaMemcpyAsync(logits, e->logits, static_cast<size_t>(B) * C * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (classify && probs) DINO_CUDA(cudaMemcpyAsync(probs, e->probs, static_cast<size_t>(B) * C * sizeof(float), cudaMemcpyDeviceToHost, st));
    DINO_CUDA(cudaStreamSynchronize(st));
    return DINO_B200_OK;
    DINO_API_END(e)
}

// Pipelined host interface.  Slot k % 2 of batch k: upload on the copy stream (may run under the forward pass of batch
// k-1), then forward + read-back on the engine stream.  At most two batches in flight.
dino_b200_status dino_b200_submit(dino_b200_engine *e, const float *images, int layout, int B, int H, int W, int flags, float *cls,
                                  float *patch, float *logits, float *probs) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!images || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit: bad batch or image size");
    if (e->n_submitted - e->n_waited >= 2) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit: two batches already in flight; call dino_b200_wait first");
    DINO_CUDA(cudaSetDevice(e->device));
    const int slot = static_cast<int>(e->n_submitted & 1);
    if (!e->copy_stream) {
        DINO_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        DINO_CUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            DINO_CUDA(cudaEventCreateWithFlags(&e->ev_up[i], cudaEventDisableTiming));
            DINO_CUDA(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
            DINO_CUDA(cudaEventCreateWithFlags(&e->ev_fwd[i], cudaEventDisableTiming));
        }
    }
    const int ps = e->hp.patch_size, D = e->hp.hidden_size, C = e->hp.num_classes;
    const size_t np = static_cast<size_t>(H / ps) * (W / ps);
    const size_t n_in = static_cast<size_t>(B) * 3 * H * W;
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    if (classify && !e->wc) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit: checkpoint has no classifier head");
    // per-slot device results, so that the read-back of batch k (own stream) may run under the forward pass of batch k+1
    const size_t n_cls = cls ? static_cast<size_t>(B) * D : 0, n_log = (classify && logits) ? static_cast<size_t>(B) * C : 0;
    const size_t n_prob = (classify && probs) ? static_cast<size_t>(B) * C : 0, n_patch = patch ? static_cast<size_t>(B) * np * D : 0;
    const size_t n_res = n_cls + n_log + n_prob + n_patch;
    // anything that (re)allocates waits for the work in flight first
    if (n_in > e->cap_in[slot] || n_res > e->cap_r[slot]) {
        DINO_CUDA(cudaStreamSynchronize(e->copy_stream));
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        DINO_CUDA(cudaStreamSynchronize(e->d2h_stream));
        if (n_in > e->cap_in[slot]) {
            if (e->d_in[slot]) DINO_CUDA(cudaFree(e->d_in[slot]));
            e->d_in[slot] = nullptr;
            e->cap_in[slot] = 0;
            DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->d_in[slot]), n_in * sizeof(float)));
            e->cap_in[slot] = n_in;
        }
        if (n_res > e->cap_r[slot]) {
            if (e->r_buf[slot]) DINO_CUDA(cudaFree(e->r_buf[slot]));
            e->r_buf[slot] = nullptr;
            e->cap_r[slot] = 0;
            DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->r_buf[slot]), n_res * sizeof(float)));
            e->cap_r[slot] = n_res;
        }
    }
    dino::ensure_arena(e, B, H, W);     // synchronises the engine stream itself when it has to grow
    float *r_cls = e->r_buf[slot], *r_log = r_cls + n_cls, *r_prob = r_log + n_log, *r_patch = r_prob + n_prob;
    cudaStream_t st = e->stream;
    // slot reuse (batch k-2): its forward pass has read the input slot and its read-back has left the result slot
    if (e->n_submitted >= 2) {
        DINO_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_done[slot], 0));
        DINO_CUDA(cudaStreamWaitEvent(st, e->ev_done[slot], 0));
    }
    DINO_CUDA(cudaMemcpyAsync(e->d_in[slot], images, n_in * sizeof(float), cudaMemcpyHostToDevice, e->copy_stream));
    DINO_CUDA(cudaEventRecord(e->ev_up[slot], e->copy_stream));
    DINO_CUDA(cudaStreamWaitEvent(st, e->ev_up[slot], 0));
    dino::forward_device(e, e->d_in[slot], layout, B, H, W, flags, n_cls ? r_cls : nullptr, n_patch ? r_patch : nullptr,
                         n_log ? r_log : nullptr, n_prob ? r_prob : nullptr, st);
    DINO_CUDA(cudaEventRecord(e->ev_fwd[slot], st));
    DINO_CUDA(cudaStreamWaitEvent(e->d2h_stream, e->ev_fwd[slot], 0));
    if (n_cls) DINO_CUDA(cudaMemcpyAsync(cls, r_cls, n_cls * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (n_patch) DINO_CUDA(cudaMemcpyAsync(patch, r_patch, n_patch * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (n_log) DINO_CUDA(cudaMemcpyAsync(logits, r_log, n_log * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (n_prob) DINO_CUDA(cudaMemcpyAsync(probs, r_prob, n_prob * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    DINO_CUDA(cudaEventRecord(e->ev_done[slot], e->d2h_stream));
    e->n_submitted++;
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_wait(dino_b200_engine *e) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (e->n_waited >= e->n_submitted) throw dino::StatusError(DINO_B200_ERR_INVALID, "wait: nothing in flight");
    DINO_CUDA(cudaSetDevice(e->device));
    DINO_CUDA(cudaEventSynchronize(e->ev_done[e->n_waited & 1]));
    e->n_waited++;
    return DINO_B200_OK;
    DINO_API_END(e)
}

// ------------------------------------------------------------------------------------------------ PCA colouring
namespace dino {
static void pca_reserve(dino_b200_engine *e, int B, int NP, bool host_staging) {
    const int D = e->hp.hidden_size;
    if (static_cast<size_t>(B) > e->pca_cap_b || static_cast<size_t>(NP) > e->pca_cap_np) {
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        for (float **q : {&e->pca_mean, &e->pca_v, &e->pca_w, &e->pca_y}) {
            if (*q) DINO_CUDA(cudaFree(*q));
            *q = nullptr;
        }
        if (e->pca_rgb) DINO_CUDA(cudaFree(e->pca_rgb));
        e->pca_rgb = nullptr;
        const size_t nb = std::max<size_t>(B, e->pca_cap_b), np = std::max<size_t>(NP, e->pca_cap_np);
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->pca_mean), nb * D * sizeof(float)));
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->pca_v), nb * D * 3 * sizeof(float)));
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->pca_w), nb * D * 3 * sizeof(float)));
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->pca_y), nb * np * 3 * sizeof(float)));
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->pca_rgb), nb * np * 3));
        e->pca_cap_b = nb;
        e->pca_cap_np = np;
    }
    const size_t nx = static_cast<size_t>(B) * NP * D;
    if (host_staging && nx > e->pca_cap_x) {
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        if (e->pca_x) DINO_CUDA(cudaFree(e->pca_x));
        e->pca_x = nullptr;
        e->pca_cap_x = 0;
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->pca_x), nx * sizeof(float)));
        e->pca_cap_x = nx;
    }
}

// X: device [B][NP][D]; rgb: device [B][NP][3] u8 (may be null); proj: device [B][NP][3] (may be null)
static void pca_device(dino_b200_engine *e, const float *X, int B, int NP, uint8_t *rgb, float *proj, cudaStream_t st) {
    const int D = e->hp.hidden_size;
    const dim3 gcol((D + 127) / 128, B);
    pca_mean_kernel<<<gcol, 128, 0, st>>>(X, e->pca_mean, NP, D);
    pca_init_kernel<<<gcol, 128, 0, st>>>(e->pca_w, D);
    pca_orth_kernel<<<B, 256, 0, st>>>(e->pca_w, e->pca_v, D, 0);
    const dim3 grow((NP + 7) / 8, B);
    const int splits = std::max(1, std::min(16, NP / 64));
    const int rows_per_split = (NP + splits - 1) / splits;
    const dim3 gback((D + 127) / 128, B, splits);
    for (int it = 0; it < PCA_ITERS; ++it) {
        pca_project_kernel<<<grow, 256, 0, st>>>(X, e->pca_mean, e->pca_v, e->pca_y, NP, D);
        pca_backproject_kernel<<<gback, 128, 0, st>>>(X, e->pca_mean, e->pca_y, e->pca_w, NP, D, rows_per_split);
        pca_orth_kernel<<<B, 256, 0, st>>>(e->pca_w, e->pca_v, D, it == PCA_ITERS - 1);
    }
    pca_project_kernel<<<grow, 256, 0, st>>>(X, e->pca_mean, e->pca_v, e->pca_y, NP, D);
    if (rgb) pca_to_u8_kernel<<<B, 256, 0, st>>>(e->pca_y, rgb, NP * 3);
    DINO_CUDA(cudaGetLastError());
    if (proj) DINO_CUDA(cudaMemcpyAsync(proj, e->pca_y, static_cast<size_t>(B) * NP * 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    e->launches += 4 + 3 * PCA_ITERS + (rgb ? 1 : 0);
}
}  // namespace dino

dino_b200_status dino_b200_pca_rgb_device(dino_b200_engine *e, const float *patch, int B, int NP, uint8_t *rgb, float *proj, void *stream) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!patch || B <= 0 || NP < 3 || (!rgb && !proj)) throw dino::StatusError(DINO_B200_ERR_INVALID, "pca_rgb: bad argument (needs at least 3 patches)");
    DINO_CUDA(cudaSetDevice(e->device));
    dino::pca_reserve(e, B, NP, false);
    dino::pca_device(e, patch, B, NP, rgb, proj, stream ? static_cast<cudaStream_t>(stream) : e->stream);
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_pca_rgb(dino_b200_engine *e, const float *patch, int B, int NP, uint8_t *rgb, float *proj) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!patch || B <= 0 || NP < 3 || (!rgb && !proj)) throw dino::StatusError(DINO_B200_ERR_INVALID, "pca_rgb: bad argument (needs at least 3 patches)");
    DINO_CUDA(cudaSetDevice(e->device));
    dino::pca_reserve(e, B, NP, true);
    const size_t nx = static_cast<size_t>(B) * NP * e->hp.hidden_size;
    DINO_CUDA(cudaMemcpyAsync(e->pca_x, patch, nx * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    dino::pca_device(e, e->pca_x, B, NP, rgb ? e->pca_rgb : nullptr, nullptr, e->stream);
    if (rgb) DINO_CUDA(cudaMemcpyAsync(rgb, e->pca_rgb, static_cast<size_t>(B) * NP * 3, cudaMemcpyDeviceToHost, e->stream));
    if (proj) DINO_CUDA(cudaMemcpyAsync(proj, e->pca_y, static_cast<size_t>(B) * NP * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    return DINO_B200_OK;
    DINO_API_END(e)
}

dino_b200_status dino_b200_synchronize(dino_b200_engine *e) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    DINO_CUDA(cudaSetDevice(e->device));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    return DINO_B200_OK;
    DINO_API_END(e)
}

const char *dino_b200_last_error(const dino_b200_engine *e) { return e ? e->err.c_str() : dino::g_last_error.c_str(); }

uint64_t dino_b200_kernel_launches(const dino_b200_engine *e) { return e ? e->launches : 0; }

dino_b200_status dino_b200_set_profiling(dino_b200_engine *e, int on) {
    if (!e) return DINO_B200_ERR_INVALID;
    e->profiling = on != 0;
    return DINO_B200_OK;
}

dino_b200_status dino_b200_get_profile(dino_b200_engine *e, float *gemm_ms, float *attn_ms, float *other_ms, float *total_ms) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!e->ev_t1) throw dino::StatusError(DINO_B200_ERR_INVALID, "no profiled forward has run");
    DINO_CUDA(cudaEventSynchronize(e->ev_t1));
    float acc[3] = {0.f, 0.f, 0.f};
    for (auto &pe : e->prof) {
        float ms = 0.f;
        DINO_CUDA(cudaEventElapsedTime(&ms, pe.a, pe.b));
        acc[pe.kind] += ms;
    }
    float tot = 0.f;
    DINO_CUDA(cudaEventElapsedTime(&tot, e->ev_t0, e->ev_t1));
    if (gemm_ms) *gemm_ms = acc[0];
    if (attn_ms) *attn_ms = acc[1];
    if (other_ms) *other_ms = acc[2];
    if (total_ms) *total_ms = tot;
    return DINO_B200_OK;
    DINO_API_END(e)
}

// ------------------------------------------------------------------------------------------------ kernel hooks
dino_b200_status dino_b200_kernel_gemm(int epi, const void *A, int lda, const void *W, int ldw, int M, int N, int K, const float *bias,
                                       const float *lscale, void *out, int ldo, const float *pos, int np, int ntok, int tok_off,
                                       void *stream) {
    DINO_API_BEGIN
    if (!A || !W || !out || !bias) throw dino::StatusError(DINO_B200_ERR_INVALID, "kernel_gemm: NULL argument");
    int dev = 0;
    DINO_CUDA(cudaGetDevice(&dev));
    if (!device_is_sm100(dev)) throw dino::StatusError(DINO_B200_ERR_NO_DEVICE, "current device is not sm_100");
    const int num_sms = configure_current_device();
    const dino::GemmPlan pl = dino::plan_gemm(epi, M, N, num_sms);
    const CUtensorMap tmA = dino::make_tmap_f16(A, K, M, lda, GEMM_BM);
    const CUtensorMap tmB = dino::make_tmap_f16(W, K, N, ldw, pl.BN / pl.CG / pl.MC);
    dino::GemmParams gp{};
    gp.M = M; gp.N = N; gp.K = K; gp.bias = bias; gp.lscale = lscale; gp.out = out; gp.ldo = ldo;
    gp.pos = pos; gp.np = np; gp.ntok = ntok; gp.tok_off = tok_off;
    const int out_cols = epi == DINO_B200_EPI_SWIGLU_F16 ? N / 2 : N;
    const uint64_t out_rows = epi == DINO_B200_EPI_PATCH_F32 ? static_cast<uint64_t>(M / (np > 0 ? np : 1)) * ntok : static_cast<uint64_t>(M);
    const CUtensorMap tmC = dino::make_tmap_out(epi, out, out_cols, out_rows, ldo);
    dino::launch_gemm(epi, pl, tmA, tmB, tmC, gp, static_cast<cudaStream_t>(stream));
    return DINO_B200_OK;
    DINO_API_END(static_cast<dino_b200_engine *>(nullptr))
}

dino_b200_status dino_b200_kernel_gemm_resid_ln(const void *A, int lda, const void *W, int ldw, int M, int N, int K, const float *bias,
                                                const float *lscale, float *X, const float *gamma, const float *beta, float eps,
                                                void *ln_out, int *counters, void *stream) {
    DINO_API_BEGIN
    if (!A || !W || !X || !bias || !lscale || !gamma || !beta || !ln_out || !counters)
        throw dino::StatusError(DINO_B200_ERR_INVALID, "kernel_gemm_resid_ln: NULL argument");
    if (!dino::gemm_ln_width_ok(N)) throw dino::StatusError(DINO_B200_ERR_UNSUPPORTED, "kernel_gemm_resid_ln: N must be 384, 768, 1024 or 1536");
    int dev = 0;
    DINO_CUDA(cudaGetDevice(&dev));
    if (!device_is_sm100(dev)) throw dino::StatusError(DINO_B200_ERR_NO_DEVICE, "current device is not sm_100");
    const int num_sms = configure_current_device();
    const dino::GemmPlan pl = dino::plan_gemm(dino::EPI_RESID_LN_F32, M, N, num_sms);
    const CUtensorMap tmA = dino::make_tmap_f16(A, K, M, lda, GEMM_BM);
    const CUtensorMap tmB = dino::make_tmap_f16(W, K, N, ldw, pl.BN / pl.CG);
    dino::GemmParams gp{};
    gp.M = M; gp.N = N; gp.K = K; gp.bias = bias; gp.lscale = lscale; gp.out = X; gp.ldo = N;
    gp.ln_gamma = gamma; gp.ln_beta = beta; gp.ln_out = static_cast<__half *>(ln_out); gp.ln_eps = eps; gp.ln_count = counters;
    const CUtensorMap tmC = dino::make_tmap_out(dino::EPI_RESID_LN_F32, X, N, M, N);
    dino::launch_gemm(dino::EPI_RESID_LN_F32, pl, tmA, tmB, tmC, gp, static_cast<cudaStream_t>(stream));
    return DINO_B200_OK;
    DINO_API_END(static_cast<dino_b200_engine *>(nullptr))
}

dino_b200_status dino_b200_kernel_attention(const void *qkv, void *out, int B, int n_tok, int D, void *stream) {
    DINO_API_BEGIN
    if (!qkv || !out || B <= 0 || n_tok <= 0 || D <= 0 || D % ATT_HD) throw dino::StatusError(DINO_B200_ERR_INVALID, "kernel_attention: bad argument");
    int dev = 0;
    DINO_CUDA(cudaGetDevice(&dev));
    if (!device_is_sm100(dev)) throw dino::StatusError(DINO_B200_ERR_NO_DEVICE, "current device is not sm_100");
    const int num_sms = configure_current_device();
    const CUtensorMap tm = dino::make_tmap_2d(qkv, false, 3 * D, static_cast<uint64_t>(B) * n_tok, 3 * D, ATT_BKV, dino::attn_promotion());
    dino::launch_attention(tm, static_cast<__half *>(out), B, n_tok, D, num_sms, static_cast<cudaStream_t>(stream));
    return DINO_B200_OK;
    DINO_API_END(static_cast<dino_b200_engine *>(nullptr))
}

dino_b200_status dino_b200_kernel_layernorm(const float *X, const float *gamma, const float *beta, void *out, int rows, int D, float eps,
                                            int out_half, void *stream) {
    DINO_API_BEGIN
    if (!X || !gamma || !beta || !out || rows <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "kernel_layernorm: bad argument");
    dino::launch_layernorm(X, gamma, beta, out, rows, D, eps, out_half != 0, static_cast<cudaStream_t>(stream));
    return DINO_B200_OK;
    DINO_API_END(static_cast<dino_b200_engine *>(nullptr))
}

}  // extern "C"

namespace dino {
// output size of dino_preprocess / dino_classify_preprocess for an H x W frame (reference dinov2.cpp:106-156)
static void preprocess_size(const dino_b200_engine *e, int H, int W, bool classify, int &RH, int &RW, int &OH, int &OW, int &cy, int &cx) {
    const int ps = e->hp.patch_size;
    cy = cx = 0;
    if (classify) {
        RH = RW = 256;
        OH = OW = 224;
        cy = (RH - OH) / 2;
        cx = (RW - OW) / 2;
    } else {
        RW = (W / ps + 1) * ps;
        RH = (H / ps + 1) * ps;
        OH = RH;
        OW = RW;
    }
}
// runs the preprocessing kernel: raw frames at src (device), float32 BGR [B][OH][OW][3] to dst (device)
static void preprocess_launch(dino_b200_engine *e, const uint8_t *src, float *dst, int B, int H, int W, bool classify, cudaStream_t st) {
    int RH, RW, OH, OW, cy, cx;
    preprocess_size(e, H, W, classify, RH, RW, OH, OW, cy, cx);
    // IMAGENET mean / std are R,G,B (reference dinov2.h:16-17); the image is B,G,R
    const float3 mean = make_float3(0.406f, 0.456f, 0.485f);
    const float3 inv_std = make_float3(1.0f / 0.225f, 1.0f / 0.224f, 1.0f / 0.229f);
    const long long total = static_cast<long long>(B) * OH * OW;
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, static_cast<long long>(e->num_sms) * 32));
    preprocess_bicubic_kernel<<<grid, 256, 0, st>>>(src, dst, B, H, W, RH, RW, OH, OW, cy, cx, mean, inv_std);
    DINO_CUDA(cudaGetLastError());
    e->launches++;
}
// frames already in e->d_u8, result in e->d_img; returns output size
static void preprocess_device(dino_b200_engine *e, int B, int H, int W, bool classify, int &OH, int &OW) {
    int RH, RW, cy, cx;
    preprocess_size(e, H, W, classify, RH, RW, OH, OW, cy, cx);
    ensure_arena(e, B, OH, OW);
    preprocess_launch(e, e->d_u8, e->d_img, B, H, W, classify, e->stream);
}

static void upload_frames(dino_b200_engine *e, const uint8_t *images, int B, int H, int W) {
    const size_t bytes = static_cast<size_t>(B) * H * W * 3;
    if (bytes > e->cap_u8) {
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        if (e->d_u8) DINO_CUDA(cudaFree(e->d_u8));
        e->d_u8 = nullptr;
        e->cap_u8 = 0;
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->d_u8), bytes));
        e->cap_u8 = bytes;
    }
    DINO_CUDA(cudaMemcpyAsync(e->d_u8, images, bytes, cudaMemcpyHostToDevice, e->stream));
}
}  // namespace dino

extern "C" dino_b200_status dino_b200_preprocess(dino_b200_engine *e, const uint8_t *images, int B, int H, int W, int classify,
                                                 float *out, int *out_h, int *out_w) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!images || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "preprocess: bad batch or image size");
    DINO_CUDA(cudaSetDevice(e->device));
    dino::upload_frames(e, images, B, H, W);
    int OH = 0, OW = 0;
    dino::preprocess_device(e, B, H, W, classify != 0, OH, OW);
    if (out) DINO_CUDA(cudaMemcpyAsync(out, e->d_img, static_cast<size_t>(B) * OH * OW * 3 * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    if (out_h) *out_h = OH;
    if (out_w) *out_w = OW;
    return DINO_B200_OK;
    DINO_API_END(e)
}

extern "C" dino_b200_status dino_b200_forward_u8(dino_b200_engine *e, const uint8_t *images, int B, int H, int W, int flags, float *cls,
                                                 float *patch, float *logits, float *probs) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!images || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "forward_u8: bad batch or image size");
    DINO_CUDA(cudaSetDevice(e->device));
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    dino::upload_frames(e, images, B, H, W);
    int OH = 0, OW = 0;
    dino::preprocess_device(e, B, H, W, classify, OH, OW);
    const int ps = e->hp.patch_size, D = e->hp.hidden_size, C = e->hp.num_classes;
    const size_t np = static_cast<size_t>(OH / ps) * (OW / ps);
    if (patch && static_cast<size_t>(B) * np * D > e->cap_o_patch) {
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        if (e->o_patch) DINO_CUDA(cudaFree(e->o_patch));
        e->o_patch = nullptr;
        e->cap_o_patch = 0;
        DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->o_patch), static_cast<size_t>(B) * np * D * sizeof(float)));
        e->cap_o_patch = static_cast<size_t>(B) * np * D;
    }
    cudaStream_t st = e->stream;
    dino::forward_device(e, e->d_img, DINO_B200_LAYOUT_BGR_HWC, B, OH, OW, flags, cls ? e->o_cls : nullptr, patch ? e->o_patch : nullptr,
                         (classify && logits) ? e->logits : nullptr, (classify && probs) ? e->probs : nullptr, st);
    if (cls) DINO_CUDA(cudaMemcpyAsync(cls, e->o_cls, static_cast<size_t>(B) * D * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (patch) DINO_CUDA(cudaMemcpyAsync(patch, e->o_patch, static_cast<size_t>(B) * np * D * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (classify && logits) DINO_CUDA(cudaMemcpyAsync(logits, e->logits, static_cast<size_t>(B) * C * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (classify && probs) DINO_CUDA(cudaMemcpyAsync(probs, e->probs, static_cast<size_t>(B) * C * sizeof(float), cudaMemcpyDeviceToHost, st));
    DINO_CUDA(cudaStreamSynchronize(st));
    return DINO_B200_OK;
    DINO_API_END(e)
}


// ================================================================================================ pipelined raw-frame interface
// The realtime caller's per-frame loop (reference realtime.cpp:75-101: capture -> dino_preprocess -> dino_predict -> PCA -> show)
// as one asynchronous submission: uint8 frames in, features and / or PCA colours out, nothing but the frame and the results
// crosses PCIe.  Shares the two-slot pipeline (and dino_b200_wait) with dino_b200_submit.
extern "C" dino_b200_status dino_b200_submit_u8(dino_b200_engine *e, const uint8_t *frames, int B, int H, int W, int flags, float *cls,
                                                float *patch, float *logits, float *probs, uint8_t *pca_rgb, int *out_h, int *out_w) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!frames || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit_u8: bad batch or frame size");
    if (e->n_submitted - e->n_waited >= 2) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit_u8: two batches already in flight; call dino_b200_wait first");
    DINO_CUDA(cudaSetDevice(e->device));
    const int slot = static_cast<int>(e->n_submitted & 1);
    if (!e->copy_stream) {
        DINO_CUDA(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        DINO_CUDA(cudaStreamCreateWithFlags(&e->d2h_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            DINO_CUDA(cudaEventCreateWithFlags(&e->ev_up[i], cudaEventDisableTiming));
            DINO_CUDA(cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
            DINO_CUDA(cudaEventCreateWithFlags(&e->ev_fwd[i], cudaEventDisableTiming));
        }
    }
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    if (classify && !e->wc) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit_u8: checkpoint has no classifier head");
    int RH, RW, OH, OW, cy, cx;
    dino::preprocess_size(e, H, W, classify, RH, RW, OH, OW, cy, cx);
    if (out_h) *out_h = OH;
    if (out_w) *out_w = OW;
    const int ps = e->hp.patch_size, D = e->hp.hidden_size, C = e->hp.num_classes;
    const size_t np = static_cast<size_t>(OH / ps) * (OW / ps);
    if (pca_rgb && np < 3) throw dino::StatusError(DINO_B200_ERR_INVALID, "submit_u8: PCA colouring needs at least 3 patches");
    const size_t n_u8 = static_cast<size_t>(B) * H * W * 3, n_in = static_cast<size_t>(B) * 3 * OH * OW;
    const bool need_patch = patch || pca_rgb;            // the PCA reads the patch tokens on the device even when the host does not want them
    const size_t n_cls = cls ? static_cast<size_t>(B) * D : 0, n_log = (classify && logits) ? static_cast<size_t>(B) * C : 0;
    const size_t n_prob = (classify && probs) ? static_cast<size_t>(B) * C : 0, n_patch = need_patch ? static_cast<size_t>(B) * np * D : 0;
    const size_t n_res = n_cls + n_log + n_prob + n_patch, n_rgb = pca_rgb ? static_cast<size_t>(B) * np * 3 : 0;
    if (n_u8 > e->cap_u8s[slot] || n_in > e->cap_in[slot] || n_res > e->cap_r[slot] || n_rgb > e->cap_rgb[slot]) {
        DINO_CUDA(cudaStreamSynchronize(e->copy_stream));
        DINO_CUDA(cudaStreamSynchronize(e->stream));
        DINO_CUDA(cudaStreamSynchronize(e->d2h_stream));
        auto grow = [&](void **ptr, size_t &cap, size_t need, size_t elem) {
            if (need <= cap) return;
            if (*ptr) DINO_CUDA(cudaFree(*ptr));
            *ptr = nullptr;
            cap = 0;
            DINO_CUDA(cudaMalloc(ptr, need * elem));
            cap = need;
        };
        grow(reinterpret_cast<void **>(&e->d_u8s[slot]), e->cap_u8s[slot], n_u8, 1);
        grow(reinterpret_cast<void **>(&e->d_in[slot]), e->cap_in[slot], n_in, sizeof(float));
        grow(reinterpret_cast<void **>(&e->r_buf[slot]), e->cap_r[slot], n_res, sizeof(float));
        grow(reinterpret_cast<void **>(&e->r_rgb[slot]), e->cap_rgb[slot], n_rgb, 1);
    }
    dino::ensure_arena(e, B, OH, OW);
    if (pca_rgb) dino::pca_reserve(e, B, static_cast<int>(np), false);
    float *r_cls = e->r_buf[slot], *r_log = r_cls + n_cls, *r_prob = r_log + n_log, *r_patch = r_prob + n_prob;
    cudaStream_t st = e->stream;
    if (e->n_submitted >= 2) {
        DINO_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_done[slot], 0));
        DINO_CUDA(cudaStreamWaitEvent(st, e->ev_done[slot], 0));
    }
    DINO_CUDA(cudaMemcpyAsync(e->d_u8s[slot], frames, n_u8, cudaMemcpyHostToDevice, e->copy_stream));
    DINO_CUDA(cudaEventRecord(e->ev_up[slot], e->copy_stream));
    DINO_CUDA(cudaStreamWaitEvent(st, e->ev_up[slot], 0));
    dino::preprocess_launch(e, e->d_u8s[slot], e->d_in[slot], B, H, W, classify, st);
    dino::forward_device(e, e->d_in[slot], DINO_B200_LAYOUT_BGR_HWC, B, OH, OW, flags & (DINO_B200_CLASSIFY | DINO_B200_FLASH_ATTN_COMPAT), n_cls ? r_cls : nullptr,
                         n_patch ? r_patch : nullptr, n_log ? r_log : nullptr, n_prob ? r_prob : nullptr, st);
    if (pca_rgb) dino::pca_device(e, r_patch, B, static_cast<int>(np), e->r_rgb[slot], nullptr, st);
    DINO_CUDA(cudaEventRecord(e->ev_fwd[slot], st));
    DINO_CUDA(cudaStreamWaitEvent(e->d2h_stream, e->ev_fwd[slot], 0));
    if (n_cls) DINO_CUDA(cudaMemcpyAsync(cls, r_cls, n_cls * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (patch) DINO_CUDA(cudaMemcpyAsync(patch, r_patch, n_patch * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (n_log) DINO_CUDA(cudaMemcpyAsync(logits, r_log, n_log * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (n_prob) DINO_CUDA(cudaMemcpyAsync(probs, r_prob, n_prob * sizeof(float), cudaMemcpyDeviceToHost, e->d2h_stream));
    if (n_rgb) DINO_CUDA(cudaMemcpyAsync(pca_rgb, e->r_rgb[slot], n_rgb, cudaMemcpyDeviceToHost, e->d2h_stream));
    DINO_CUDA(cudaEventRecord(e->ev_done[slot], e->d2h_stream));
    e->n_submitted++;
    return DINO_B200_OK;
    DINO_API_END(e)
}

// ================================================================================================ feature all-gather
// SURVEY.md 8e / BASELINE north_star: the data-parallel path has exactly one optional exchange — every rank ends up with the
// [cls] or patch embeddings of the WHOLE global batch.  Here it is not a separate collective: the final LayerNorm kernel of
// each rank stores its rows directly into every rank's gather buffer (peer memory over NVLink / NVSwitch).
namespace dino {
static void gather_release(dino_b200_engine *e) {
    for (int r = 0; r < 8; ++r) {
        if (e->g_ipc[r] && e->g_peer[r]) cudaIpcCloseMemHandle(e->g_peer[r]);
        e->g_peer[r] = nullptr;
        e->g_ipc[r] = false;
    }
    if (e->g_buf) cudaFree(e->g_buf);
    e->g_buf = nullptr;
    e->g_world = 0;
}
}  // namespace dino

extern "C" dino_b200_status dino_b200_gather_init(dino_b200_engine *e, int rank, int world, int what, int max_batch, int H, int W,
                                                  void **local_buf, unsigned char *ipc_handle) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    const int ps = e->hp.patch_size;
    if (world < 1 || world > 8 || rank < 0 || rank >= world || max_batch <= 0 || H < ps || W < ps || H % ps || W % ps ||
        (what != DINO_B200_GATHER_CLS && what != DINO_B200_GATHER_PATCH))
        throw dino::StatusError(DINO_B200_ERR_INVALID, "gather_init: bad rank / world (1..8) / batch / image size / selector");
    DINO_CUDA(cudaSetDevice(e->device));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    dino::drop_graphs(e);
    dino::gather_release(e);
    e->g_rank = rank;
    e->g_world = world;
    e->g_what = what;
    e->g_max_batch = max_batch;
    e->g_H = H;
    e->g_W = W;
    e->g_rpi = what == DINO_B200_GATHER_CLS ? 1 : (H / ps) * (W / ps);
    const size_t bytes = static_cast<size_t>(world) * max_batch * e->g_rpi * e->hp.hidden_size * sizeof(float);
    DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->g_buf), bytes));
    DINO_CUDA(cudaMemsetAsync(e->g_buf, 0, bytes, e->stream));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    e->g_peer[rank] = e->g_buf;
    if (local_buf) *local_buf = e->g_buf;
    if (ipc_handle) {
        cudaIpcMemHandle_t h;
        DINO_CUDA(cudaIpcGetMemHandle(&h, e->g_buf));
        static_assert(sizeof(h) == DINO_B200_IPC_HANDLE_BYTES, "CUDA IPC handle size");
        memcpy(ipc_handle, &h, sizeof(h));
    }
    return DINO_B200_OK;
    DINO_API_END(e)
}

extern "C" dino_b200_status dino_b200_gather_set_peer(dino_b200_engine *e, int r, void *dev_ptr, const unsigned char *ipc_handle) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (e->g_world == 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "gather_set_peer: call dino_b200_gather_init first");
    if (r < 0 || r >= e->g_world || r == e->g_rank || (!dev_ptr && !ipc_handle))
        throw dino::StatusError(DINO_B200_ERR_INVALID, "gather_set_peer: bad peer rank or no buffer given");
    DINO_CUDA(cudaSetDevice(e->device));
    DINO_CUDA(cudaStreamSynchronize(e->stream));
    dino::drop_graphs(e);
    if (e->g_ipc[r] && e->g_peer[r]) cudaIpcCloseMemHandle(e->g_peer[r]);
    e->g_peer[r] = nullptr;
    e->g_ipc[r] = false;
    if (dev_ptr) {
        // a buffer of another engine of THIS process: make the owning device's memory addressable from ours
        cudaPointerAttributes at{};
        DINO_CUDA(cudaPointerGetAttributes(&at, dev_ptr));
        if (at.type != cudaMemoryTypeDevice) throw dino::StatusError(DINO_B200_ERR_INVALID, "gather_set_peer: not a device pointer");
        if (at.device != e->device) {
            int can = 0;
            DINO_CUDA(cudaDeviceCanAccessPeer(&can, e->device, at.device));
            if (!can) throw dino::StatusError(DINO_B200_ERR_UNSUPPORTED, "gather_set_peer: no peer access between the two devices");
            const cudaError_t pe = cudaDeviceEnablePeerAccess(at.device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) DINO_CUDA(pe);
            cudaGetLastError();
        }
        e->g_peer[r] = static_cast<float *>(dev_ptr);
    } else {
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handle, sizeof(h));
        void *p = nullptr;
        DINO_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        e->g_peer[r] = static_cast<float *>(p);
        e->g_ipc[r] = true;
    }
    return DINO_B200_OK;
    DINO_API_END(e)
}

extern "C" dino_b200_status dino_b200_forward_gather_device(dino_b200_engine *e, const float *images, int layout, int B, int H, int W,
                                                            int flags, float *cls, float *patch, float *logits, float *probs, void *stream) {
    if (!e) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (e->g_world == 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "forward_gather: call dino_b200_gather_init first");
    for (int r = 0; r < e->g_world; ++r)
        if (!e->g_peer[r]) throw dino::StatusError(DINO_B200_ERR_INVALID, "forward_gather: the gather buffer of rank " + std::to_string(r) + " has not been registered");
    if (B > e->g_max_batch || H != e->g_H || W != e->g_W)
        throw dino::StatusError(DINO_B200_ERR_INVALID, "forward_gather: batch / image size differ from dino_b200_gather_init");
    DINO_CUDA(cudaSetDevice(e->device));
    dino::ensure_arena(e, B > 0 ? B : 1, H, W);
    dino::forward_device(e, images, layout, B, H, W, (flags & (DINO_B200_CLASSIFY | DINO_B200_FLASH_ATTN_COMPAT)) | dino::kFlagGather, cls, patch, logits, probs,
                         stream ? static_cast<cudaStream_t>(stream) : e->stream);
    return DINO_B200_OK;
    DINO_API_END(e)
}

// ================================================================================================ engine group (one process, n GPUs)
struct dino_b200_group {
    std::vector<dino_b200_engine *> eng;
    int g_what = 0, g_per = 0, g_H = 0, g_W = 0;      // gather configuration currently set up on the engines
    std::string err;
};

extern "C" dino_b200_status dino_b200_group_create_from_gguf(const char *path, const int *devices, int n, dino_b200_group **out) {
    if (!path || !devices || !out || n < 1 || n > 8) {
        dino::g_last_error = "group_create: bad argument (1..8 devices)";
        return DINO_B200_ERR_INVALID;
    }
    *out = nullptr;
    dino_b200_group *g = new (std::nothrow) dino_b200_group();
    if (!g) return DINO_B200_ERR_INVALID;
    for (int i = 0; i < n; ++i) {
        dino_b200_engine *e = nullptr;
        const dino_b200_status st = dino_b200_create_from_gguf(path, devices[i], &e);
        if (st != DINO_B200_OK) {
            for (dino_b200_engine *x : g->eng) dino_b200_destroy(x);
            delete g;
            return st;
        }
        g->eng.push_back(e);
    }
    *out = g;
    return DINO_B200_OK;
}

extern "C" void dino_b200_group_destroy(dino_b200_group *g) {
    if (!g) return;
    for (dino_b200_engine *e : g->eng) dino_b200_destroy(e);
    delete g;
}

extern "C" int dino_b200_group_size(const dino_b200_group *g) { return g ? static_cast<int>(g->eng.size()) : 0; }

extern "C" dino_b200_engine *dino_b200_group_engine(dino_b200_group *g, int i) {
    return (g && i >= 0 && i < static_cast<int>(g->eng.size())) ? g->eng[i] : nullptr;
}

#define DINO_GROUP_END(grp)                                                 \
    }                                                                       \
    catch (const dino::StatusError &ex) {                                   \
        dino::g_last_error = ex.what();                                     \
        return ex.st;                                                       \
    }                                                                       \
    catch (const dino::CudaError &ex) {                                     \
        dino::g_last_error = ex.what();                                     \
        return DINO_B200_ERR_CUDA;                                          \
    }                                                                       \
    catch (const std::exception &ex) {                                      \
        dino::g_last_error = ex.what();                                     \
        return DINO_B200_ERR_INVALID;                                       \
    }

// Data-parallel forward over the group: the global batch is split into contiguous shards (image i -> engine i / ceil(B / n)),
// every device uploads, computes and reads back its shard concurrently; one host thread drives all of them.
extern "C" dino_b200_status dino_b200_group_forward(dino_b200_group *g, const float *images, int layout, int B, int H, int W, int flags,
                                                    float *cls, float *patch, float *logits, float *probs) {
    if (!g || g->eng.empty()) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!images || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "group_forward: bad batch or image size");
    const int n = static_cast<int>(g->eng.size());
    const int per = (B + n - 1) / n;
    const bool classify = (flags & DINO_B200_CLASSIFY) != 0;
    const size_t img_elems = static_cast<size_t>(3) * H * W;
    for (int i = 0; i < n; ++i) {
        dino_b200_engine *e = g->eng[i];
        const int b0 = i * per, nb = std::min(per, B - b0);
        if (nb <= 0) break;
        DINO_CUDA(cudaSetDevice(e->device));
        dino::ensure_arena(e, nb, H, W);
        const int ps = e->hp.patch_size, D = e->hp.hidden_size, C = e->hp.num_classes;
        const size_t np = static_cast<size_t>(H / ps) * (W / ps);
        if (patch && static_cast<size_t>(nb) * np * D > e->cap_o_patch) {
            DINO_CUDA(cudaStreamSynchronize(e->stream));
            if (e->o_patch) DINO_CUDA(cudaFree(e->o_patch));
            e->o_patch = nullptr;
            e->cap_o_patch = 0;
            DINO_CUDA(cudaMalloc(reinterpret_cast<void **>(&e->o_patch), static_cast<size_t>(nb) * np * D * sizeof(float)));
            e->cap_o_patch = static_cast<size_t>(nb) * np * D;
        }
        cudaStream_t st = e->stream;
        DINO_CUDA(cudaMemcpyAsync(e->d_img, images + b0 * img_elems, nb * img_elems * sizeof(float), cudaMemcpyHostToDevice, st));
        dino::forward_device(e, e->d_img, layout, nb, H, W, flags & (DINO_B200_CLASSIFY | DINO_B200_FLASH_ATTN_COMPAT), cls ? e->o_cls : nullptr, patch ? e->o_patch : nullptr,
                             (classify && logits) ? e->logits : nullptr, (classify && probs) ? e->probs : nullptr, st);
        if (cls) DINO_CUDA(cudaMemcpyAsync(cls + static_cast<size_t>(b0) * D, e->o_cls, static_cast<size_t>(nb) * D * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (patch) DINO_CUDA(cudaMemcpyAsync(patch + static_cast<size_t>(b0) * np * D, e->o_patch, static_cast<size_t>(nb) * np * D * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (classify && logits) DINO_CUDA(cudaMemcpyAsync(logits + static_cast<size_t>(b0) * C, e->logits, static_cast<size_t>(nb) * C * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (classify && probs) DINO_CUDA(cudaMemcpyAsync(probs + static_cast<size_t>(b0) * C, e->probs, static_cast<size_t>(nb) * C * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < n; ++i) {
        DINO_CUDA(cudaSetDevice(g->eng[i]->device));
        DINO_CUDA(cudaStreamSynchronize(g->eng[i]->stream));
    }
    return DINO_B200_OK;
    DINO_GROUP_END(g)
}

// Forward + all-gather of the features: afterwards EVERY device of the group holds the [cls] (what = DINO_B200_GATHER_CLS) or
// patch (DINO_B200_GATHER_PATCH) embeddings of the whole batch in its own memory, laid out [n][ceil(B/n)][rows][D] (rank-major,
// unused image slots of the last shard zero).  device_bufs (may be NULL) receives the n device pointers; host_out (may be NULL)
// receives [B][rows][D] copied from the buffer of device `host_from` — any rank serves, that is the point of the gather.
extern "C" dino_b200_status dino_b200_group_allgather_features(dino_b200_group *g, const float *images, int layout, int B, int H, int W,
                                                               int what, float **device_bufs, float *host_out, int host_from) {
    if (!g || g->eng.empty()) return DINO_B200_ERR_INVALID;
    DINO_API_BEGIN
    if (!images || B <= 0 || H <= 0 || W <= 0) throw dino::StatusError(DINO_B200_ERR_INVALID, "group_allgather: bad batch or image size");
    const int n = static_cast<int>(g->eng.size());
    if (host_from < 0 || host_from >= n) throw dino::StatusError(DINO_B200_ERR_INVALID, "group_allgather: host_from is not a rank of the group");
    const int per = (B + n - 1) / n;
    if (g->g_what != what || g->g_per < per || g->g_H != H || g->g_W != W) {
        std::vector<void *> bufs(n, nullptr);
        for (int i = 0; i < n; ++i) {
            const dino_b200_status st = dino_b200_gather_init(g->eng[i], i, n, what, per, H, W, &bufs[i], nullptr);
            if (st != DINO_B200_OK) return st;
        }
        for (int i = 0; i < n; ++i)
            for (int r = 0; r < n; ++r)
                if (r != i) {
                    const dino_b200_status st = dino_b200_gather_set_peer(g->eng[i], r, bufs[r], nullptr);
                    if (st != DINO_B200_OK) return st;
                }
        g->g_what = what;
        g->g_per = per;
        g->g_H = H;
        g->g_W = W;
    }
    const int slot = g->g_per;                               // image slots per rank in the gather buffers
    const size_t img_elems = static_cast<size_t>(3) * H * W;
    for (int i = 0; i < n; ++i) {
        dino_b200_engine *e = g->eng[i];
        const int b0 = i * per, nb = std::min(per, B - b0);
        if (nb <= 0) break;
        DINO_CUDA(cudaSetDevice(e->device));
        dino::ensure_arena(e, nb, H, W);
        DINO_CUDA(cudaMemcpyAsync(e->d_img, images + b0 * img_elems, nb * img_elems * sizeof(float), cudaMemcpyHostToDevice, e->stream));
        dino::forward_device(e, e->d_img, layout, nb, H, W, dino::kFlagGather, nullptr, nullptr, nullptr, nullptr, e->stream);
    }
    // every rank's peer stores have landed once every stream has drained
    for (int i = 0; i < n; ++i) {
        DINO_CUDA(cudaSetDevice(g->eng[i]->device));
        DINO_CUDA(cudaStreamSynchronize(g->eng[i]->stream));
    }
    if (device_bufs)
        for (int i = 0; i < n; ++i) device_bufs[i] = g->eng[i]->g_buf;
    if (host_out) {
        dino_b200_engine *e = g->eng[host_from];
        const size_t row = static_cast<size_t>(e->g_rpi) * e->hp.hidden_size;
        DINO_CUDA(cudaSetDevice(e->device));
        for (int i = 0; i < n; ++i) {
            const int b0 = i * per, nb = std::min(per, B - b0);
            if (nb <= 0) break;
            DINO_CUDA(cudaMemcpyAsync(host_out + static_cast<size_t>(b0) * row, e->g_buf + static_cast<size_t>(i) * slot * row,
                                      static_cast<size_t>(nb) * row * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
        }
        DINO_CUDA(cudaStreamSynchronize(e->stream));
    }
    return DINO_B200_OK;
    DINO_GROUP_END(g)
}
