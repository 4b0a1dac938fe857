/*
 * dinov2_b200.h — C ABI of the B200-native DINOv2 forward engine.
 *
 * This is the drop-in boundary for the reference's hot path: everything the reference does between
 * "weights + a preprocessed image on the host" and "probabilities / patch tokens on the host", i.e.
 * build_graph + ggml_backend_graph_compute inside dino_predict (reference dinov2.cpp:900-948) and all of
 * ggml beneath it.  Plain C types only (pointers + sizes), status-code returns, caller-allocated outputs,
 * no exceptions cross the boundary.  The C++ layer that keeps the reference's own API (dino_model_load /
 * dino_predict ..., reference dinov2.h:94-118) on top of these calls is dinov2.cpp_b200/host/; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Threading: any number of engines per process, on one or several devices; one caller at a time per engine (the reference
 * is single-caller too: dino_predict rebuilds its graph and shares one allocator, dinov2.cpp:907-910).
 */
#ifndef DINOV2_B200_H
#define DINOV2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DINO_B200_API __attribute__((visibility("default")))
#else
#define DINO_B200_API
#endif

typedef struct dino_b200_engine dino_b200_engine;

typedef enum {
    DINO_B200_OK = 0,
    DINO_B200_ERR_INVALID = 1,     /* bad argument (NULL, size not a patch multiple, ...) */
    DINO_B200_ERR_IO = 2,          /* file could not be read */
    DINO_B200_ERR_FORMAT = 3,      /* not a GGUF / missing tensor or key / unsupported tensor type */
    DINO_B200_ERR_CUDA = 4,        /* CUDA runtime/driver failure; see dino_b200_last_error */
    DINO_B200_ERR_UNSUPPORTED = 5, /* model shape outside what the kernels implement (head_dim != 64 ...) */
    DINO_B200_ERR_NO_DEVICE = 6    /* no sm_100 device visible: there is deliberately NO CPU fallback */
} dino_b200_status;

/* Mirrors the numeric fields of the reference's dino_hparams (dinov2.h:25-45). */
typedef struct {
    uint32_t hidden_size;
    uint32_t num_hidden_layers;
    uint32_t num_attention_heads;
    uint32_t num_classes;
    uint32_t num_register_tokens;
    uint32_t patch_size;
    uint32_t img_size;
    uint32_t ftype;
    float eps;
} dino_b200_hparams;

/* ggml_type ids accepted for weights (reference ggml.h enum ggml_type). */
enum { DINO_B200_TYPE_F32 = 0, DINO_B200_TYPE_F16 = 1, DINO_B200_TYPE_Q4_0 = 2, DINO_B200_TYPE_Q4_1 = 3, DINO_B200_TYPE_Q5_0 = 6,
       DINO_B200_TYPE_Q5_1 = 7, DINO_B200_TYPE_Q8_0 = 8 };   /* = ggml_type; the five quantised types of the reference's quantize tool */

/* One checkpoint tensor as the reference holds it in dino_model::tensors (dinov2.h:54): ggml `ne`
 * (fastest dimension first) and a host pointer to the raw bytes. */
typedef struct {
    const char *name;
    int32_t type;
    int32_t n_dims;
    int64_t ne[4];
    const void *data;
    uint64_t nbytes;
} dino_b200_tensor;

typedef struct {
    dino_b200_hparams hparams;
    int32_t n_tensors;
    const dino_b200_tensor *tensors;
} dino_b200_model_desc;

/* image layouts of the `images` argument */
enum {
    DINO_B200_LAYOUT_RGB_PLANAR = 0, /* [B][3][H][W] float32 — what the reference uploads (dinov2.cpp:914-933) */
    DINO_B200_LAYOUT_BGR_HWC = 1     /* [B][H][W][3] float32 — the cv::Mat dino_preprocess returns (dinov2.cpp:135-156) */
};
/* forward flags */
enum {
    DINO_B200_CLASSIFY = 1,          /* dino_params::classify (dinov2.h:63): run forward_head (dinov2.cpp:792-821) */
    /* dino_params::enable_flash_attn (-fa, dinov2.h:65): reproduce the SEMANTIC difference of the reference's flash path
     * (dinov2.cpp:499-525): tokens are zero-padded to a multiple of 32 and ggml_flash_attn_ext runs without a mask, so each
     * query also attends to the padding keys (score 0, value 0) — their weight is added to the softmax denominator.  Not
     * reproduced: the fp16 running accumulator of ggml's CPU flash kernel (ops.cpp:6858-7075) and, for quantised checkpoints,
     * the cast of K / V to the weight type.  Without this flag attention is exact (the reference's default path). */
    DINO_B200_FLASH_ATTN_COMPAT = 2
};

/* Number of usable sm_100 devices (0 when none / no driver). */
DINO_B200_API int dino_b200_device_count(void);

/* Replaces the upload half of dino_model_load (dinov2.cpp:341-348: ggml_backend_alloc_ctx_tensors +
 * ggml_backend_tensor_set): copies every tensor to `device`, converting to the engine's layouts
 * (q8_0 -> fp16, patch-embed K padding, SwiGLU row interleave).  Host data may be freed afterwards. */
DINO_B200_API dino_b200_status dino_b200_create(const dino_b200_model_desc *desc, int device, dino_b200_engine **out);

/* Replaces dino_model_load end to end (dinov2.cpp:239-352: gguf_init_from_file + KV -> hparams + upload)
 * with the engine's own GGUF reader. */
DINO_B200_API dino_b200_status dino_b200_create_from_gguf(const char *path, int device, dino_b200_engine **out);

/* Replaces ggml_backend_free / ggml_backend_buffer_free on the model (inference.cpp:72-73). */
DINO_B200_API void dino_b200_destroy(dino_b200_engine *e);

DINO_B200_API dino_b200_status dino_b200_get_hparams(const dino_b200_engine *e, dino_b200_hparams *out);

/* id2label entry (dino_hparams::id2label, dinov2.cpp:300-304); NULL when the checkpoint has none. */
DINO_B200_API const char *dino_b200_label(const dino_b200_engine *e, int class_id);

/* Pre-sizes the activation arena for up to max_batch images of H x W (replaces ggml_gallocr_new +
 * ggml_gallocr_alloc_graph, inference.cpp:63 / dinov2.cpp:910).  Optional: forward grows it on demand. */
DINO_B200_API dino_b200_status dino_b200_reserve(dino_b200_engine *e, int max_batch, int H, int W);

/* Optional override of the resampled positional embedding for a (gh x gw) patch grid:
 * pos is [(1 + gh*gw)][D] float32 on the host, e.g. the output of the reference's host-side
 * interpolate_pos_embed (dinov2.cpp:159-225, uploaded at :942).  Without it the engine resamples on the
 * device with the same bicubic convention and caches the result per grid. */
DINO_B200_API dino_b200_status dino_b200_set_pos_embed(dino_b200_engine *e, int gh, int gw, const float *pos);

/* Copies the positional embedding the engine would use for an H x W input to the host
 * ([(1 + (H/ps)*(W/ps))][D]); the device counterpart of interpolate_pos_embed. */
DINO_B200_API dino_b200_status dino_b200_get_pos_embed(dino_b200_engine *e, int H, int W, float *out);

/* The hot path: replaces build_graph + ggml_backend_graph_compute + output read-back of dino_predict
 * (dinov2.cpp:907-991) for a batch of B already-preprocessed images in HOST memory.  Synchronous.
 * Any output pointer may be NULL.
 *   cls    [B][D]            final-LayerNorm class token                    (graph node "cls_token", :764-768)
 *   patch  [B][NP][D]        final-LayerNorm patch tokens, registers stripped (node "patch_tokens", :770-789;
 *                            what dino_predict copies into dino_output::patch_tokens, :980-991)
 *   logits [B][num_classes]  classifier output before softmax (requires DINO_B200_CLASSIFY)
 *   probs  [B][num_classes]  softmax(logits)                  (node "probs", :815-818)
 * H and W must be multiples of patch_size (dino_preprocess guarantees it, :140-141). */
DINO_B200_API dino_b200_status dino_b200_forward(dino_b200_engine *e, const float *images, int layout, int B, int H, int W,
                                                 int flags, float *cls, float *patch, float *logits, float *probs);

/* Same with every pointer in DEVICE memory; enqueues on `stream` (a cudaStream_t; NULL = the engine's own
 * stream) and returns without synchronising. */
DINO_B200_API dino_b200_status dino_b200_forward_device(dino_b200_engine *e, const float *images, int layout, int B, int H,
                                                        int W, int flags, float *cls, float *patch, float *logits,
                                                        float *probs, void *stream);

/* Pipelined form of dino_b200_forward for a stream of batches (what realtime.cpp's capture loop needs, realtime.cpp:75-101):
 * submit() enqueues the upload of the batch, its forward pass and the read-back of the requested outputs and returns at once;
 * wait() blocks until the OLDEST submitted batch has completed and its outputs are in place.  Up to two batches may be in
* flight: the upload of batch k+1 and the read-back of batch k-1 (two copy streams) run under the forward pass of batch k.  `images` and the output buffers
 * must stay valid (and should be pinned host memory for the overlap to be real) until the matching wait() returns. */
DINO_B200_API dino_b200_status dino_b200_submit(dino_b200_engine *e, const float *images, int layout, int B, int H, int W,
                                                int flags, float *cls, float *patch, float *logits, float *probs);
DINO_B200_API dino_b200_status dino_b200_wait(dino_b200_engine *e);

/* The realtime caller's whole per-frame loop (reference realtime.cpp:75-101: VideoCapture -> dino_preprocess -> dino_predict ->
 * cv::PCA colouring -> imshow) as ONE asynchronous submission on raw frames: uint8 BGR [B][H][W][3] in host memory (pinned for
 * real overlap) -> upload -> device preprocessing (as dino_b200_preprocess; mode follows DINO_B200_CLASSIFY) -> forward ->
 * optional PCA colouring (as dino_b200_pca_rgb: pca_rgb [B][NP][3] uint8, may be NULL) -> read-back.  Only the frame and the
 * requested results cross PCIe.  Shares the two-slot pipeline of dino_b200_submit: complete each submission with
 * dino_b200_wait.  out_h / out_w (may be NULL) receive the preprocessed size; NP = (out_h / ps) * (out_w / ps). */
DINO_B200_API dino_b200_status dino_b200_submit_u8(dino_b200_engine *e, const uint8_t *frames, int B, int H, int W, int flags,
                                                   float *cls, float *patch, float *logits, float *probs, uint8_t *pca_rgb,
                                                   int *out_h, int *out_w);

/* "Next row" of the hot path: the reference's host-side OpenCV preprocessing on the device.
 * images: B raw frames, uint8 BGR interleaved [B][H][W][3] in HOST memory (what cv::imread / VideoCapture deliver).
 * classify == 0 replaces dino_preprocess (dinov2.cpp:135-156): x/255, bicubic resize UP to the next patch multiple
 *   ((W/ps+1)*ps x (H/ps+1)*ps, even when already a multiple), per-channel (v - mean)/std.
 * classify != 0 replaces dino_classify_preprocess (dinov2.cpp:106-132): x/255, bicubic squash to 256x256, centre crop
 *   224x224, standardise.
 * The result (float32 BGR [B][out_h][out_w][3]) stays on the device for dino_b200_forward_preprocessed and is
 * optionally copied to `out` (host, may be NULL). */
DINO_B200_API dino_b200_status dino_b200_preprocess(dino_b200_engine *e, const uint8_t *images, int B, int H, int W, int classify,
                                                    float *out, int *out_h, int *out_w);

/* dino_preprocess / dino_classify_preprocess + dino_predict in one call on raw uint8 frames (host in, host out):
 * what inference.cpp:48-65 does per image, batched.  Preprocessing mode follows DINO_B200_CLASSIFY in flags. */
DINO_B200_API dino_b200_status dino_b200_forward_u8(dino_b200_engine *e, const uint8_t *images, int B, int H, int W, int flags,
                                                    float *cls, float *patch, float *logits, float *probs);

/* The callers' post-processing: inference.cpp:76-86 and realtime.cpp colour every patch by projecting its
 * feature vector on the top-3 principal components of the image's NP x D patch tokens (cv::PCA DATA_AS_ROW, 3 components;
 * pca.project) and min-max normalising the NP x 3 block to 0..255 (cv::normalize NORM_MINMAX -> CV_8U).  Done here on the
 * device, batched: patch [B][NP][D] float32 -> rgb [B][NP][3] uint8 and/or proj [B][NP][3] float32 (either may be NULL).
 * Components are ordered by decreasing variance; the sign of a component is arbitrary in any PCA — here the largest-
 * magnitude loading of each is positive (OpenCV's solver leaves it as it falls, so a channel may come out inverted).
 * _device: all pointers in device memory, asynchronous on `stream`; the plain form takes and fills host buffers. */
DINO_B200_API dino_b200_status dino_b200_pca_rgb(dino_b200_engine *e, const float *patch, int B, int NP, uint8_t *rgb, float *proj);
DINO_B200_API dino_b200_status dino_b200_pca_rgb_device(dino_b200_engine *e, const float *patch, int B, int NP, uint8_t *rgb,
                                                        float *proj, void *stream);

/* Replaces dino_model_quantize / the `quantize` tool (dinov2.cpp:354-452, quantize.cpp): host-only, needs no GPU.
 * Re-encodes every 2-D "*weight" tensor of a F16/F32 gguf as ggml_type 2 (q4_0), 3 (q4_1), 6 (q5_0), 7 (q5_1) or 8 (q8_0)
 * with the reference's deterministic quantisers and writes the file the reference tool would write. */
DINO_B200_API dino_b200_status dino_b200_quantize_gguf(const char *fname_inp, const char *fname_out, int ggml_type);

/* ---- multi-GPU data parallelism from C / C++ (the "allgather_features over engines[]" entry) ----
 * Images are independent: weights are replicated, the batch is sharded, and the ONLY exchange is the optional all-gather of the
 * per-image features (class token or patch tokens).  It is fused into the last kernel of the forward pass: the final LayerNorm
 * of every rank stores its rows straight into the gather buffer of every rank (peer memory over NVLink / NVSwitch), so there is
 * no separate collective and no staging copy.  The reference has no counterpart (batch 1, one device: dinov2.cpp:630).
 *
 * Low level (one engine = one rank; works across processes through CUDA IPC handles, or inside one process):
 *   gather_init      allocates this rank's gather buffer [world * max_batch][rows][D] fp32 (rows = 1 for DINO_B200_GATHER_CLS,
 *                    NP for DINO_B200_GATHER_PATCH at H x W) and returns its device pointer and/or a 64-byte cudaIpcMemHandle_t
 *   gather_set_peer  registers rank r's buffer: a device pointer of the same process (peer access is enabled) OR an IPC handle
 *   forward_gather_device   dino_b200_forward_device + the fused final-LayerNorm / peer-store all-gather: this rank's rows land
 *                    in slot [rank * max_batch + image] of EVERY registered buffer.  Asynchronous; the rows of the other ranks
 *                    are complete in the local buffer once every rank's stream has been synchronised (caller's barrier).
 *                    cls / patch / logits / probs (device, may be NULL) are the usual local outputs. */
enum { DINO_B200_GATHER_CLS = 1, DINO_B200_GATHER_PATCH = 2 };
#define DINO_B200_IPC_HANDLE_BYTES 64
DINO_B200_API dino_b200_status dino_b200_gather_init(dino_b200_engine *e, int rank, int world, int what, int max_batch, int H, int W,
                                                     void **local_buf, unsigned char *ipc_handle);
DINO_B200_API dino_b200_status dino_b200_gather_set_peer(dino_b200_engine *e, int r, void *dev_ptr, const unsigned char *ipc_handle);
DINO_B200_API dino_b200_status dino_b200_forward_gather_device(dino_b200_engine *e, const float *images, int layout, int B, int H,
                                                               int W, int flags, float *cls, float *patch, float *logits,
                                                               float *probs, void *stream);

/* High level: n engines on n devices of ONE process (the host code stays C / C++, no torch, no NCCL).
 *   group_forward             dino_b200_forward over a global batch: image i runs on engine i / ceil(B / n); uploads, forward
 *                             passes and read-backs of all devices overlap; host in, host out (pin the buffers for real overlap)
 *   group_allgather_features  forward + fused all-gather: every device ends up with the features of the WHOLE batch, laid out
 *                             [n][ceil(B/n)][rows][D]; device_bufs (NULL or n pointers) receives the per-device buffers,
 *                             host_out (NULL or [B][rows][D]) a copy taken from device `host_from` */
typedef struct dino_b200_group dino_b200_group;
DINO_B200_API dino_b200_status dino_b200_group_create_from_gguf(const char *path, const int *devices, int n, dino_b200_group **out);
DINO_B200_API void dino_b200_group_destroy(dino_b200_group *g);
DINO_B200_API int dino_b200_group_size(const dino_b200_group *g);
DINO_B200_API dino_b200_engine *dino_b200_group_engine(dino_b200_group *g, int i);
DINO_B200_API dino_b200_status dino_b200_group_forward(dino_b200_group *g, const float *images, int layout, int B, int H, int W,
                                                       int flags, float *cls, float *patch, float *logits, float *probs);
DINO_B200_API dino_b200_status dino_b200_group_allgather_features(dino_b200_group *g, const float *images, int layout, int B, int H,
                                                                  int W, int what, float **device_bufs, float *host_out,
                                                                  int host_from);

/* Replaces ggml_backend_synchronize (inference.cpp:62,66). */
DINO_B200_API dino_b200_status dino_b200_synchronize(dino_b200_engine *e);

/* Last error text of the engine (or of the failed create call when e == NULL). Never NULL. */
DINO_B200_API const char *dino_b200_last_error(const dino_b200_engine *e);

/* Number of engine kernels launched since creation (bench.py reports the per-step delta as gpu_launches). */
DINO_B200_API uint64_t dino_b200_kernel_launches(const dino_b200_engine *e);

/* cudaEvent-timed duration (ms) of the GEMM kernels / attention kernels of the most recent
 * dino_b200_forward* call when profiling is switched on (off by default; adds event records). */
DINO_B200_API dino_b200_status dino_b200_set_profiling(dino_b200_engine *e, int on);
DINO_B200_API dino_b200_status dino_b200_get_profile(dino_b200_engine *e, float *gemm_ms, float *attn_ms, float *other_ms,
                                                     float *total_ms);

/* ---- kernel-level entry points (device pointers), used by tests/ and bench.py to check and time each
 * hand-written kernel in isolation against a plain fp32 restatement. stream may be NULL. ---- */
enum { DINO_B200_EPI_BIAS_F16 = 0, DINO_B200_EPI_GELU_F16 = 1, DINO_B200_EPI_RESID_F32 = 2, DINO_B200_EPI_SWIGLU_F16 = 3,
       DINO_B200_EPI_PATCH_F32 = 4 };
/* C = A[M,K](fp16) x W[N,K]^T(fp16) with fused epilogue `epi` (see csrc/gemm.cuh). lda/ldw in elements.
 * out: fp16 [M, ldo] for BIAS/GELU/SWIGLU (SWIGLU writes N/2 columns; W rows and bias pre-interleaved
 * 128 gate | 128 up), fp32 [.., ldo] updated in place for RESID, written for PATCH. */
DINO_B200_API dino_b200_status dino_b200_kernel_gemm(int epi, const void *A, int lda, const void *W, int ldw, int M, int N, int K,
                                                     const float *bias, const float *lscale, void *out, int ldo,
                                                     const float *pos, int np, int ntok, int tok_off, void *stream);
/* The fused residual + LayerNorm GEMM (reference dinov2.cpp:546-551 + 708-714 + 722-728):
 * X[M,N](fp32) += lscale * (A x W^T + bias), then ln_out[M,N](fp16) = LayerNorm(X row; eps) * gamma + beta, written by the
 * kernel's LayerNorm worker warps as soon as a 128-row block of X is complete (bit-identical to dino_b200_kernel_layernorm on
 * the updated X).  counters: 2 * ceil(M/128) + 2 ints, zero on entry, zero again on exit.  N must be 384, 768, 1024 or 1536. */
DINO_B200_API dino_b200_status dino_b200_kernel_gemm_resid_ln(const void *A, int lda, const void *W, int ldw, int M, int N, int K,
                                                              const float *bias, const float *lscale, float *X,
                                                              const float *gamma, const float *beta, float eps, void *ln_out,
                                                              int *counters, void *stream);
/* out[B*N, D](fp16) = MHA(qkv[B*N, 3D](fp16)), head_dim 64 */
DINO_B200_API dino_b200_status dino_b200_kernel_attention(const void *qkv, void *out, int B, int n_tok, int D, void *stream);
/* LayerNorm rows of X[rows, D] -> fp16 (out_half != 0) or fp32 */
DINO_B200_API dino_b200_status dino_b200_kernel_layernorm(const float *X, const float *gamma, const float *beta, void *out,
                                                          int rows, int D, float eps, int out_half, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DINOV2_B200_H */
