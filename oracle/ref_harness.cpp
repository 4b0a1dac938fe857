// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" wrapper around the *unmodified* reference implementation
// (reference dinov2.cpp compiled from /root/reference by oracle/Makefile) so
// that the parity tests and bench.py's cpu_baseline / --impl reference legs
// can drive it from Python through ctypes.  Nothing here re-implements the
// forward pass: ref_forward() performs exactly the steps of the reference's
// dino_predict (dinov2.cpp:900-948 — build_graph, gallocr, BGR->RGB planar,
// interpolate_pos_embed, ggml_backend_graph_compute) by calling the
// reference's own functions, and then reads every graph output instead of
// only the printed top-k, because dino_predict discards logits/cls and stores
// `(uint32_t)prob` into preds (dinov2.cpp:975).  ref_predict() calls the real
// dino_predict for a cross-check of the features path.
#include "dinov2.h"
#include "ggml-backend.h"
#include "ggml-cpu.h"
#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

#include <chrono>
#include <cstdio>
#include <cstring>

struct ref_handle {
    dino_model model;
    dino_params params;
    ggml_gallocr_t allocr = nullptr;
};

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

extern "C" {

void *ref_load(const char *gguf_path, int n_threads, int classify, int flash_attn, int H, int W) {
    auto *h = new ref_handle();
    h->params.model = gguf_path;
    h->params.n_threads = n_threads;
    h->params.classify = classify != 0;
    h->params.enable_flash_attn = flash_attn != 0;
    if (!dino_model_load(cv::Size(W, H), gguf_path, h->model, h->params)) {
        delete h;
        return nullptr;
    }
    h->allocr = ggml_gallocr_new(ggml_backend_get_default_buffer_type(h->model.backend));
    return h;
}

void ref_hparams(void *hv, uint32_t out[8]) {
    auto *h = (ref_handle *) hv;
    const auto &p = h->model.hparams;
    out[0] = p.hidden_size; out[1] = p.num_hidden_layers; out[2] = p.num_attention_heads; out[3] = p.num_classes;
    out[4] = p.num_register_tokens; out[5] = p.patch_size; out[6] = p.img_size; out[7] = p.ftype;
}

// img: H x W x 3 float32, BGR interleaved (what dino_preprocess hands to dino_predict).
// Outputs may be NULL. Returns wall milliseconds of graph build + compute, <0 on failure.
//   cls    [D]              final-LN cls token              (graph node "cls_token")
//   patch  [n_patch_out, D] graph node "patch_tokens" (features mode: registers stripped;
//                           classify mode: includes registers, dinov2.cpp:770-776)
//   logits [num_classes]    input of the final soft_max (probs->src[0])
//   probs  [num_classes]
double ref_forward(void *hv, const float *img, int H, int W, float *cls, float *patch, float *logits, float *probs) {
    auto *h = (ref_handle *) hv;
    const dino_model &model = h->model;
    const double t0 = now_ms();

    struct ggml_init_params ip = {ggml_tensor_overhead() * GGML_DEFAULT_GRAPH_SIZE + ggml_graph_overhead(), nullptr, true};
    struct ggml_context *ctx = ggml_init(ip);
    struct ggml_cgraph *gf = build_graph(cv::Size(W, H), ctx, model, h->params);
    ggml_gallocr_alloc_graph(h->allocr, gf);

    const size_t plane = (size_t) H * W;
    std::vector<float> planar(plane * 3);
    for (size_t i = 0; i < plane; ++i) {           // BGR interleaved -> RGB planar (dinov2.cpp:914-931)
        planar[0 * plane + i] = img[i * 3 + 2];
        planar[1 * plane + i] = img[i * 3 + 1];
        planar[2 * plane + i] = img[i * 3 + 0];
    }
    struct ggml_tensor *input = ggml_graph_get_tensor(gf, "input");
    ggml_backend_tensor_set(input, planar.data(), 0, ggml_nbytes(input));

    const struct ggml_tensor *pos = ggml_get_tensor(model.ctx, "embeddings.position_embeddings");
    const std::vector<float> pos_fixed = interpolate_pos_embed(cv::Size(W, H), (const float *) pos->data, model.hparams);
    struct ggml_tensor *pos_t = ggml_graph_get_tensor(gf, "pos_embed_fixed");
    ggml_backend_tensor_set(pos_t, pos_fixed.data(), 0, ggml_nbytes(pos_t));

    if (ggml_backend_graph_compute(model.backend, gf) != GGML_STATUS_SUCCESS) {
        ggml_free(ctx);
        return -1.0;
    }
    const double t1 = now_ms();

    if (cls) {
        struct ggml_tensor *t = ggml_graph_get_tensor(gf, "cls_token");
        std::memcpy(cls, t->data, ggml_nbytes(t));
    }
    if (patch) {
        struct ggml_tensor *t = ggml_graph_get_tensor(gf, "patch_tokens");
        // a view with contiguous rows: ne0 = D, ne1 = tokens
        std::memcpy(patch, t->data, (size_t) t->ne[0] * t->ne[1] * sizeof(float));
    }
    if (h->params.classify) {
        struct ggml_tensor *p = ggml_graph_get_tensor(gf, "probs");
        if (probs) std::memcpy(probs, p->data, ggml_nbytes(p));
        if (logits) std::memcpy(logits, p->src[0]->data, ggml_nbytes(p->src[0]));
    }
    ggml_free(ctx);
    return t1 - t0;
}

// The reference's real entry point, features mode only (classify mode only prints).
// patch: [ (H/ps)*(W/ps), D ].  Returns wall ms of dino_predict, <0 on failure.
double ref_predict(void *hv, const float *img, int H, int W, float *patch) {
    auto *h = (ref_handle *) hv;
    cv::Mat m(H, W, CV_32FC3, (void *) img);
    const double t0 = now_ms();
    std::unique_ptr<dino_output> out = dino_predict(h->model, m, h->params, h->allocr);
    const double t1 = now_ms();
    if (!out) return -1.0;
    if (patch && out->patch_tokens.has_value()) {
        const cv::Mat &pt = out->patch_tokens.value();
        std::memcpy(patch, pt.data, (size_t) pt.rows * pt.cols * sizeof(float));
    }
    return t1 - t0;
}

// reference host-side helpers, exposed for the "next rows" parity tests
int ref_interpolate_pos_embed(void *hv, int H, int W, float *out, int64_t out_cap) {
    auto *h = (ref_handle *) hv;
    const struct ggml_tensor *pos = ggml_get_tensor(h->model.ctx, "embeddings.position_embeddings");
    const std::vector<float> v = interpolate_pos_embed(cv::Size(W, H), (const float *) pos->data, h->model.hparams);
    if ((int64_t) v.size() > out_cap) return -1;
    std::memcpy(out, v.data(), v.size() * sizeof(float));
    return (int) v.size();
}

// u8 BGR HWC image -> dino_preprocess / dino_classify_preprocess; returns out H,W via pointers
int ref_preprocess(void *hv, const uint8_t *bgr, int H, int W, int classify, float *out, int64_t out_cap, int *oh, int *ow) {
    auto *h = (ref_handle *) hv;
    cv::Mat m(H, W, CV_8UC3, (void *) bgr);
    cv::Mat r = classify ? dino_classify_preprocess(m, cv::Size(W, H), h->model.hparams)
                         : dino_preprocess(m, cv::Size(W, H), h->model.hparams);
    *oh = r.rows; *ow = r.cols;
    const int64_t n = (int64_t) r.rows * r.cols * 3;
    if (n > out_cap) return -1;
    for (int y = 0; y < r.rows; ++y) std::memcpy(out + (size_t) y * r.cols * 3, r.ptr<float>(y), (size_t) r.cols * 3 * sizeof(float));
    return 0;
}

void ref_free(void *hv) {
    auto *h = (ref_handle *) hv;
    if (!h) return;
    ggml_gallocr_free(h->allocr);
    ggml_free(h->model.ctx);
    ggml_backend_buffer_free(h->model.buffer);
    ggml_backend_free(h->model.backend);
    delete h;
}

}  // extern "C"
