// Drop-in implementation of the reference's own C++ API (the prototypes of reference dinov2.h:20-118) on top of the
// C ABI in include/dinov2_b200.h.  It is compiled AGAINST THE REFERENCE'S UNMODIFIED dinov2.h (build with
// -I<reference root> -I<reference root>/ggml/include), so the structs (dino_hparams, dino_model, dino_params,
// dino_output) and signatures are, by construction, the reference's; the reference's inference.cpp / realtime.cpp
// compile unchanged against it and link with libdinov2_host.so + libdinov2_b200.so instead of dinov2.cpp + ggml.
//
// What replaces what:
//   dino_model_load    (dinov2.cpp:239-352)  -> own GGUF reader + dino_b200_create; model.tensors still lists every
//                                               checkpoint tensor (host copies), model.backend wraps the engine
//   dino_predict       (dinov2.cpp:900-999)  -> dino_b200_forward (BGR cv::Mat in, probabilities / patch tokens out)
//   dino_preprocess / dino_classify_preprocess / interpolate_pos_embed -> same documented steps on the host
//   the nine ggml symbols the apps call directly (inference.cpp:25,62-73) -> tiny definitions at the end of this file
//   graph builders (build_graph, forward_features, attn, ...) -> not available: the engine has no ggml graph
//
// Deliberate, documented deviations: dino_output::preds holds the top-k CLASS IDS (the reference stores
// `(uint32_t)probability`, i.e. zeros, dinov2.cpp:975); `-o` sets image_out (the reference overwrites fname_inp,
// dinov2.cpp:875-876); -fa (enable_flash_attn) selects DINO_B200_FLASH_ATTN_COMPAT: the phantom zero keys of the reference's
// unmasked flash path are reproduced, the fp16 accumulator noise of ggml's CPU flash kernel is not.
#include "dinov2.h"
#include "ggml-backend.h"
#include "dinov2_b200_batch.h"

#include <opencv2/core.hpp>
#include <opencv2/imgproc.hpp>

#include "../../include/dinov2_b200.h"
#include "../csrc/gguf_reader.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

// ------------------------------------------------------------------------------------------------
// The opaque ggml handle types embedded in dino_model.  Only pointers to them cross the API, so this file is free
// to give them its own definitions.
struct ggml_context {
    dino::GGUFFile file;                       // keeps every tensor's host bytes alive (model.tensors points into it)
    std::vector<ggml_tensor *> tensors;
};
struct ggml_backend {
    dino_b200_engine *engine = nullptr;
};
struct ggml_backend_buffer {
    int unused = 0;
};
struct ggml_backend_buffer_type {
    int unused = 0;
};
struct ggml_gallocr {
    int unused = 0;
};
// the key / value section of a checkpoint as get_val_u32 / get_val_str see it (reference dinov2.h:20-23)
struct gguf_context {
    const dino::GGUFFile *file = nullptr;
};

static dino_b200_engine *engine_of(const dino_model &model) { return model.backend ? model.backend->engine : nullptr; }

// ------------------------------------------------------------------------------------------------ hparams
uint32_t dino_hparams::n_enc_head_dim() const { return hidden_size / num_attention_heads; }
uint32_t dino_hparams::n_img_size() const { return img_size; }
uint32_t dino_hparams::n_patch_size() const { return patch_size; }
uint32_t dino_hparams::n_img_embd() const { return img_size / patch_size; }

// get_val_u32 / get_val_str (dinov2.cpp:55-67): the reference looks the key up with gguf_find_key and reads it; a missing
// key is a fatal invariant there (GGML_ASSERT inside gguf_get_val_*), here it aborts with the same kind of message.
uint32_t get_val_u32(const struct gguf_context *ctx, const char *key) {
    auto it = ctx->file->kv_u.find(key);
    if (it == ctx->file->kv_u.end()) {
        fprintf(stderr, "%s: key '%s' not found in the gguf\n", __func__, key);
        abort();
    }
    return (uint32_t) it->second;
}
const char *get_val_str(const struct gguf_context *ctx, const char *key) {
    auto it = ctx->file->kv_s.find(key);
    if (it == ctx->file->kv_s.end()) {
        fprintf(stderr, "%s: key '%s' not found in the gguf\n", __func__, key);
        abort();
    }
    return it->second.c_str();
}

bool do_quantize(const char *name, const struct ggml_tensor *tensor) {
    // 2-D tensors whose name ends in "weight" (reference PATTERN ".*weight", dinov2.h:18 / dinov2.cpp:227-236)
    const size_t n = std::strlen(name);
    const bool match = n >= 6 && std::strcmp(name + n - 6, "weight") == 0;
    int dims = 4;
    while (dims > 1 && tensor->ne[dims - 1] == 1) --dims;
    return match && dims == 2;
}

// ------------------------------------------------------------------------------------------------ load
bool dino_model_load(const cv::Size img_size, const std::string &fname, dino_model &model, const dino_params &params) {
    (void) img_size;
    printf("%s: loading model from '%s' - please wait\n", __func__, fname.c_str());
    fprintf(stderr, "%s: using the dinov2_b200 engine (sm_100a)\n", __func__);
    auto *ctx = new ggml_context();
    try {
        dino::gguf_read(fname, ctx->file);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s: gguf read failed: %s\n", __func__, e.what());
        delete ctx;
        return false;
    }
    const gguf_context kv{&ctx->file};
    auto u32 = [&](const char *k, uint32_t &dst) {
        if (!ctx->file.kv_u.count(k)) {
            fprintf(stderr, "%s: key '%s' missing from gguf\n", __func__, k);
            return false;
        }
        dst = get_val_u32(&kv, k);
        return true;
    };
    auto &hp = model.hparams;
    if (!u32("hidden_size", hp.hidden_size) || !u32("num_hidden_layers", hp.num_hidden_layers) ||
        !u32("num_attention_heads", hp.num_attention_heads) || !u32("patch_size", hp.patch_size) ||
        !u32("img_size", hp.img_size) || !u32("ftype", hp.ftype) || !u32("num_register_tokens", hp.num_register_tokens)) {
        delete ctx;
        return false;
    }
    const int32_t qntvr = hp.ftype / 1000;   // GGML_QNT_VERSION_FACTOR
    printf("%s: hidden_size            = %d\n", __func__, hp.hidden_size);
    printf("%s: num_hidden_layers      = %d\n", __func__, hp.num_hidden_layers);
    printf("%s: num_register_tokens    = %d\n", __func__, hp.num_register_tokens);
    printf("%s: num_attention_heads    = %d\n", __func__, hp.num_attention_heads);
    printf("%s: patch_size             = %d\n", __func__, hp.patch_size);
    printf("%s: img_size               = %d\n", __func__, hp.img_size);
    printf("%s: ftype                  = %d\n", __func__, hp.ftype);
    printf("%s: qntvr                  = %d\n", __func__, qntvr);
    if (params.classify) {
        if (!u32("num_classes", hp.num_classes)) {
            delete ctx;
            return false;
        }
        printf("%s: num_classes            = %d\n", __func__, hp.num_classes);
        for (uint32_t i = 0; i < hp.num_classes; ++i) {
            auto it = ctx->file.kv_s.find(std::to_string(i));
            hp.id2label[(int) i] = it == ctx->file.kv_s.end() ? std::string() : it->second;
        }
    } else {
        hp.num_classes = 0;          // as the reference: the key is only read for -c (dinov2.cpp:297-304)
    }
    hp.ftype %= 1000;

    // model.tensors: one ggml_tensor per checkpoint tensor, data pointing at the host copy
    // bytes per block of each ggml type (ggml_type_size): F32, F16, Q4_0, Q4_1, -, -, Q5_0, Q5_1, Q8_0
    static const size_t kTypeSize[9] = {4, 2, 18, 20, 0, 0, 22, 24, 34};
    std::vector<dino_b200_tensor> table;
    table.reserve(ctx->file.tensors.size());
    for (size_t i = 0; i < ctx->file.tensors.size(); ++i) {
        const auto &t = ctx->file.tensors[i];
        auto *gt = new ggml_tensor();
        std::memset(gt, 0, sizeof(*gt));
        gt->type = (ggml_type) t.type;
        for (int d = 0; d < 4; ++d) gt->ne[d] = t.ne[d];
        gt->nb[0] = (t.type >= 0 && t.type < 9) ? kTypeSize[t.type] : 0;      // ggml: nb[0] = type size, nb[1] = nb[0] * ne[0] / block
        gt->nb[1] = t.nbytes / std::max<int64_t>(1, t.ne[1] * t.ne[2] * t.ne[3]);
        gt->nb[2] = gt->nb[1] * t.ne[1];
        gt->nb[3] = gt->nb[2] * t.ne[2];
        gt->data = const_cast<uint8_t *>(t.data);
        std::snprintf(gt->name, sizeof(gt->name), "%s", t.name.c_str());
        ctx->tensors.push_back(gt);
        model.tensors[t.name] = gt;
        // the engine uploads the classifier head only when the model is loaded for classification (model.tensors lists it either way)
        if (!params.classify && t.name.rfind("classifier.", 0) == 0) continue;
        table.push_back(dino_b200_tensor{t.name.c_str(), t.type, t.n_dims, {t.ne[0], t.ne[1], t.ne[2], t.ne[3]}, t.data, t.nbytes});
    }
    dino_b200_model_desc desc{};
    desc.hparams = dino_b200_hparams{hp.hidden_size, hp.num_hidden_layers, hp.num_attention_heads, hp.num_classes,
                                     hp.num_register_tokens, hp.patch_size, hp.img_size, hp.ftype, hp.eps};
    desc.n_tensors = (int32_t) table.size();
    desc.tensors = table.data();
    int device = 0;
    if (const char *d = std::getenv("DINO_B200_DEVICE")) device = std::atoi(d);
    dino_b200_engine *eng = nullptr;
    const dino_b200_status st = dino_b200_create(&desc, device, &eng);
    if (st != DINO_B200_OK) {
        fprintf(stderr, "%s: dino_b200_create() failed: %s\n", __func__, dino_b200_last_error(nullptr));
        for (auto *t : ctx->tensors) delete t;
        delete ctx;
        model.tensors.clear();
        return false;
    }
    model.ctx = ctx;
    model.backend = new ggml_backend{eng};
    model.buffer = new ggml_backend_buffer();
    return true;
}

// ------------------------------------------------------------------------------------------------ preprocessing
static cv::Mat standardize_bgr(const cv::Mat &image) {
    // channel order of a cv::Mat is B, G, R; IMAGENET_DEFAULT_MEAN/STD are R, G, B (dinov2.h:16-17)
    cv::Mat out(image.rows, image.cols, CV_32FC3);
    for (int y = 0; y < image.rows; ++y) {
        const float *s = image.ptr<float>(y);
        float *d = out.ptr<float>(y);
        for (int x = 0; x < image.cols; ++x)
            for (int c = 0; c < 3; ++c) d[3 * x + c] = (s[3 * x + c] - IMAGENET_DEFAULT_MEAN[2 - c]) / IMAGENET_DEFAULT_STD[2 - c];
    }
    return out;
}

cv::Mat dino_classify_preprocess(cv::Mat &img, const cv::Size, const dino_hparams &) {
    // [0,1] floats, squash to 256x256 (bicubic), centre-crop 224x224, standardise   (dinov2.cpp:106-132)
    cv::Mat f;
    img.convertTo(f, CV_32FC3, 1.0 / 255.0);
    cv::resize(f, f, cv::Size(256, 256), 0, 0, cv::INTER_CUBIC);
    const int crop = 224;
    cv::Mat roi = f(cv::Rect((f.cols - crop) / 2, (f.rows - crop) / 2, crop, crop));
    return standardize_bgr(roi);
}

cv::Mat dino_preprocess(cv::Mat &img, const cv::Size, const dino_hparams &params) {
    // [0,1] floats, bicubic resize UP to the next patch multiple (even when already a multiple), standardise
    // (dinov2.cpp:135-156)
    cv::Mat f;
    img.convertTo(f, CV_32FC3, 1.0 / 255.0);
    const int ps = (int) params.patch_size;
    cv::resize(f, f, cv::Size((f.cols / ps + 1) * ps, (f.rows / ps + 1) * ps), 0, 0, cv::INTER_CUBIC);
    return standardize_bgr(f);
}

std::vector<float> interpolate_pos_embed(const cv::Size img_size, const float *pos, const dino_hparams &hp) {
    // cls row copied, the M x M grid resampled per channel with cv::resize(INTER_CUBIC); identity when the patch COUNT
    // matches (dinov2.cpp:159-225)
    const int gh = img_size.height / (int) hp.patch_size, gw = img_size.width / (int) hp.patch_size;
    const int M = (int) hp.n_img_embd(), D = (int) hp.hidden_size;
    if (gh * gw == M * M) return std::vector<float>(pos, pos + (size_t) (M * M + 1) * D);
    std::vector<float> out((size_t) (gh * gw + 1) * D);
    std::copy(pos, pos + D, out.begin());
    cv::Mat plane(M, M, CV_32F), resized;
    for (int c = 0; c < D; ++c) {
        for (int i = 0; i < M * M; ++i) plane.at<float>(i / M, i % M) = pos[(size_t) (i + 1) * D + c];
        cv::resize(plane, resized, cv::Size(gw, gh), 0, 0, cv::INTER_CUBIC);
        for (int i = 0; i < gh * gw; ++i) out[(size_t) (i + 1) * D + c] = resized.at<float>(i / gw, i % gw);
    }
    return out;
}

// ------------------------------------------------------------------------------------------------ predict
std::unique_ptr<dino_output> dino_predict(const dino_model &model, const cv::Mat &img, const dino_params &params, ggml_gallocr_t) {
    dino_b200_engine *eng = engine_of(model);
    if (!eng) {
        fprintf(stderr, "%s: model has no dinov2_b200 engine\n", __func__);
        return {};
    }
    if (img.type() != CV_32FC3) {
        fprintf(stderr, "%s: expected a CV_32FC3 image (the output of dino_preprocess)\n", __func__);
        return {};
    }
    cv::Mat contiguous = img.isContinuous() ? img : img.clone();
    const auto &hp = model.hparams;
    const int ps = (int) hp.patch_size;
    // the reference's conv ignores pixels past the last full patch; crop to a patch multiple for the engine
    const int H = img.rows / ps * ps, W = img.cols / ps * ps;
    if (H != img.rows || W != img.cols) contiguous = cv::Mat(contiguous(cv::Rect(0, 0, W, H))).clone();
    auto output = std::make_unique<dino_output>();
    const int fa = params.enable_flash_attn ? DINO_B200_FLASH_ATTN_COMPAT : 0;
    if (params.classify) {
        std::vector<float> probs(hp.num_classes);
        if (dino_b200_forward(eng, (const float *) contiguous.data, DINO_B200_LAYOUT_BGR_HWC, 1, H, W, DINO_B200_CLASSIFY | fa, nullptr,
                              nullptr, nullptr, probs.data()) != DINO_B200_OK) {
            fprintf(stderr, "%s: dino_b200_forward() failed: %s\n", __func__, dino_b200_last_error(eng));
            return {};
        }
        std::vector<int> order(hp.num_classes);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return probs[a] > probs[b]; });
        fprintf(stderr, "\n");
        std::vector<uint32_t> preds(params.topk);
        for (uint32_t i = 0; i < params.topk && i < order.size(); ++i) {
            auto it = hp.id2label.find(order[i]);
            printf(" > %s : %.2f\n", it == hp.id2label.end() ? "?" : it->second.c_str(), probs[order[i]]);
            preds[i] = (uint32_t) order[i];
        }
        output->preds = preds;
    } else {
        const int np = (H / ps) * (W / ps);
        cv::Mat patch_tokens(np, (int) hp.hidden_size, CV_32F);
        if (dino_b200_forward(eng, (const float *) contiguous.data, DINO_B200_LAYOUT_BGR_HWC, 1, H, W, fa, nullptr,
                              (float *) patch_tokens.data, nullptr, nullptr) != DINO_B200_OK) {
            fprintf(stderr, "%s: dino_b200_forward() failed: %s\n", __func__, dino_b200_last_error(eng));
            return {};
        }
        output->patch_tokens = patch_tokens;
    }
    return output;
}

// Batched dino_predict (new surface: the reference is batch 1, dinov2.cpp:630): B preprocessed images of one size in, one
// dino_output per image out, ONE forward pass on the device.  Same per-image semantics as dino_predict (top-k printed and
// returned for -c, patch tokens otherwise).
std::vector<std::unique_ptr<dino_output>> dino_predict_batch(const dino_model &model, const std::vector<cv::Mat> &imgs, const dino_params &params) {
    std::vector<std::unique_ptr<dino_output>> outs;
    dino_b200_engine *eng = engine_of(model);
    if (!eng || imgs.empty()) {
        fprintf(stderr, "%s: no engine or empty batch\n", __func__);
        return outs;
    }
    const auto &hp = model.hparams;
    const int ps = (int) hp.patch_size, B = (int) imgs.size();
    const int H = imgs[0].rows / ps * ps, W = imgs[0].cols / ps * ps;
    std::vector<float> packed((size_t) B * H * W * 3);
    for (int b = 0; b < B; ++b) {
        const cv::Mat &im = imgs[b];
        if (im.type() != CV_32FC3 || im.rows != imgs[0].rows || im.cols != imgs[0].cols) {
            fprintf(stderr, "%s: every image must be CV_32FC3 and of the same size\n", __func__);
            return outs;
        }
        for (int y = 0; y < H; ++y) std::memcpy(&packed[((size_t) b * H + y) * W * 3], im.ptr<float>(y), (size_t) W * 3 * sizeof(float));
    }
    const int np = (H / ps) * (W / ps), D = (int) hp.hidden_size, C = (int) hp.num_classes;
    std::vector<float> probs(params.classify ? (size_t) B * C : 0), patch(params.classify ? 0 : (size_t) B * np * D);
    if (dino_b200_forward(eng, packed.data(), DINO_B200_LAYOUT_BGR_HWC, B, H, W,
                          (params.classify ? DINO_B200_CLASSIFY : 0) | (params.enable_flash_attn ? DINO_B200_FLASH_ATTN_COMPAT : 0), nullptr,
                          params.classify ? nullptr : patch.data(), nullptr, params.classify ? probs.data() : nullptr) != DINO_B200_OK) {
        fprintf(stderr, "%s: dino_b200_forward() failed: %s\n", __func__, dino_b200_last_error(eng));
        return outs;
    }
    for (int b = 0; b < B; ++b) {
        auto o = std::make_unique<dino_output>();
        if (params.classify) {
            const float *p = &probs[(size_t) b * C];
            std::vector<int> order(C);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return p[x] > p[y]; });
            std::vector<uint32_t> preds(params.topk);
            for (uint32_t i = 0; i < params.topk && i < order.size(); ++i) {
                auto it = hp.id2label.find(order[i]);
                printf(" [%d] > %s : %.2f\n", b, it == hp.id2label.end() ? "?" : it->second.c_str(), p[order[i]]);
                preds[i] = (uint32_t) order[i];
            }
            o->preds = preds;
        } else {
            cv::Mat pt(np, D, CV_32F);
            std::memcpy(pt.data, &patch[(size_t) b * np * D], (size_t) np * D * sizeof(float));
            o->patch_tokens = pt;
        }
        outs.push_back(std::move(o));
    }
    return outs;
}

// ------------------------------------------------------------------------------------------------ CLI helpers
void print_usage(int, char **argv, const dino_params &params) {
    fprintf(stderr, "usage: %s [options]\n\noptions:\n", argv[0]);
    fprintf(stderr, "  -h, --help              show this help message and exit\n");
    fprintf(stderr, "  -m FNAME, --model       model path (default: %s)\n", params.model.c_str());
    fprintf(stderr, "  -i FNAME, --inp         input file (default: %s)\n", params.fname_inp.c_str());
    fprintf(stderr, "  -o FNAME, --out         output file for backbone PCA features (default: %s)\n", params.image_out.c_str());
    fprintf(stderr, "  -k N, --topk            top k classes to print (default: %d)\n", params.topk);
    fprintf(stderr, "  -t N, --threads         accepted for compatibility; the GPU engine ignores it (default: %d)\n", params.n_threads);
    fprintf(stderr, "  -c, --classify          classify the image instead of extracting backbone features (default: %d)\n", params.classify);
    fprintf(stderr, "  -fa, --flash_attn       reproduce the reference flash path's unmasked zero-padding keys (default: %d)\n", params.enable_flash_attn);
    fprintf(stderr, "  -cid, --camera_id       camera id for realtime PCA feature streaming (default: %d)\n\n", params.camera_id);
}

bool dino_params_parse(int argc, char **argv, dino_params &params) {
    auto next = [&](int &i) -> const char * {
        if (i + 1 >= argc) {
            fprintf(stderr, "error: missing value for %s\n", argv[i]);
            print_usage(argc, argv, params);
            exit(0);
        }
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-s" || a == "--seed") params.seed = std::stoi(next(i));
        else if (a == "-m" || a == "--model") params.model = next(i);
        else if (a == "-i" || a == "--inp") params.fname_inp = next(i);
        else if (a == "-o" || a == "--out") params.image_out = next(i);
        else if (a == "-t" || a == "--threads") params.n_threads = std::stoi(next(i));
        else if (a == "-k" || a == "--topk") params.topk = std::stoi(next(i));
        else if (a == "-cid" || a == "--camera_id") params.camera_id = (uint8_t) std::stoi(next(i));
        else if (a == "-fa" || a == "--flash_attn") params.enable_flash_attn = true;
        else if (a == "-c" || a == "--classify") params.classify = true;
        else {
            if (a != "-h" && a != "--help") fprintf(stderr, "error: unknown argument: %s\n", a.c_str());
            print_usage(argc, argv, params);
            exit(0);
        }
    }
    return true;
}

void print_t_f32(const char *title, const struct ggml_tensor *t, int n) {
    printf("%s\ndims: %lld %lld %lld %lld f32\n", title, (long long) t->ne[0], (long long) t->ne[1], (long long) t->ne[2], (long long) t->ne[3]);
    const float *d = (const float *) t->data;
    const long long total = (long long) (t->ne[0] * t->ne[1] * t->ne[2] * t->ne[3]);
    double sum = 0;
    for (long long i = 0; i < total; ++i) sum += d[i];
    for (long long i = 0; i < std::min<long long>(n, total); ++i) printf("%.5f ", d[i]);
    printf("\nsum:  %f\n\n", sum);
}

// ---- parts of the reference surface that only exist as ggml graph construction: not provided by this engine ------
static void no_graph(const char *fn) {
    fprintf(stderr, "%s: not available — the dinov2_b200 engine runs a fixed fused pipeline, there is no ggml graph\n", fn);
}
struct ggml_tensor *attn(struct ggml_tensor *, int, struct ggml_context *, const dino_model &, const dino_params &) { no_graph(__func__); return nullptr; }
struct ggml_tensor *mlp(struct ggml_tensor *, int, struct ggml_context *, const dino_model &, const dino_params &) { no_graph(__func__); return nullptr; }
struct ggml_tensor *swiglu_ffn(struct ggml_tensor *, int, struct ggml_context *, const dino_model &, const dino_params &) { no_graph(__func__); return nullptr; }
void forward_features(cv::Size, struct ggml_cgraph *, struct ggml_context *, const dino_model &, const dino_params &) { no_graph(__func__); }
void forward_head(cv::Size, struct ggml_cgraph *, struct ggml_context *, const dino_model &, const dino_params &) { no_graph(__func__); }
struct ggml_cgraph *build_graph(cv::Size, struct ggml_context *, const dino_model &, const dino_params &) { no_graph(__func__); return nullptr; }
bool dino_model_quantize(const std::string &fname_inp, const std::string &fname_out, int itype) {
    // reference dinov2.cpp:354-452, without ggml: same tensor selection, same deterministic quantisers, same file layout
    if (dino_b200_quantize_gguf(fname_inp.c_str(), fname_out.c_str(), itype) != DINO_B200_OK) {
        fprintf(stderr, "%s: %s\n", __func__, dino_b200_last_error(nullptr));
        return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// The nine ggml entry points the reference apps call directly (inference.cpp:25,62-73; realtime.cpp:53,68-72,102-105)
extern "C" {

void ggml_time_init(void) {}
int64_t ggml_time_ms(void) {
    using namespace std::chrono;
    return duration_cast<milliseconds>(steady_clock::now().time_since_epoch()).count();
}
void ggml_backend_synchronize(ggml_backend_t backend) {
    if (backend && backend->engine) dino_b200_synchronize(backend->engine);
}
ggml_backend_buffer_type_t ggml_backend_get_default_buffer_type(ggml_backend_t) {
    static ggml_backend_buffer_type host_type;
    return &host_type;
}
ggml_gallocr_t ggml_gallocr_new(ggml_backend_buffer_type_t) { return new ggml_gallocr(); }
void ggml_gallocr_free(ggml_gallocr_t g) { delete g; }
void ggml_free(struct ggml_context *ctx) {
    if (!ctx) return;
    for (auto *t : ctx->tensors) delete t;
    delete ctx;
}
void ggml_backend_buffer_free(ggml_backend_buffer_t b) { delete b; }
void ggml_backend_free(ggml_backend_t backend) {
    if (!backend) return;
    dino_b200_destroy(backend->engine);
    delete backend;
}

}  // extern "C"
