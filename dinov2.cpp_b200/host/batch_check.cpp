// Link + behaviour check of dino_predict_batch (tests/test_host_dropin.py): loads a model through the reference's own
// dino_model_load signature, runs the same three synthetic images once one by one through dino_predict and once through
// dino_predict_batch, and prints the largest difference.
#include "dinov2_b200_batch.h"
#include "ggml-backend.h"
#include "ggml.h"

#include <opencv2/core.hpp>

#include <cmath>
#include <cstdio>

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    dino_params params;
    params.model = argv[1];
    params.classify = argc > 2 && argv[2][0] == 'c';
    params.topk = 2;
    dino_model model;
    if (!dino_model_load(cv::Size(70, 70), params.model, model, params)) return 1;
    {   // strides of a (possibly quantised) weight as ggml would report them: nb[0] = bytes per block, nb[1] = bytes per row
        const ggml_tensor *w = model.tensors.at("encoder.layer.0.mlp.fc1.weight");
        fprintf(stderr, "batch_check: fc1.weight type %d nb0=%zu nb1=%zu ne0=%lld\n", (int) w->type, (size_t) w->nb[0], (size_t) w->nb[1], (long long) w->ne[0]);
    }
    std::vector<cv::Mat> imgs;
    unsigned s = 12345;
    for (int b = 0; b < 3; ++b) {
        cv::Mat m(70, 70, CV_32FC3);
        for (int i = 0; i < 70 * 70 * 3; ++i) {
            s = s * 1664525u + 1013904223u;
            ((float *) m.data)[i] = ((s >> 8) & 0xFFFF) / 65535.0f * 4.0f - 2.0f;
        }
        imgs.push_back(m);
    }
    auto batch = dino_predict_batch(model, imgs, params);
    if (batch.size() != imgs.size()) return 3;
    double worst = 0;
    for (size_t b = 0; b < imgs.size(); ++b) {
        auto one = dino_predict(model, imgs[b], params, nullptr);
        if (!one) return 4;
        if (params.classify) {
            for (size_t i = 0; i < one->preds->size(); ++i) worst = std::fmax(worst, std::fabs((double) (*one->preds)[i] - (double) (*batch[b]->preds)[i]));
        } else {
            const cv::Mat &x = *one->patch_tokens, &y = *batch[b]->patch_tokens;
            for (int i = 0; i < x.rows * x.cols; ++i) worst = std::fmax(worst, std::fabs((double) ((float *) x.data)[i] - (double) ((float *) y.data)[i]));
        }
    }
    fprintf(stderr, "batch_check: %zu images, max difference %g\n", imgs.size(), worst);
    ggml_backend_free(model.backend);
    return worst == 0 ? 0 : 5;
}
