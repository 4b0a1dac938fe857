// Batched companion of the reference's dino_predict (dinov2.h:111, batch 1 by construction: dinov2.cpp:630) for callers that
// adopt the B200 engine: include it next to the reference's dinov2.h and link libdinov2_host.so.
#pragma once
#include "dinov2.h"

#include <memory>
#include <vector>

// B preprocessed CV_32FC3 images of one size (the output of dino_preprocess / dino_classify_preprocess) -> one dino_output per
// image, computed in ONE forward pass.  Empty vector on error (message on stderr, as dino_predict).
std::vector<std::unique_ptr<dino_output>> dino_predict_batch(const dino_model &model, const std::vector<cv::Mat> &imgs,
                                                             const dino_params &params);
