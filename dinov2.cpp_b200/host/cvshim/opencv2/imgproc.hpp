// cv::resize for the three interpolation modes the dinov2.cpp sources use
// (INTER_CUBIC in dino_preprocess / interpolate_pos_embed, reference
// dinov2.cpp:112-144,210; INTER_NEAREST in the apps).  See core.hpp for why
// this shim exists.  Follows OpenCV's documented sampling convention:
// src_x = (dst_x + 0.5) * (src_w / dst_w) - 0.5, bicubic kernel a = -0.75,
// replicated borders, separable float passes (horizontal then vertical).
#pragma once
#include "core.hpp"

namespace cv {

namespace shim_detail {
inline void cubic_coeffs(float x, float c[4]) {
    const float A = -0.75f;
    c[0] = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
    c[1] = ((A + 2) * x - (A + 3)) * x * x + 1;
    c[2] = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
    c[3] = 1.f - c[0] - c[1] - c[2];
}
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
}  // namespace shim_detail

inline void resize(const Mat &src_in, Mat &dst, Size dsize, double /*fx*/ = 0, double /*fy*/ = 0,
                   int interpolation = INTER_LINEAR) {
    using namespace shim_detail;
    Mat src = src_in;  // keeps the storage alive when &src_in == &dst
    const int cn = src.channels();
    const int sw = src.cols, sh = src.rows, dw = dsize.width, dh = dsize.height;
    if (sw == dw && sh == dh) { Mat o; src.copyTo(o); dst = o; return; }
    const double scale_x = (double) sw / dw, scale_y = (double) sh / dh;
    Mat out(dh, dw, src.type());

    if (interpolation == INTER_NEAREST) {
        const size_t es = src.elemSize();
        for (int y = 0; y < dh; ++y) {
            const int sy = std::min((int) std::floor(y * scale_y), sh - 1);
            for (int x = 0; x < dw; ++x) {
                const int sx = std::min((int) std::floor(x * scale_x), sw - 1);
                std::memcpy(out.data + (size_t) y * out.step + x * es, src.data + (size_t) sy * src.step + sx * es, es);
            }
        }
        dst = out;
        return;
    }

    const int taps = interpolation == INTER_CUBIC ? 4 : 2;
    std::vector<int> xofs((size_t) dw * taps), yofs((size_t) dh * taps);
    std::vector<float> xw((size_t) dw * taps), yw((size_t) dh * taps);
    auto build = [&](int dn, int sn, double scale, std::vector<int> &ofs, std::vector<float> &w) {
        for (int d = 0; d < dn; ++d) {
            float f = (float) ((d + 0.5) * scale - 0.5);
            int s = (int) std::floor(f);
            f -= s;
            if (taps == 4) {
                float c[4]; cubic_coeffs(f, c);
                for (int k = 0; k < 4; ++k) { ofs[(size_t) d * 4 + k] = clampi(s - 1 + k, 0, sn - 1); w[(size_t) d * 4 + k] = c[k]; }
            } else {
                if (s < 0) { s = 0; f = 0; }
                if (s >= sn - 1) { s = sn - 1; f = 0; }
                ofs[(size_t) d * 2] = s; ofs[(size_t) d * 2 + 1] = clampi(s + 1, 0, sn - 1);
                w[(size_t) d * 2] = 1.f - f; w[(size_t) d * 2 + 1] = f;
            }
        }
    };
    build(dw, sw, scale_x, xofs, xw);
    build(dh, sh, scale_y, yofs, yw);

    const bool is_f32 = src.depth() == CV_32F;
    // horizontal pass into a float image [sh][dw*cn]
    std::vector<float> tmp((size_t) sh * dw * cn);
    for (int y = 0; y < sh; ++y) {
        float *t = &tmp[(size_t) y * dw * cn];
        for (int x = 0; x < dw; ++x)
            for (int c = 0; c < cn; ++c) {
                float acc = 0.f;
                for (int k = 0; k < taps; ++k) {
                    const int sx = xofs[(size_t) x * taps + k];
                    const float v = is_f32 ? src.ptr<float>(y)[sx * cn + c] : (float) src.ptr<uint8_t>(y)[sx * cn + c];
                    acc = k == 0 ? v * xw[(size_t) x * taps] : acc + v * xw[(size_t) x * taps + k];
                }
                t[x * cn + c] = acc;
            }
    }
    // vertical pass
    for (int y = 0; y < dh; ++y) {
        for (int i = 0; i < dw * cn; ++i) {
            float acc = 0.f;
            for (int k = 0; k < taps; ++k) {
                const float v = tmp[(size_t) yofs[(size_t) y * taps + k] * dw * cn + i];
                acc = k == 0 ? v * yw[(size_t) y * taps] : acc + v * yw[(size_t) y * taps + k];
            }
            if (is_f32) out.ptr<float>(y)[i] = acc;
            else out.ptr<uint8_t>(y)[i] = (uint8_t) clampi((int) std::nearbyint(acc), 0, 255);
        }
    }
    dst = out;
}

}  // namespace cv
