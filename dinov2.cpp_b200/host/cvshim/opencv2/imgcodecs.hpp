// cv::imread / cv::imwrite for the shim (see core.hpp).  Decoding/encoding is delegated to the stb single-header
// libraries that the reference tree already ships and that its apps include themselves
// (reference inference.cpp:8 includes "ggml/examples/stb_image.h"); they are found on the include path at build
// time (-I<reference root>) and are not copied into this repo.
#pragma once
#include "core.hpp"

#ifndef STB_IMAGE_IMPLEMENTATION
#define STB_IMAGE_IMPLEMENTATION
#endif
#include "ggml/examples/stb_image.h"
#undef STB_IMAGE_IMPLEMENTATION   // the apps include stb_image.h again; its implementation part has no include guard
#ifndef STB_IMAGE_WRITE_IMPLEMENTATION
#define STB_IMAGE_WRITE_IMPLEMENTATION
#endif
#include "ggml/examples/stb_image_write.h"
#undef STB_IMAGE_WRITE_IMPLEMENTATION

namespace cv {

inline Mat imread(const std::string &path, int /*flags*/ = IMREAD_COLOR) {
    int w = 0, h = 0, c = 0;
    unsigned char *rgb = stbi_load(path.c_str(), &w, &h, &c, 3);
    if (!rgb) return Mat();
    Mat out(h, w, CV_8UC3);
    for (int y = 0; y < h; ++y) {
        uint8_t *d = out.ptr<uint8_t>(y);
        const unsigned char *s = rgb + (size_t) y * w * 3;
        for (int x = 0; x < w; ++x) { d[3 * x] = s[3 * x + 2]; d[3 * x + 1] = s[3 * x + 1]; d[3 * x + 2] = s[3 * x]; }   // RGB -> BGR
    }
    stbi_image_free(rgb);
    return out;
}

inline bool imwrite(const std::string &path, const Mat &img) {
    if (img.empty() || img.depth() != CV_8U) return false;
    const int cn = img.channels();
    std::vector<unsigned char> buf((size_t) img.rows * img.cols * cn);
    for (int y = 0; y < img.rows; ++y) {
        const uint8_t *s = img.ptr<uint8_t>(y);
        unsigned char *d = buf.data() + (size_t) y * img.cols * cn;
        for (int x = 0; x < img.cols; ++x)
            for (int k = 0; k < cn; ++k) d[x * cn + k] = s[x * cn + (cn == 3 ? 2 - k : k)];   // BGR -> RGB
    }
    auto ends_with = [&](const char *e) { const size_t n = std::strlen(e); return path.size() >= n && path.compare(path.size() - n, n, e) == 0; };
    if (ends_with(".png")) return stbi_write_png(path.c_str(), img.cols, img.rows, cn, buf.data(), img.cols * cn) != 0;
    if (ends_with(".bmp")) return stbi_write_bmp(path.c_str(), img.cols, img.rows, cn, buf.data()) != 0;
    return stbi_write_jpg(path.c_str(), img.cols, img.rows, cn, buf.data(), 95) != 0;
}

}  // namespace cv
