// Minimal header-only stand-in for the slice of OpenCV's C++ API that the
// dinov2.cpp sources touch (cv::Mat / Size / Rect, split / merge, convertTo,
// ROI views, scalar arithmetic, normalize, PCA, hconcat).
//
// Why it exists: neither this build image nor the B200 boxes carry the OpenCV
// C++ SDK (only the Python `cv2` wheel).  The reference's dinov2.h embeds
// cv::Mat / cv::Size in its public signatures (reference dinov2.h:94-111), so
// both the oracle harness (oracle/Makefile, compiling the unmodified reference
// dinov2.cpp) and the drop-in host layer (host/dinov2_host.cpp) need *some*
// definition of those types.  When a real OpenCV is installed, point the build
// at it instead (-I<opencv>/include4) and drop this directory from the include
// path: nothing in the engine depends on shim internals.
//
// Written from scratch against OpenCV's documented behaviour; numerics of
// resize(INTER_CUBIC) are cross-checked against Python cv2 in
// tests/test_cvshim.py.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 63) + 1)

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() = default;
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
    int area() const { return width * height; }
};

struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() = default;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };
enum NormTypes { NORM_MINMAX = 32 };
enum ImreadModes { IMREAD_COLOR = 1 };

class Mat;

// cv::Mat::size is an object that is both callable and indexable.
struct MatSize {
    const Mat *m = nullptr;
    Size operator()() const;
    int operator[](int i) const;
};

class Mat {
public:
    int rows = 0, cols = 0;
    int flags_type = CV_8UC1;
    size_t step = 0;  // bytes per row
    uint8_t *data = nullptr;
    MatSize size;

    Mat() { size.m = this; }
    Mat(int r, int c, int type) { size.m = this; create(r, c, type); }
    Mat(Size s, int type) { size.m = this; create(s.height, s.width, type); }
    // borrowed storage (no ownership)
    Mat(int r, int c, int type, void *ext, size_t step_ = 0) {
        size.m = this;
        rows = r; cols = c; flags_type = type;
        step = step_ ? step_ : (size_t) c * elemSize();
        data = (uint8_t *) ext;
    }
    Mat(const Mat &o) { size.m = this; assign(o); }
    Mat &operator=(const Mat &o) { if (this != &o) assign(o); return *this; }

    void create(int r, int c, int type) {
        if (data && owner && rows == r && cols == c && flags_type == type && step == (size_t) c * elemSize()) return;
        rows = r; cols = c; flags_type = type;
        step = (size_t) c * elemSize();
        size_t bytes = step * (size_t) r;
        owner = std::shared_ptr<uint8_t>(new uint8_t[bytes ? bytes : 1], std::default_delete<uint8_t[]>());
        data = owner.get();
    }

    int type() const { return flags_type; }
    int depth() const { return CV_MAT_DEPTH(flags_type); }
    int channels() const { return CV_MAT_CN(flags_type); }
    size_t elemSize1() const { return depth() == CV_32F ? 4 : 1; }
    size_t elemSize() const { return elemSize1() * channels(); }
    size_t total() const { return (size_t) rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t) cols * elemSize(); }

    template <typename T> T *ptr(int y = 0) { return (T *) (data + (size_t) y * step); }
    template <typename T> const T *ptr(int y = 0) const { return (const T *) (data + (size_t) y * step); }
    template <typename T> T &at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T &at(int y, int x) const { return ptr<T>(y)[x]; }

    Mat operator()(const Rect &r) const {
        Mat v;
        v.rows = r.height; v.cols = r.width; v.flags_type = flags_type; v.step = step;
        v.data = data + (size_t) r.y * step + (size_t) r.x * elemSize();
        v.owner = owner;
        return v;
    }

    Mat clone() const { Mat o; copyTo(o); return o; }

    void copyTo(Mat &dst) const {
        if (dst.data == nullptr || dst.rows != rows || dst.cols != cols || dst.flags_type != flags_type)
            dst.create(rows, cols, flags_type);
        const size_t rowbytes = (size_t) cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t) y * dst.step, data + (size_t) y * step, rowbytes);
    }
    // cv::split hands out temporaries that get copied into pre-bound planes
    void copyTo(Mat &&dst) const { copyTo(dst); }

    void convertTo(Mat &dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
        const int cn = channels();
        const int ddepth = CV_MAT_DEPTH(rtype);
        Mat out(rows, cols, CV_MAKETYPE(ddepth, cn));
        const int n = cols * cn;
        for (int y = 0; y < rows; ++y) {
            for (int i = 0; i < n; ++i) {
                double v = depth() == CV_32F ? (double) ptr<float>(y)[i] : (double) ptr<uint8_t>(y)[i];
                if (ddepth == CV_32F) {
                    // OpenCV's 8u->32f path evaluates src*alpha+beta in float
                    out.ptr<float>(y)[i] = depth() == CV_32F ? (float) (v * alpha + beta)
                                                            : (float) v * (float) alpha + (float) beta;
                } else {
                    double r = std::nearbyint(v * alpha + beta);
                    out.ptr<uint8_t>(y)[i] = (uint8_t) std::min(255.0, std::max(0.0, r));
                }
            }
        }
        dst = out;
    }

    // reshape(cn, rows): continuous data only
    Mat reshape(int cn, int new_rows) const {
        if (!isContinuous()) throw std::runtime_error("cvshim: reshape needs continuous data");
        const size_t scalars = total() * channels();
        Mat v;
        v.flags_type = CV_MAKETYPE(depth(), cn);
        v.rows = new_rows;
        v.cols = (int) (scalars / ((size_t) cn * new_rows));
        v.step = (size_t) v.cols * v.elemSize();
        v.data = data; v.owner = owner;
        return v;
    }

private:
    std::shared_ptr<uint8_t> owner;
    void assign(const Mat &o) {
        rows = o.rows; cols = o.cols; flags_type = o.flags_type; step = o.step; data = o.data; owner = o.owner;
    }
    friend Mat operator-(const Mat &, float);
    friend Mat operator/(const Mat &, float);
};

inline Size MatSize::operator()() const { return Size(m->cols, m->rows); }
inline int MatSize::operator[](int i) const { return i == 0 ? m->rows : m->cols; }

// OpenCV evaluates Mat-scalar expressions per element in the Mat's depth;
// for CV_32F this is a plain float subtract / divide.
inline Mat operator-(const Mat &a, float s) {
    Mat o(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    for (int y = 0; y < a.rows; ++y)
        for (int i = 0; i < n; ++i) o.ptr<float>(y)[i] = a.ptr<float>(y)[i] - s;
    return o;
}
inline Mat operator/(const Mat &a, float s) {
    // OpenCV lowers Mat / scalar to a multiply by the reciprocal (MatExpr scale = 1/s, kept in double)
    Mat o(a.rows, a.cols, a.type());
    const int n = a.cols * a.channels();
    const double inv = 1.0 / (double) s;
    for (int y = 0; y < a.rows; ++y)
        for (int i = 0; i < n; ++i) o.ptr<float>(y)[i] = (float) ((double) a.ptr<float>(y)[i] * inv);
    return o;
}

inline void split(const Mat &src, std::vector<Mat> &planes) {
    const int cn = src.channels();
    planes.resize(cn);
    for (int c = 0; c < cn; ++c) planes[c] = Mat(src.rows, src.cols, CV_MAKETYPE(src.depth(), 1));
    const size_t es = src.elemSize1();
    for (int y = 0; y < src.rows; ++y) {
        const uint8_t *s = src.ptr<uint8_t>(y);
        for (int x = 0; x < src.cols; ++x)
            for (int c = 0; c < cn; ++c)
                std::memcpy(planes[c].data + (size_t) y * planes[c].step + x * es, s + ((size_t) x * cn + c) * es, es);
    }
}

inline void merge(const std::vector<Mat> &planes, Mat &dst) {
    const int cn = (int) planes.size();
    const Mat &p0 = planes[0];
    Mat out(p0.rows, p0.cols, CV_MAKETYPE(p0.depth(), cn));
    const size_t es = p0.elemSize1();
    for (int y = 0; y < p0.rows; ++y) {
        uint8_t *d = out.ptr<uint8_t>(y);
        for (int x = 0; x < p0.cols; ++x)
            for (int c = 0; c < cn; ++c)
                std::memcpy(d + ((size_t) x * cn + c) * es, planes[c].data + (size_t) y * planes[c].step + x * es, es);
    }
    dst = out;
}

inline void hconcat(const std::vector<Mat> &src, Mat &dst) {
    int cols = 0;
    for (auto &m : src) cols += m.cols;
    Mat out(src[0].rows, cols, src[0].type());
    int x0 = 0;
    for (auto &m : src) {
        for (int y = 0; y < m.rows; ++y)
            std::memcpy(out.data + (size_t) y * out.step + (size_t) x0 * out.elemSize(), m.ptr<uint8_t>(y),
                        (size_t) m.cols * m.elemSize());
        x0 += m.cols;
    }
    dst = out;
}

// normalize(NORM_MINMAX) to [a,b] with optional depth change
inline void normalize(const Mat &src, Mat &dst, double a, double b, int norm_type, int dtype = -1) {
    (void) norm_type;
    double lo = 1e300, hi = -1e300;
    const int n = src.cols * src.channels();
    for (int y = 0; y < src.rows; ++y)
        for (int i = 0; i < n; ++i) {
            double v = src.depth() == CV_32F ? src.ptr<float>(y)[i] : src.ptr<uint8_t>(y)[i];
            lo = std::min(lo, v); hi = std::max(hi, v);
        }
    const double dmin = std::min(a, b), dmax = std::max(a, b);
    const double scale = (dmax - dmin) * (hi - lo > 2.220446049250313e-16 ? 1.0 / (hi - lo) : 0.0);
    const double shift = dmin - lo * scale;
    const int rtype = dtype < 0 ? src.type() : CV_MAKETYPE(dtype, src.channels());
    src.convertTo(dst, rtype, scale, shift);
}

// PCA (DATA_AS_ROW, keep `maxComponents`): covariance + Jacobi eigen-solver.
class PCA {
public:
    enum Flags { DATA_AS_ROW = 0 };
    Mat mean, eigenvectors, eigenvalues;
    PCA() = default;
    PCA(const Mat &data, const Mat & /*mean*/, int /*flags*/, int maxComponents) {
        const int n = data.rows, d = data.cols;
        std::vector<double> mu(d, 0.0);
        for (int i = 0; i < n; ++i) for (int j = 0; j < d; ++j) mu[j] += data.at<float>(i, j);
        for (auto &v : mu) v /= n;
        std::vector<double> C((size_t) d * d, 0.0), row(d);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < d; ++j) row[j] = data.at<float>(i, j) - mu[j];
            for (int j = 0; j < d; ++j) { const double rj = row[j]; double *c = &C[(size_t) j * d];
                for (int k = j; k < d; ++k) c[k] += rj * row[k]; }
        }
        for (int j = 0; j < d; ++j) for (int k = j; k < d; ++k) { C[(size_t) j * d + k] /= n; C[(size_t) k * d + j] = C[(size_t) j * d + k]; }
        // top-k eigenvectors by deflated power iteration (k is 3 for the apps)
        mean = Mat(1, d, CV_32F);
        for (int j = 0; j < d; ++j) mean.at<float>(0, j) = (float) mu[j];
        eigenvectors = Mat(maxComponents, d, CV_32F);
        eigenvalues = Mat(maxComponents, 1, CV_32F);
        std::vector<double> v(d), w(d);
        for (int c = 0; c < maxComponents; ++c) {
            for (int j = 0; j < d; ++j) v[j] = 1.0 / std::sqrt((double) d) * ((j * 2654435761u >> 7) & 1 ? 1.0 : -1.0);
            double lambda = 0.0;
            for (int it = 0; it < 500; ++it) {
                for (int j = 0; j < d; ++j) { double s = 0.0; const double *cr = &C[(size_t) j * d]; for (int k = 0; k < d; ++k) s += cr[k] * v[k]; w[j] = s; }
                double nrm = 0.0; for (double x : w) nrm += x * x; nrm = std::sqrt(nrm);
                if (nrm < 1e-300) break;
                double diff = 0.0;
                for (int j = 0; j < d; ++j) { double nv = w[j] / nrm; diff += std::fabs(nv - v[j]); v[j] = nv; }
                lambda = nrm;
                if (diff < 1e-12) break;
            }
            for (int j = 0; j < d; ++j) eigenvectors.at<float>(c, j) = (float) v[j];
            eigenvalues.at<float>(c, 0) = (float) lambda;
            for (int j = 0; j < d; ++j) for (int k = 0; k < d; ++k) C[(size_t) j * d + k] -= lambda * v[j] * v[k];
        }
    }
    void project(const Mat &data, Mat &out) const {
        const int n = data.rows, d = data.cols, k = eigenvectors.rows;
        Mat r(n, k, CV_32F);
        for (int i = 0; i < n; ++i)
            for (int c = 0; c < k; ++c) {
                double s = 0.0;
                for (int j = 0; j < d; ++j) s += ((double) data.at<float>(i, j) - mean.at<float>(0, j)) * eigenvectors.at<float>(c, j);
                r.at<float>(i, c) = (float) s;
            }
        out = r;
    }
};

}  // namespace cv
