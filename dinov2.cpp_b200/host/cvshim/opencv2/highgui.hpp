// Camera / window stand-ins so that the reference's realtime.cpp compiles and links where no display or capture
// device (and no OpenCV SDK) exists: VideoCapture never opens, imshow/waitKey are no-ops.  See core.hpp.
#pragma once
#include "core.hpp"

namespace cv {

class VideoCapture {
public:
    VideoCapture() = default;
    explicit VideoCapture(int /*index*/) {}
    bool isOpened() const { return false; }
    bool read(Mat &frame) { frame = Mat(); return false; }
    void release() {}
};

inline void imshow(const std::string & /*name*/, const Mat & /*img*/) {}
inline int waitKey(int /*delay*/ = 0) { return 'q'; }
inline void destroyAllWindows() {}

}  // namespace cv
