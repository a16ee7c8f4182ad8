// Shim for <cv_bridge/cv_bridge.h> — TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Lets the reference's own Labeler/label.{h,cpp} compile unmodified, from where they lie under
// /root/reference, without ROS/OpenCV.  Provides just the cv::Mat surface those files use
// (Mat::zeros, at<T>, rows, cols, copy) with LOGICAL indexing: every element owns an 8-byte slot and
// at<T>(r,c) views the first sizeof(T) bytes of the slot of element (r,c), whatever T is.  This is
// the "typed-accessor" contract of SURVEY.md appendix A (Q8): at<uint16_t> on the CV_8U label image
// addresses the label of pixel (r,c), not bytes r*W+2c of the buffer.
#pragma once
#include <sys/types.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_16UC1 2
#define CV_32F 5

namespace cv {

class Mat {
   public:
    int rows = 0, cols = 0;
    Mat() {}
    static Mat zeros(int r, int c, int type) {
        Mat m;
        m.rows = r;
        m.cols = c;
        m.type_ = type;
        m.slots_ = std::make_shared<std::vector<uint64_t>>(static_cast<size_t>(r) * c, 0);
        return m;
    }
    template <typename T>
    T& at(int r, int c) {
        static_assert(sizeof(T) <= sizeof(uint64_t), "slot too small");
        return *reinterpret_cast<T*>(&(*slots_)[static_cast<size_t>(r) * cols + c]);
    }
    template <typename T>
    const T& at(int r, int c) const {
        return *reinterpret_cast<const T*>(&(*slots_)[static_cast<size_t>(r) * cols + c]);
    }
    int type() const { return type_; }

   private:
    int type_ = 0;
    std::shared_ptr<std::vector<uint64_t>> slots_;  // shared like cv::Mat's refcounted header copies
};

}  // namespace cv
