// gsref_cv.h — stand-in for the OpenCV surface the reference touches.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Lets the reference's own translation units (src/alfa_gs.cpp, src/alfa_node.cpp, Labeler/label.cpp,
// src/cloud2range_node.cpp, src/range2cloud_node.cpp) compile UNMODIFIED from where they lie under /root/reference,
// without OpenCV.  cv::Mat here is a continuous row-major matrix with the access rules of SURVEY.md appendix A
// ("typed accessor" contract), quirk by quirk:
//   * storage is continuous, at<T>(r, c) addresses flat element r * cols + c: column index == cols aliases the next
//     row exactly like a continuous cv::Mat (Q2); an access past the end of the buffer lands in a scratch slot
//     (the literal code writes out of bounds there; contract: dropped).
//   * CV_8U matrices keep 2-byte slots so that at<uint16_t>(r, c) on the label / dilated images addresses the label
//     of pixel (r, c) and not bytes r * W + 2c (Q8: logical indexing) ...
//   * ... except 8-bit COLUMN vectors (cols == 1), which are byte-true with slack behind the buffer: GetUniformKernel
//     writes two float 1.0 into a ks x 1 CV_8U kernel and the resulting BYTES are the structuring element (Q13, literal).
//   * at<float> on an integer image (RepairGaps, Q22) is the typed reading: the element's value as a float; what the
//     caller stores through the reference is written back truncated to the element type when that element is next
//     looked at (or at the next access of another kind).
// cv::dilate / cv::filter2D / Mat::operator/= restate what OpenCV 4.13 computes for the calls the reference makes;
// they are pinned against the committed OpenCV vectors in tests/golden (tests/test_oracle_vs_ref.py).
#pragma once
#include <sys/types.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>

#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_8UC1 0
#define CV_16UC1 2
#define CV_32FC1 5
#define CV_8UC3 16

namespace cv {

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Point {
    int x = 0, y = 0;
    Point() {}
    Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
};
struct Vec3b {
    unsigned char val[3];
    unsigned char& operator[](int i) { return val[i]; }
    const unsigned char& operator[](int i) const { return val[i]; }
};
template <typename T>
struct DataType;
template <>
struct DataType<float> { enum { type = CV_32F }; };
template <>
struct DataType<unsigned char> { enum { type = CV_8U }; };
template <>
struct DataType<unsigned short> { enum { type = CV_16U }; };

enum { BORDER_CONSTANT = 0, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4 };

class Mat {
    struct Store {
        std::vector<unsigned char> bytes;  // rows * cols slots + slack
        struct Shadow { float v; unsigned pending; };
        std::vector<Shadow> shadow;        // typed at<float> views of an integer image (Q22), allocated on first use;
                                           // pending: handed out through at<float> and not yet written back
        std::vector<size_t> plist;         // the pending elements
        alignas(16) unsigned char sink[16];
    };
    std::shared_ptr<Store> st_;  // shared like cv::Mat's refcounted header copies
    size_t off_ = 0;             // first element (row views)
    int type_ = 0, es_ = 0;      // element type, slot bytes

    static int slot_bytes(int type, int cols) {
        switch (type) {
            case CV_8U: return cols == 1 ? 1 : 2;  // Q13 literal column vectors / Q8 logical label images
            case CV_16U: return 2;
            case CV_32F: return 4;
            case CV_8UC3: return 3;
        }
        return 8;
    }
    bool int_image() const { return (type_ == CV_16U) || (type_ == CV_8U && cols != 1); }
    size_t count() const { return static_cast<size_t>(rows) * cols; }
    double elem(size_t i) const {
        const unsigned char* p = &st_->bytes[(off_ + i) * es_];
        switch (type_) {
            case CV_8U: return *p;
            case CV_16U: { uint16_t v; std::memcpy(&v, p, 2); return v; }
            case CV_32F: { float v; std::memcpy(&v, p, 4); return v; }
        }
        return *p;
    }
    void set_elem(size_t i, double v) {
        unsigned char* p = &st_->bytes[(off_ + i) * es_];
        switch (type_) {
            case CV_8U: *p = static_cast<unsigned char>(v); if (es_ == 2) p[1] = 0; break;
            case CV_16U: { const uint16_t u = static_cast<uint16_t>(v); std::memcpy(p, &u, 2); break; }
            case CV_32F: { const float f = static_cast<float>(v); std::memcpy(p, &f, 4); break; }
            default: *p = static_cast<unsigned char>(v);
        }
    }
    void write_back(size_t abs_idx) const {  // typed store through an at<float> reference: truncated to the element type
        Store& s = *st_;
        unsigned char* p = &s.bytes[abs_idx * es_];
        const float f = s.shadow[abs_idx].v;
        if (type_ == CV_16U) { const uint16_t u = static_cast<uint16_t>(f); std::memcpy(p, &u, 2); }
        else { p[0] = static_cast<unsigned char>(f); }
    }

   public:
    int rows = 0, cols = 0;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); fill(s); }
    Mat(Size sz, int type, const Scalar& s) { create(sz.height, sz.width, type); fill(s); }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size sz, int type) { return Mat(sz.height, sz.width, type); }

    void create(int r, int c, int type) {
        rows = r;
        cols = c;
        type_ = type;
        es_ = slot_bytes(type, c);
        off_ = 0;
        st_ = std::make_shared<Store>();
        st_->bytes.assign(count() * es_ + 16, 0);
    }
    void fill(const Scalar& s) {
        for (size_t i = 0; i < count(); ++i) {
            if (type_ == CV_8UC3) for (int k = 0; k < 3; ++k) st_->bytes[(off_ + i) * 3 + k] = static_cast<unsigned char>(s.val[k]);
            else set_elem(i, s.val[0]);
        }
    }
    // written-back state of every outstanding at<float> reference (call before any other access)
    // any other kind of access ends the life of the outstanding at<float> references
    void settle() const {
        if (!st_ || st_->plist.empty()) return;
        for (size_t idx : st_->plist) {
            write_back(idx);
            st_->shadow[idx].pending = 0;
        }
        st_->plist.clear();
    }
    void sync() const { settle(); }
    bool empty() const { return !st_ || count() == 0; }
    int type() const { return type_; }
    Size size() const { return Size(cols, rows); }
    size_t slot() const { return static_cast<size_t>(es_); }

    template <typename T>
    T& at(int r, int c) const {
        Store& s = *st_;
        const size_t idx = off_ + static_cast<size_t>(static_cast<long long>(r) * cols + c);  // flat: continuous matrix (Q2)
        const size_t total = (s.bytes.size() - 16) / es_;
        if (std::is_floating_point<T>::value && int_image()) {
            // typed reading of an integer image (Q22)
            static_assert(!std::is_floating_point<T>::value || sizeof(T) <= 8, "");
            if (s.shadow.empty()) s.shadow.assign(total + 4, Store::Shadow{0.0f, 0u});
            if (idx >= total) {  // out of bounds in the literal code: reads as empty, stores are dropped
                std::memset(s.sink, 0, sizeof(s.sink));
                return *reinterpret_cast<T*>(s.sink);
            }
            // a store made through an earlier reference to this element becomes visible (truncated to the element type) now;
            // stores to other elements are written back when they are next looked at, or at the next access of another kind
            Store::Shadow& sh = s.shadow[idx];
            if (sh.pending) {
                const float t = type_ == CV_16U ? static_cast<float>(static_cast<uint16_t>(sh.v)) : static_cast<float>(static_cast<unsigned char>(sh.v));
                if (t != sh.v) {
                    write_back(idx);
                    sh.v = t;
                }
            } else {
                sh.pending = 1;
                s.plist.push_back(idx);
                sh.v = static_cast<float>(elem(idx - off_));
            }
            return *reinterpret_cast<T*>(&sh.v);
        }
        settle();
        const size_t need = idx * es_ + sizeof(T);
        if (need > s.bytes.size() - (es_ == 1 ? 0 : 16)) {  // byte-true column vectors may overrun into the slack (Q13)
            std::memset(s.sink, 0, sizeof(s.sink));
            return *reinterpret_cast<T*>(s.sink);
        }
        return *reinterpret_cast<T*>(&s.bytes[idx * es_]);
    }
    template <typename T>
    T* ptr(int r = 0) const {
        settle();
        return reinterpret_cast<T*>(&st_->bytes[(off_ + static_cast<size_t>(r) * cols) * es_]);
    }
    unsigned char* raw() const { settle(); return &st_->bytes[off_ * es_]; }

    Mat clone() const {
        Mat m;
        if (!st_) return m;
        settle();
        m.create(rows, cols, type_);
        std::memcpy(m.st_->bytes.data(), raw(), count() * es_);
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    void copyTo(Mat&& dst) const {  // into a view (row(r)): same shape assumed
        settle();
        std::memcpy(dst.raw(), raw(), std::min(count() * es_, dst.count() * dst.es_));
    }
    Mat row(int r) const {
        Mat v = *this;
        v.rows = 1;
        v.off_ = off_ + static_cast<size_t>(r) * cols;
        return v;
    }
    // Mat /= s is Mat::convertTo(m, -1, 1. / s): every element becomes saturate_cast<T>(v * (float)(1. / s)), cvRound
    // (round half to even) for the integer types
    Mat& operator/=(double s) {
        settle();
        const float alpha = static_cast<float>(1.0 / s);
        for (size_t i = 0; i < count(); ++i) {
            const float v = static_cast<float>(elem(i)) * alpha;
            if (type_ == CV_32F) set_elem(i, v);
            else {
                const long q = std::lrint(v);
                const long hi = type_ == CV_8U ? 255 : 65535;
                set_elem(i, static_cast<double>(std::min(std::max(q, 0l), hi)));
            }
        }
        return *this;
    }
    double value(int r, int c) const { settle(); return elem(static_cast<size_t>(r) * cols + c); }
    void set_value(int r, int c, double v) { settle(); set_elem(static_cast<size_t>(r) * cols + c, v); }
};

// only what CreateAngleImg2 (never executed on the path) needs to compile
inline Mat operator*(const Mat& m, double s) {
    Mat o = m.clone();
    for (int r = 0; r < o.rows; ++r)
        for (int c = 0; c < o.cols; ++c) o.set_value(r, c, m.value(r, c) * s);
    return o;
}

inline int countNonZero(const Mat& m) {
    int n = 0;
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c) n += m.value(r, c) != 0;
    return n;
}

inline int borderReflect101(int i, int n) {  // gfedcb|abcdefgh|gfedcba
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// cv::dilate(src, dst, kernel) with the defaults the reference uses (anchor = centre, one iteration, constant border
// that never wins the maximum): structuring element = the kernel's non-zero pattern; an all-zero kernel degenerates to the
// single tap at kernel coordinate (0, 0) (OpenCV preprocess2DKernel).
inline void dilate(const Mat& src, Mat& dst, const Mat& kernel) {
    std::vector<std::pair<int, int>> taps;
    for (int ky = 0; ky < kernel.rows; ++ky)
        for (int kx = 0; kx < kernel.cols; ++kx)
            if (kernel.value(ky, kx) != 0) taps.emplace_back(ky - kernel.rows / 2, kx - kernel.cols / 2);
    if (taps.empty()) taps.emplace_back(0 - kernel.rows / 2, 0 - kernel.cols / 2);
    Mat out = Mat::zeros(src.size(), src.type());
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            double m = 0;
            for (const auto& t : taps) {
                const int rr = r + t.first, cc = c + t.second;
                if (rr >= 0 && rr < src.rows && cc >= 0 && cc < src.cols) m = std::max(m, src.value(rr, cc));
            }
            out.set_value(r, c, m);
        }
    dst = out;
}

// cv::filter2D(src, dst, -1, kernel, anchor, delta, border): correlation with float32 accumulation in tap order, result
// rounded half to even and saturated to the source type (pinned against OpenCV 4.13 for the ws x 1 Savitzky-Golay kernels)
inline void filter2D(const Mat& src, Mat& dst, int /*ddepth*/, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0,
                     int /*borderType*/ = BORDER_REFLECT101) {
    const int ay = anchor.y < 0 ? kernel.rows / 2 : anchor.y, ax = anchor.x < 0 ? kernel.cols / 2 : anchor.x;
    Mat out = Mat::zeros(src.size(), src.type());
    const long hi = src.type() == CV_8U ? 255 : 65535;
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) {
            float acc = 0.0f;
            for (int ky = 0; ky < kernel.rows; ++ky)
                for (int kx = 0; kx < kernel.cols; ++kx) {
                    const int rr = borderReflect101(r + ky - ay, src.rows), cc = borderReflect101(c + kx - ax, src.cols);
                    const float prod = static_cast<float>(kernel.value(ky, kx)) * static_cast<float>(src.value(rr, cc));
                    acc = acc + prod;
                }
            acc = acc + static_cast<float>(delta);
            if (src.type() == CV_32F) out.set_value(r, c, acc);
            else out.set_value(r, c, static_cast<double>(std::min(std::max(std::lrint(acc), 0l), hi)));
        }
    dst = out;
}

}  // namespace cv
