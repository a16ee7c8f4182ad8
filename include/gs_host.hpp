// gs_host.hpp — the reference's C++ API surface for the cloud -> range -> ground -> cloud path, on light types, over
// the C ABI of gs_b200.h (libgs_b200.so, CUDA sm_100a).  Header-only, C++14, no ROS / PCL / OpenCV.
//
// What mirrors what (reference file:line, relative to matigolias/GroundSegmentationROSNode):
//   gs::Alfa_GS            class Alfa_GS                 lib/alfa_gs.h:33-83, src/alfa_gs.cpp
//   gs::Cloud2RangeNode    class Cloud2RangeNode         src/cloud2range_node.cpp:22-57, :381-480, :701-744
//   gs::Range2CloudNode    class Range2CloudNode         src/range2cloud_node.cpp:13-99
//   gs::cloud2range::Label class cloud2range::Label      Labeler/label.h:11-42, Labeler/label.cpp
//   gs::cloud2range::Point{Range,Azimuth,Altitude}, Deg_Rad, Rad_Deg      lib/utils.h:8-54
// Same member names, argument order and meaning; cv::Mat becomes gs::Image<T>, pcl::PointCloud<P> becomes
// std::vector<P>, sensor_msgs::CameraInfo becomes gs::CameraInfo (K[0..5] as the reference fills them,
// src/alfa_gs.cpp:58-63).  Every image / cloud producing member runs on the GPU through the C ABI; where the reference
// throws std::logic_error (bad window sizes, src/alfa_gs.cpp:631-633, :645-649) so do these.  ROS plumbing
// (subscribers, publishers, the config service, the FPGA registers) stays with the caller.
#ifndef GS_HOST_HPP
#define GS_HOST_HPP

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gs_b200.h"

namespace gs {

// ---- light stand-ins for cv::Mat / pcl points / sensor_msgs::CameraInfo --------------------------------------------
template <class T>
struct Image {
    int rows = 0, cols = 0;
    std::vector<T> data;
    Image() {}
    Image(int r, int c, T fill = T()) : rows(r), cols(c), data(static_cast<size_t>(r) * c, fill) {}
    static Image zeros(int r, int c) { return Image(r, c, T()); }
    T& at(int r, int c) { return data[static_cast<size_t>(r) * cols + c]; }  // logical element (r, c): SURVEY.md Q8
    const T& at(int r, int c) const { return data[static_cast<size_t>(r) * cols + c]; }
    bool empty() const { return data.empty(); }
    T* ptr() { return data.data(); }
    const T* ptr() const { return data.data(); }
};

struct PointXYZ {  // pcl::PointXYZ: 16 bytes
    float x, y, z, pad;
};
struct PointXYZI {  // pcl::PointXYZI: 16 bytes (cloud2range_node.cpp:18) — laid out as x, y, z, intensity for the device path
    float x, y, z, intensity;
};
struct PointXYZRGB {  // pcl::PointXYZRGB: 32 bytes (alfa_gs.h:25)
    float x, y, z, pad;
    uint8_t b, g, r, a;
    float pad2[3];
};
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 16 && sizeof(PointXYZRGB) == 32, "PCL point layouts");

struct CameraInfo {  // the fields the reference uses: K[0..5] = min_angle, max_angle, min_range, max_range, d_azimuth, d_altitude
    int height = 0, width = 0;
    std::array<double, 9> K{{0, 0, 0, 0, 0, 0, 0, 0, 0}};
};

// ---- lib/utils.h:8-54 ------------------------------------------------------------------------------------------------
namespace cloud2range {
inline double PointAltitude(double x, double y, double z) { return std::atan2(z, std::hypot(x, y)); }  // utils.h:8-11
inline double PointAzimuth(double x, double y) {                                                        // utils.h:24-28
    const double a = std::atan2(y, x);
    return (y >= 0 ? a : a + M_PI * 2);
}
inline double PointRange(double x, double y, double z) { return std::sqrt(x * x + y * y + z * z); }     // utils.h:41-43
inline double Deg_Rad(double rad) { return rad * 180 / M_PI; }                                           // utils.h:50
inline double Rad_Deg(double deg) { return deg / 180 * M_PI; }                                           // utils.h:54
}  // namespace cloud2range

// ---- shared GPU context -----------------------------------------------------------------------------------------------
class Context {
public:
    explicit Context(const gs_params& p, int device = 0) {
        gs_ctx* c = nullptr;
        const int rc = gs_create(&p, device, &c);
        if (rc == GS_EINVAL) throw std::logic_error("gs_b200: invalid parameters (window size must be odd, beams >= 2, ...)");
        if (rc != GS_OK) throw std::runtime_error("gs_b200: no usable CUDA device (there is no CPU fallback)");
        ctx_.reset(c, gs_destroy);
        gs_get_derived(c, &der_);
        gs_get_params(c, &par_);
    }
    gs_ctx* get() const { return ctx_.get(); }
    const gs_derived& derived() const { return der_; }
    const gs_params& params() const { return par_; }
    void set_params(const gs_params& p) {
        check(gs_set_params(ctx_.get(), &p));
        gs_get_derived(ctx_.get(), &der_);
        gs_get_params(ctx_.get(), &par_);
    }
    void check(int rc) const {
        if (rc == GS_OK) return;
        const std::string msg = gs_last_error(ctx_.get());
        if (rc == GS_EINVAL) throw std::logic_error(msg);  // the reference throws std::logic_error for bad windows
        throw std::runtime_error(msg);
    }

private:
    std::shared_ptr<gs_ctx> ctx_;
    gs_derived der_;
    gs_params par_;
};

namespace cloud2range {

// ---- Labeler/label.h:11-42 --------------------------------------------------------------------------------------------
class Label {
public:
    struct pixel_coord {
        pixel_coord() : row(0), col(0) {}
        pixel_coord(uint16_t row_, uint16_t col_) : row(row_), col(col_) {}
        pixel_coord operator+(const pixel_coord& o) const { return pixel_coord(row + o.row, col + o.col); }
        uint16_t row, col;
    };
    // Label(rows, cols) of the reference (label.cpp:8-21); `ctx` carries the image shape and the GPU
    Label(int n_rows, int n_cols, std::shared_ptr<Context> ctx) : label_image(Image<uint8_t>::zeros(n_rows, n_cols)), ctx_(std::move(ctx)) {
        if (n_rows != ctx_->derived().n_rows || n_cols != ctx_->derived().n_cols) throw std::logic_error("Label: shape differs from the context's sensor shape");
    }
    Image<uint8_t> GetLabelImage() { return label_image; }                       // label.cpp:28-31
    const Image<uint8_t>* GetLabelImagePtr() { return &label_image; }            // label.h:36
    uint8_t CheckLabelAt(pixel_coord c) { return label_image.at(c.row, c.col); } // label.h:37
    uint8_t CheckLabelAt(int row, int col) { return label_image.at(row, col); }  // label.h:38
    void SetLabelAt(int label, pixel_coord c) { label_image.at(c.row, c.col) = static_cast<uint8_t>(label); }  // label.h:39
    uint16_t WrapCols(int16_t col) {  // label.cpp:33-45
        if (col < 0) return static_cast<uint16_t>(col + label_image.cols);
        if (col >= label_image.cols) return static_cast<uint16_t>(col - label_image.cols);
        return static_cast<uint16_t>(col);
    }
    // label.cpp:48-104 — the BFS, run on the GPU as connected components over the unlabelled non-empty pixels
    void LabelOneComponent(uint8_t label, int row, int col, float ground_angle_threshold, const Image<uint16_t>& range_image,
                           const Image<uint16_t>& smoothed_image) {
        ctx_->check(gs_label_one_component(ctx_->get(), label, row, col, ground_angle_threshold, range_image.ptr(), smoothed_image.ptr(),
                                           label_image.ptr()));
    }

private:
    Image<uint8_t> label_image;
    std::shared_ptr<Context> ctx_;
};

}  // namespace cloud2range

// ---- lib/alfa_gs.h:33-83 ----------------------------------------------------------------------------------------------
class Alfa_GS {
public:
    // Alfa_GS::Alfa_GS (alfa_gs.cpp:12-93): the hard-coded constants are the GS_PRESET_ALFA_GS_DEFAULT parameter block
    explicit Alfa_GS(const gs_params* params = nullptr, int device = 0) {
        gs_params p;
        if (params) p = *params;
        else gs_default_params(GS_PRESET_ALFA_GS_DEFAULT, &p);
        ctx_ = std::make_shared<Context>(p, device);
        refresh_cinfo();
    }
    std::shared_ptr<Context> context() const { return ctx_; }
    const CameraInfo& cinfo() const { return cinfo_; }
    int n_rows() const { return ctx_->derived().n_rows; }
    int n_cols() const { return ctx_->derived().n_cols; }

    // alfa_gs.cpp:247-298
    template <class PointT>
    Image<uint16_t> CreateRangeImage(const std::vector<PointT>& input_cloud) {
        Image<uint16_t> out(n_rows(), n_cols());
        ctx_->check(gs_project(ctx_->get(), input_cloud.data(), input_cloud.size(), sizeof(PointT), out.ptr()));
        return out;
    }
    // the C ABI reads n_rows * n_cols pixels from every image it is handed: anything else is a caller error
    template <class T>
    void require_shape(const Image<T>& image, const char* what) const {
        if (image.rows != n_rows() || image.cols != n_cols())
            throw std::logic_error(std::string(what) + ": image shape differs from the sensor's range-image shape");
    }
    // alfa_gs.cpp:300-333 (typed reading)
    Image<uint16_t> RepairGaps(const Image<uint16_t>& no_ground_image, int step, float depth_threshold) {
        require_shape(no_ground_image, "RepairGaps");
        Image<uint16_t> out(n_rows(), n_cols());
        ctx_->check(gs_repair_gaps(ctx_->get(), no_ground_image.ptr(), step, depth_threshold, out.ptr()));
        return out;
    }
    // alfa_gs.cpp:335-386
    Image<uint16_t> CreateAngleImg(const Image<uint16_t>& range_image) {
        require_shape(range_image, "CreateAngleImg");
        Image<uint16_t> out(n_rows(), n_cols());
        ctx_->check(gs_angle_image(ctx_->get(), range_image.ptr(), out.ptr()));
        return out;
    }
    // alfa_gs.cpp:447-582: the reference computes it and discards the result (:212 vs :220), reading a 16U image through
    // at<float> with a hard-coded 3600-column loop.  The contract: the pipeline's output does not depend
    // on it; the call returns its input.
    Image<uint16_t> MovingAverageSmoothing(const Image<uint16_t>& angle_image, int /*window_size*/) { return angle_image; }
    // alfa_gs.cpp:630-639, executed on bytes as the reference does: two `float 1.0f` stores into a window_size-byte u8
    // column (the second overruns it), then `/= 2` with u8 rounding.
    Image<uint8_t> GetUniformKernel(int window_size, int /*type*/) {
        if (window_size % 2 == 0) throw std::logic_error("Window size must be odd");  // alfa_gs.cpp:631-633
        std::vector<uint8_t> buf(static_cast<size_t>(window_size) + 3, 0);
        const float one = 1.0f;
        std::memcpy(&buf[0], &one, 4);
        std::memcpy(&buf[window_size - 1], &one, 4);
        Image<uint8_t> k(window_size, 1);
        for (int i = 0; i < window_size; ++i) k.at(i, 0) = static_cast<uint8_t>(std::nearbyint(buf[i] / 2.0));
        return k;
    }
    // alfa_gs.cpp:641-703
    Image<float> GetSavitskyGolayKernel(int window_size) {
        if (window_size % 2 == 0) throw std::logic_error("Window size must be odd");  // :642-644
        static const float k5[] = {-3, 12, 17, 12, -3}, k7[] = {-2, 3, 6, 7, 6, 3, -2};
        static const float k9[] = {-21, 14, 39, 54, 59, 54, 39, 14, -21}, k11[] = {-36, 9, 44, 69, 84, 89, 84, 69, 44, 9, -36};
        const float* src;
        float norm;
        switch (window_size) {
            case 5: src = k5; norm = 35.f; break;
            case 7: src = k7; norm = 21.f; break;
            case 9: src = k9; norm = 231.f; break;
            case 11: src = k11; norm = 429.f; break;
            default: throw std::logic_error("bad window size");  // :645-649
        }
        Image<float> k(window_size, 1);
        const float alpha = static_cast<float>(1.0 / static_cast<double>(norm));  // `kernel /= norm` = convertTo(alpha)
        for (int i = 0; i < window_size; ++i) k.at(i, 0) = src[i] * alpha;
        return k;
    }
    // alfa_gs.cpp:705-713 (cv::filter2D with BORDER_REFLECT101 on the GPU)
    Image<uint16_t> SavitskyGolaySmoothing(const Image<uint16_t>& image, int window_size) {
        require_shape(image, "SavitskyGolaySmoothing");
        Image<uint16_t> out(n_rows(), n_cols());
        ctx_->check(gs_smooth(ctx_->get(), image.ptr(), window_size, out.ptr()));
        return out;
    }
    // alfa_gs.cpp:715-766 (Labeler BFS + GetUniformKernel + cv::dilate + erase on the GPU)
    Image<uint16_t> EraseBFS(const Image<uint16_t>& range_image, const Image<uint16_t>& smoothed_image, double ground_angle_threshold,
                             double start_angle_threshold, int kernel_size) {
        if (range_image.rows != smoothed_image.rows || range_image.cols != smoothed_image.cols) return Image<uint16_t>();  // :742-745: empty image
        require_shape(range_image, "EraseBFS");
        Image<uint16_t> out(n_rows(), n_cols());
        ctx_->check(gs_erase_ground(ctx_->get(), range_image.ptr(), smoothed_image.ptr(), ground_angle_threshold, start_angle_threshold,
                                    kernel_size, out.ptr(), nullptr));
        return out;
    }
    // alfa_gs.cpp:997-1074: decodes with the parameters of the CameraInfo it is handed (:1010-1015; the node passes the one it
    // filled at construction, :55-63), range rounded to float (:1027,1045)
    std::vector<PointXYZRGB> CameraCb(const Image<uint16_t>& range_image, const CameraInfo& cinfo) {
        require_shape(range_image, "CameraCb");
        std::vector<float> xyz(static_cast<size_t>(n_rows()) * n_cols() * 4);
        uint32_t n = 0;
        ctx_->check(gs_backproject_camera_info(ctx_->get(), range_image.ptr(), range_image.rows, range_image.cols, cinfo.K.data(), 0,
                                               xyz.data(), &n));
        std::vector<PointXYZRGB> cloud(n);
        for (uint32_t i = 0; i < n; ++i) {
            PointXYZRGB& p = cloud[i];
            std::memset(&p, 0, sizeof(p));
            p.x = xyz[4 * i + 0];
            p.y = xyz[4 * i + 1];
            p.z = xyz[4 * i + 2];
            p.pad = 1.0f;
            p.a = 255;  // PCL default: rgba = 0, 0, 0, 255
        }
        return cloud;
    }
    // alfa_gs.cpp:945-995: 1 everywhere, 0 where a kept point lands with the (sticky) ground-colour flag up
    Image<uint16_t> CreateLabeledRangeImage(const std::vector<PointXYZRGB>& input_cloud) {
        Image<uint16_t> out(n_rows(), n_cols());
        ctx_->check(gs_create_labeled_range_image(ctx_->get(), input_cloud.data(), input_cloud.size(), out.ptr()));
        return out;
    }
    // alfa_gs.cpp:831-943: the per-frame counters and IoU_g / Recall_g / Recall_mo; the running means the reference keeps in
    // function-local statics for its log line (:890-927) are accumulated here per object
    gs_eval CheckNumberOfDetectedRIdGnd(const Image<uint16_t>& og_range_image, const Image<uint16_t>& seg_range_image,
                                        const Image<uint16_t>& labeled_range_image) {
        require_shape(og_range_image, "CheckNumberOfDetectedRIdGnd");
        require_shape(seg_range_image, "CheckNumberOfDetectedRIdGnd");
        require_shape(labeled_range_image, "CheckNumberOfDetectedRIdGnd");
        gs_eval e;
        ctx_->check(gs_evaluate(ctx_->get(), og_range_image.ptr(), seg_range_image.ptr(), labeled_range_image.ptr(), &e));
        eval_sum_[0] += e.iou_g;
        eval_sum_[1] += e.recall_g;
        eval_sum_[2] += e.recall_mo;
        ++eval_cnt_;
        return e;
    }
    // mean IoU_g, Recall_g, Recall_mo over the frames evaluated so far (alfa_gs.cpp:921-927)
    void EvaluationMeans(float* iou_g, float* recall_g, float* recall_mo) const {
        const float n = static_cast<float>(eval_cnt_ ? eval_cnt_ : 1);
        *iou_g = eval_sum_[0] / n;
        *recall_g = eval_sum_[1] / n;
        *recall_mo = eval_sum_[2] / n;
    }
    // alfa_gs.cpp:175-244, the hw == 0 branch: range image -> angle image -> (discarded moving average) -> EraseBFS
    // -> CameraCb, fused on the GPU; the caller publishes the two results (alfa_node.cpp:27-57)
    struct Frame {
        Image<uint16_t> range_image, angle_image, no_ground_image;
        std::vector<PointXYZRGB> seg_cloud;
    };
    template <class PointT>
    Frame process_pointcloud(const std::vector<PointT>& input_cloud) {
        Frame f;
        f.range_image = Image<uint16_t>(n_rows(), n_cols());
        f.angle_image = Image<uint16_t>(n_rows(), n_cols());
        f.no_ground_image = Image<uint16_t>(n_rows(), n_cols());
        std::vector<float> xyz(static_cast<size_t>(n_rows()) * n_cols() * 4);
        uint32_t n = 0;
        ctx_->check(gs_process_frame(ctx_->get(), input_cloud.data(), input_cloud.size(), sizeof(PointT), f.range_image.ptr(),
                                     f.angle_image.ptr(), f.no_ground_image.ptr(), xyz.data(), &n));
        f.seg_cloud.resize(n);
        for (uint32_t i = 0; i < n; ++i) {
            PointXYZRGB& p = f.seg_cloud[i];
            std::memset(&p, 0, sizeof(p));
            p.x = xyz[4 * i + 0];
            p.y = xyz[4 * i + 1];
            p.z = xyz[4 * i + 2];
            p.pad = 1.0f;
            p.a = 255;
        }
        return f;
    }
    // alfa_gs.cpp:1122-1136: positional configuration vector (indices 0 erase_ground, 1 n_beams, 2 rpm, 3 sample_freq,
    // 4/5 min/max angle, 6/7 min/max range, 10 start threshold, 11 ground threshold, 13 window).  Unlike the reference
    // the derived values (n_cols, d_azimuth, d_altitude, cinfo) are refreshed and the vector is bounds-checked.
    void UpdateSegmentationSettings(const std::vector<double>& configurations) {
        if (configurations.size() < 14) throw std::logic_error("UpdateSegmentationSettings: 14 configuration values expected");
        gs_params p = ctx_->params();
        p.erase_ground = static_cast<int>(configurations[0]);
        p.n_beams = static_cast<int>(configurations[1]);
        p.rpm = static_cast<int>(configurations[2]);
        p.sample_freq = static_cast<int>(configurations[3]);
        p.min_angle = configurations[4];
        p.max_angle = configurations[5];
        p.min_range = configurations[6];
        p.max_range = configurations[7];
        p.start_angle_threshold = configurations[10];
        p.ground_angle_threshold = configurations[11];
        p.window_size = static_cast<int>(configurations[13]);
        ctx_->set_params(p);
        refresh_cinfo();
    }

private:
    float eval_sum_[3] = {0.f, 0.f, 0.f};
    int eval_cnt_ = 0;
    void refresh_cinfo() {  // alfa_gs.cpp:55-63
        const gs_params& p = ctx_->params();
        const gs_derived& d = ctx_->derived();
        cinfo_.height = d.n_rows;
        cinfo_.width = d.n_cols;
        cinfo_.K = {{p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth, d.d_altitude, 0, 0, 0}};
    }
    std::shared_ptr<Context> ctx_;
    CameraInfo cinfo_;
};

// ---- src/cloud2range_node.cpp ------------------------------------------------------------------------------------------
class Cloud2RangeNode {
public:
    // Cloud2RangeNode::Cloud2RangeNode (cloud2range_node.cpp:274-329): the OS1-64 parameter block = GS_PRESET_CLOUD2RANGE_DEFAULT
    explicit Cloud2RangeNode(const gs_params* params = nullptr, int device = 0) {
        gs_params p;
        if (params) p = *params;
        else gs_default_params(GS_PRESET_CLOUD2RANGE_DEFAULT, &p);
        p.variant = GS_VARIANT_CLOUD2RANGE;
        ctx_ = std::make_shared<Context>(p, device);
        const gs_derived& d = ctx_->derived();
        cinfo_.height = d.n_rows;
        cinfo_.width = d.n_cols;
        cinfo_.K = {{p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth, d.d_altitude, 0, 0, 0}};  // :319-328
    }
    std::shared_ptr<Context> context() const { return ctx_; }
    const CameraInfo& cinfo() const { return cinfo_; }
    // CloudCb (cloud2range_node.cpp:381-480): projection -> (RepairGaps) -> angle image -> Savitzky-Golay -> EraseGroundBFS;
    // returns the mono16 no-ground image the node publishes with its CameraInfo (:477-479)
    Image<uint16_t> CloudCb(const std::vector<PointXYZI>& cloud) {
        const gs_derived& d = ctx_->derived();
        Image<uint16_t> out(d.n_rows, d.n_cols);
        ctx_->check(gs_process_frame(ctx_->get(), cloud.data(), cloud.size(), sizeof(PointXYZI), nullptr, nullptr, out.ptr(), nullptr, nullptr));
        return out;
    }
    // EraseGroundBFS (cloud2range_node.cpp:701-744)
    Image<uint16_t> EraseGroundBFS(const Image<uint16_t>& range_image, const Image<uint16_t>& smoothed_image, double ground_angle_threshold,
                                   double start_angle_threshold, int kernel_size) {
        const gs_derived& d = ctx_->derived();
        if (range_image.rows != smoothed_image.rows || range_image.cols != smoothed_image.cols) return Image<uint16_t>();  // :728-731: empty image
        if (range_image.rows != d.n_rows || range_image.cols != d.n_cols)
            throw std::logic_error("EraseGroundBFS: image shape differs from the sensor's range-image shape");
        Image<uint16_t> out(d.n_rows, d.n_cols);
        ctx_->check(gs_erase_ground(ctx_->get(), range_image.ptr(), smoothed_image.ptr(), ground_angle_threshold, start_angle_threshold,
                                    kernel_size, out.ptr(), nullptr));
        return out;
    }

private:
    std::shared_ptr<Context> ctx_;
    CameraInfo cinfo_;
};

// ---- src/range2cloud_node.cpp ------------------------------------------------------------------------------------------
class Range2CloudNode {
public:
    // CameraCb (range2cloud_node.cpp:36-99): decodes with the parameters carried in CameraInfo.K (:51-58) — min_angle = K[0],
    // min / max range = K[2] / K[3], d_azimuth = K[4], d_altitude = K[5] are used as published, nothing is re-derived
    std::vector<PointXYZ> CameraCb(const Image<uint16_t>& range_image, const CameraInfo& cinfo) {
        if (range_image.rows != static_cast<int>(cinfo.height) || range_image.cols != static_cast<int>(cinfo.width))
            throw std::logic_error("Range2CloudNode::CameraCb: image shape differs from the CameraInfo's height x width");
        if (!ctx_ || range_image.rows != rows_ || range_image.cols != cols_) {
            // a context of the image's shape lends its device buffers; the decoding parameters travel with every call
            gs_params p;
            gs_default_params(GS_PRESET_CLOUD2RANGE_DEFAULT, &p);
            p.variant = GS_VARIANT_CLOUD2RANGE;
            p.n_beams = range_image.rows;
            p.rpm = 60;
            p.sample_freq = range_image.cols;  // n_cols = sample_freq * 60 / rpm
            p.smoothing = GS_SMOOTH_NONE;
            ctx_ = std::make_shared<Context>(p);
            rows_ = range_image.rows;
            cols_ = range_image.cols;
        }
        std::vector<float> xyz(static_cast<size_t>(range_image.rows) * range_image.cols * 4);
        uint32_t n = 0;
        ctx_->check(gs_backproject_camera_info(ctx_->get(), range_image.ptr(), range_image.rows, range_image.cols, cinfo.K.data(), 1,
                                               xyz.data(), &n));
        std::vector<PointXYZ> cloud(n);
        for (uint32_t i = 0; i < n; ++i) cloud[i] = PointXYZ{xyz[4 * i], xyz[4 * i + 1], xyz[4 * i + 2], 1.0f};
        return cloud;
    }

private:
    std::shared_ptr<Context> ctx_;
    int rows_ = 0, cols_ = 0;
};

// ---- multi-GPU host driver -----------------------------------------------------------------------------
// One host thread + one context + its streams per GPU, frames in contiguous blocks, no collective: N of the reference's
// one-frame-per-callback nodes (alfa_node.cpp:298-319) on N GPUs.  Results are identical to one GPU's, frame by frame.
class MultiGpu {
public:
    explicit MultiGpu(const gs_params& params, const std::vector<int>& devices = std::vector<int>()) {
        const int rc = gs_multi_create(&params, devices.empty() ? nullptr : devices.data(), static_cast<int>(devices.size()), &m_);
        if (rc != GS_OK) throw std::runtime_error(rc == GS_ECUDA ? "gs::MultiGpu: no usable CUDA device" : "gs::MultiGpu: invalid parameters or devices");
    }
    ~MultiGpu() { gs_multi_destroy(m_); }
    MultiGpu(const MultiGpu&) = delete;
    MultiGpu& operator=(const MultiGpu&) = delete;
    int n_gpus() const { return gs_multi_n_gpus(m_); }
    // page-locked memory first touched shard by shard (NUMA-local to the GPU that copies it)
    void* host_alloc(size_t bytes) {
        void* p = nullptr;
        if (gs_multi_host_alloc(m_, bytes, &p) != GS_OK) throw std::bad_alloc();
        return p;
    }
    void host_free(void* p, size_t bytes) { gs_multi_host_free(m_, p, bytes); }
    // shard g of G processes frames [g * F / G, (g + 1) * F / G); frame f's cloud is records [first[f], first[f] + count[f]) of xyz4_out
    void process_batch_host(const void* points, size_t stride_bytes, const int64_t* offsets, int n_frames, uint16_t* no_ground_out, float* xyz4_out,
                            uint64_t xyz4_capacity, int64_t* first, uint32_t* count) {
        if (gs_multi_process_batch_host(m_, points, stride_bytes, offsets, n_frames, no_ground_out, xyz4_out, xyz4_capacity, first, count) != GS_OK)
            throw std::runtime_error(std::string("gs::MultiGpu: ") + gs_multi_last_error(m_));
    }

private:
    gs_multi* m_ = nullptr;
};

}  // namespace gs
#endif  // GS_HOST_HPP
