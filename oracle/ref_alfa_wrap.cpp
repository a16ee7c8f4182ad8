// ref_alfa_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN node, compiled verbatim from /root/reference into oracle/_ref/libgs_ref.so:
//   src/alfa_gs.cpp    (class Alfa_GS: CreateRangeImage, CreateAngleImg, SavitskyGolaySmoothing, GetUniformKernel, EraseBFS,
//                       CameraCb, RepairGaps, CreateLabeledRangeImage, CheckNumberOfDetectedRIdGnd, CreateColoredAngleImage,
//                       process_pointcloud)
//   src/alfa_node.cpp  (class AlfaNode: cloud_cb, the publishers, the FPGA-format codecs store_pointcloud_hardware /
//                       read_hardware_pointcloud / read_hardware_filtered_angle_image)
//   Labeler/label.cpp
// against the stand-in headers of oracle/shim (no ROS, PCL or OpenCV in this image).  Nothing of the reference is copied
// into this repository; this file only builds inputs, calls the reference's members and copies their results out.
// The whole frame runs exactly as under ROS: a sensor_msgs::PointCloud2 goes to the callback AlfaNode subscribed
// (AlfaNode::cloud_cb -> virtual Alfa_GS::process_pointcloud), and what the node publishes is read back.
//
// The Alfa_GS constructor hard-codes its parameters (src/alfa_gs.cpp:16-38) and derives n_cols_ / d_azimuth_ / d_altitude_ /
// cinfo_ only there (:48-63).  For other sensor shapes the private fields are overwritten after construction with the same
// three formulas (`derive` below); gsref_alfa_create checks that `derive` reproduces what the constructor itself computed.
#include "shim/gsref_ros.h"

#define private public
#include "alfa_gs.h"  // -I/root/reference/lib
#undef private

#include "../include/gs_b200.h" /* POD parameter struct only */

namespace {

struct Node {
    std::vector<alfa_msg::ConfigMessage> defaults;
    Alfa_GS* gs = nullptr;
};

void derive(Alfa_GS& g) {  // src/alfa_gs.cpp:48-63, the statements of the constructor
    g.n_cols_ = g.sample_freq_ * 60 / g.rpm_;
    g.d_azimuth_ = cloud2range::Rad_Deg(360.0 / g.n_cols_);
    g.d_altitude_ = (g.max_angle_ - g.min_angle_) / (g.n_beams_ - 1);
    g.cinfo_.height = g.n_beams_;
    g.cinfo_.width = g.n_cols_;
    g.cinfo_.K[0] = g.min_angle_;
    g.cinfo_.K[1] = g.max_angle_;
    g.cinfo_.K[2] = g.min_range_;
    g.cinfo_.K[3] = g.max_range_;
    g.cinfo_.K[4] = g.d_azimuth_;
    g.cinfo_.K[5] = g.d_altitude_;
}

cv::Mat mat_u16(const uint16_t* src, int rows, int cols) {
    cv::Mat m = cv::Mat::zeros(rows, cols, CV_16UC1);
    std::memcpy(m.raw(), src, static_cast<size_t>(rows) * cols * 2);
    return m;
}
void out_u16(const cv::Mat& m, uint16_t* dst) { std::memcpy(dst, m.raw(), static_cast<size_t>(m.rows) * m.cols * 2); }

pcl::PointCloud<pcl::PointXYZRGB>::Ptr cloud_from(const void* pts, size_t n, size_t stride, const uint8_t* rgb) {
    pcl::PointCloud<pcl::PointXYZRGB>::Ptr c(new pcl::PointCloud<pcl::PointXYZRGB>);
    c->points.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const unsigned char* rec = static_cast<const unsigned char*>(pts) + i * stride;
        pcl::PointXYZRGB& p = c->points[i];
        std::memcpy(&p.x, rec, 4);
        std::memcpy(&p.y, rec + 4, 4);
        std::memcpy(&p.z, rec + 8, 4);
        if (rgb) { p.r = rgb[3 * i]; p.g = rgb[3 * i + 1]; p.b = rgb[3 * i + 2]; }
    }
    return c;
}

}  // namespace

extern "C" {

// what the UNTOUCHED constructor computed (64 beams, 600 rpm, 18 kHz): for pinning gs_derive on the reference's own default
void gsref_alfa_ctor_values(int* n_cols, double* d_azimuth, double* d_altitude, double* K6, int* hw) {
    std::vector<alfa_msg::ConfigMessage> d;
    Alfa_GS g("alfa_gs", "gs", &d);
    *n_cols = g.n_cols_;
    *d_azimuth = g.d_azimuth_;
    *d_altitude = g.d_altitude_;
    for (int i = 0; i < 6; ++i) K6[i] = g.cinfo_.K[i];
    *hw = g.hw ? 1 : 0;
}

void* gsref_alfa_create(const gs_params* p) {
    Node* n = new Node;
    n->gs = new Alfa_GS("alfa_gs", "gs", &n->defaults);  // subscribes cloud_cb, advertises the publishers; hw stays 0 (no /dev/mem)
    Alfa_GS& g = *n->gs;
    if (g.hw) { delete n->gs; delete n; return nullptr; }
    {   // `derive` must reproduce the constructor on the constructor's own parameters
        const int nc = g.n_cols_;
        const double da = g.d_azimuth_, dl = g.d_altitude_;
        derive(g);
        if (nc != g.n_cols_ || da != g.d_azimuth_ || dl != g.d_altitude_) { delete n->gs; delete n; return nullptr; }
    }
    if (p) {
        g.n_beams_ = p->n_beams;
        g.rpm_ = p->rpm;
        g.sample_freq_ = p->sample_freq;
        g.min_angle_ = p->min_angle;
        g.max_angle_ = p->max_angle;
        g.min_range_ = p->min_range;
        g.max_range_ = p->max_range;
        g.ground_angle_threshold = p->ground_angle_threshold;
        g.start_angle_threshold = p->start_angle_threshold;
        g.window_size = p->window_size;
        g.erase_ground = p->erase_ground;
        derive(g);
    }
    return n;
}
void gsref_alfa_destroy(void* h) {
    Node* n = static_cast<Node*>(h);
    if (!n) return;
    delete n->gs;
    delete n;
}
void gsref_alfa_shape(void* h, int* rows, int* cols) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    *rows = g.n_beams_;
    *cols = g.n_cols_;
}
long gsref_assert_failures(void) { return gsref::state().assert_failures; }
void gsref_reset_asserts(void) { gsref::state().assert_failures = 0; }
// the last ROS_INFO lines, newest last, '\n'-separated
int gsref_log(char* out, int cap) {
    std::string all;
    for (const std::string& s : gsref::state().log) all += s + "\n";
    const int n = static_cast<int>(std::min<size_t>(all.size(), cap > 0 ? cap - 1 : 0));
    if (cap > 0) { std::memcpy(out, all.data(), n); out[n] = 0; }
    return static_cast<int>(all.size());
}

// ---- the stages, one reference member each ------------------------------------------------------------------------
void gsref_alfa_create_range_image(void* h, const void* pts, size_t n, size_t stride, uint16_t* range_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    out_u16(g.CreateRangeImage(cloud_from(pts, n, stride, nullptr)), range_out);
}
void gsref_alfa_create_angle_img(void* h, const uint16_t* range, uint16_t* angle_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    out_u16(g.CreateAngleImg(mat_u16(range, g.n_beams_, g.n_cols_)), angle_out);
}
void gsref_alfa_repair_gaps(void* h, const uint16_t* range, int step, float depth_threshold, uint16_t* out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    out_u16(g.RepairGaps(mat_u16(range, g.n_beams_, g.n_cols_), step, depth_threshold), out);
}
// returns 0, or 1 when the reference threw std::logic_error (even / unknown window)
int gsref_alfa_sg_smooth(void* h, const uint16_t* img, int rows, int cols, int window, uint16_t* out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    try {
        out_u16(g.SavitskyGolaySmoothing(mat_u16(img, rows, cols), window), out);
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}
int gsref_alfa_sg_kernel(void* h, int window, float* taps_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    try {
        const cv::Mat k = g.GetSavitskyGolayKernel(window);
        for (int i = 0; i < window; ++i) taps_out[i] = k.at<float>(i, 0);
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}
int gsref_alfa_uniform_kernel(void* h, int ks, uint8_t* bytes_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    try {
        const cv::Mat k = g.GetUniformKernel(ks, CV_8U);
        for (int i = 0; i < ks; ++i) bytes_out[i] = k.at<uint8_t>(i, 0);
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}
// cv::dilate of the shim with the reference's own kernel, for the OpenCV golden vectors
int gsref_alfa_dilate(void* h, const uint8_t* label, int rows, int cols, int ks, uint8_t* out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    try {
        cv::Mat src = cv::Mat::zeros(rows, cols, CV_8U), dst;
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) src.set_value(r, c, label[static_cast<size_t>(r) * cols + c]);
        cv::dilate(src, dst, g.GetUniformKernel(ks, CV_8U));
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) out[static_cast<size_t>(r) * cols + c] = static_cast<uint8_t>(dst.value(r, c));
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}
int gsref_alfa_erase_bfs(void* h, const uint16_t* range, const uint16_t* angle, double ground_thr, double start_thr, int kernel_size,
                         uint16_t* res_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    try {
        out_u16(g.EraseBFS(mat_u16(range, g.n_beams_, g.n_cols_), mat_u16(angle, g.n_beams_, g.n_cols_), ground_thr, start_thr, kernel_size),
                res_out);
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}
// Alfa_GS::CameraCb with the node's own cinfo_; xyz_rgb_out: 32-byte pcl::PointXYZRGB records (may be NULL: count only)
size_t gsref_alfa_camera_cb(void* h, const uint16_t* image, void* points_out, size_t cap) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    const pcl::PointCloud<pcl::PointXYZRGB>::Ptr c = g.CameraCb(mat_u16(image, g.n_beams_, g.n_cols_), g.cinfo_);
    if (points_out) std::memcpy(points_out, c->points.data(), std::min(cap, c->points.size()) * sizeof(pcl::PointXYZRGB));
    return c->points.size();
}
void gsref_alfa_labeled_range_image(void* h, const void* pts, size_t n, size_t stride, const uint8_t* rgb, uint16_t* out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    out_u16(g.CreateLabeledRangeImage(cloud_from(pts, n, stride, rgb)), out);
}
// CheckNumberOfDetectedRIdGnd reports only through ROS_INFO: read the lines back with gsref_log
void gsref_alfa_check_detected(void* h, const uint16_t* og, const uint16_t* seg, const uint16_t* labeled) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    g.CheckNumberOfDetectedRIdGnd(mat_u16(og, g.n_beams_, g.n_cols_), mat_u16(seg, g.n_beams_, g.n_cols_), mat_u16(labeled, g.n_beams_, g.n_cols_));
}
void gsref_alfa_colored_angle_image(void* h, const uint16_t* angle, int rows, int cols, uint8_t* rgb_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    const cv::Mat m = g.CreateColoredAngleImage(mat_u16(angle, rows, cols));
    std::memcpy(rgb_out, m.raw(), static_cast<size_t>(rows) * cols * 3);
}

// ---- the whole frame, as under ROS: PointCloud2 -> AlfaNode::cloud_cb -> Alfa_GS::process_pointcloud -> publishers ------
// pts: n records of `stride` bytes, x / y / z floats first.  no_ground_out: H x W u16; points_out: 32-byte PointXYZRGB records.
// Returns the number of published points, or -1 when nothing was published.
long gsref_alfa_process_frame(void* h, const void* pts, size_t n, size_t stride, uint16_t* no_ground_out, void* points_out, size_t cap) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    msg->width = static_cast<uint32_t>(n);
    msg->height = 1;
    msg->point_step = static_cast<uint32_t>(stride);
    msg->row_step = static_cast<uint32_t>(stride * n);
    msg->layout = stride >= 32 ? 1 : 0;
    msg->data.assign(static_cast<const uint8_t*>(pts), static_cast<const uint8_t*>(pts) + n * stride);
    gsref::CameraFrame& cam = gsref::CameraFrame::get();
    gsref::Topic<sensor_msgs::PointCloud2>& top = gsref::Topic<sensor_msgs::PointCloud2>::get();
    const long before_img = cam.n_published, before_cloud = top.n_published;
    top.cb(msg);  // AlfaNode::cloud_cb (src/alfa_node.cpp:298-319)
    if (cam.n_published == before_img || top.n_published == before_cloud) return -1;
    if (no_ground_out) std::memcpy(no_ground_out, cam.image.data.data(), std::min(cam.image.data.size(), static_cast<size_t>(g.n_beams_) * g.n_cols_ * 2));
    const size_t np = static_cast<size_t>(top.last.width) * top.last.height;
    if (points_out) std::memcpy(points_out, top.last.data.data(), std::min(cap, np) * top.last.point_step);
    return static_cast<long>(np);
}

// ---- FPGA wire formats of AlfaNode (src/alfa_node.cpp:105-271) ------------------------------------------------------
void gsref_node_store_pointcloud_hardware(void* h, const void* pts, size_t n, size_t stride, uint64_t* ddr_out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    g.store_pointcloud_hardware(cloud_from(pts, n, stride, nullptr), reinterpret_cast<u64*>(ddr_out));
}
// the reference hard-codes a 64 x 1800 image here (alfa_node.cpp:162)
void gsref_node_read_hardware_pointcloud(void* h, const uint64_t* ddr, unsigned rows, unsigned cols, uint16_t* out64x1800) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    const cv::Mat m = g.read_hardware_pointcloud(reinterpret_cast<u64*>(const_cast<uint64_t*>(ddr)), rows, cols);
    out_u16(m, out64x1800);
}
// consumes (shifts) the words it reads, like the reference
void gsref_node_read_hardware_filtered_angle_image(void* h, uint64_t* ddr, unsigned rows, unsigned cols, uint16_t* out) {
    Alfa_GS& g = *static_cast<Node*>(h)->gs;
    out_u16(g.read_hardware_filtered_angle_image(reinterpret_cast<u64*>(ddr), rows, cols), out);
}

}  // extern "C"
