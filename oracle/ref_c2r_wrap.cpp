// ref_c2r_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// The reference's stand-alone cloud -> range-image node, src/cloud2range_node.cpp, compiled verbatim (the file is pulled in
// as it lies under /root/reference; its `main` is renamed away) against the stand-in headers of oracle/shim, into
// oracle/_ref/libgs_ref.so.  Parameters reach the node the way ROS delivers them: NodeHandle::param answers from a table
// filled here, so the constructor's own derivations (n_cols_, d_azimuth_, d_altitude_, cinfo_; :305-328) run untouched.
// A frame goes to the callback the node subscribed (Cloud2RangeNode::CloudCb, :381-480) as a sensor_msgs::PointCloud2 and
// the published mono16 image + CameraInfo are read back.
#include "shim/gsref_ros.h"

#define main gsref_cloud2range_main
#define private public
#include "src/cloud2range_node.cpp"  // -I/root/reference
#undef private
#undef main

#include "../include/gs_b200.h" /* POD parameter struct only */

using cloud2range::Cloud2RangeNode;

namespace {
cv::Mat mat_u16(const uint16_t* src, int rows, int cols) {
    cv::Mat m = cv::Mat::zeros(rows, cols, CV_16UC1);
    std::memcpy(m.raw(), src, static_cast<size_t>(rows) * cols * 2);
    return m;
}
void out_u16(const cv::Mat& m, uint16_t* dst) { std::memcpy(dst, m.raw(), static_cast<size_t>(m.rows) * m.cols * 2); }
}  // namespace

extern "C" {

void* gsref_c2r_create(const gs_params* p) {
    std::map<std::string, double>& t = gsref::state().params;
    t.clear();
    if (p) {
        t["window_size"] = p->window_size;
        t["ground_angle_threshold"] = p->ground_angle_threshold;
        t["start_angle_threshold"] = p->start_angle_threshold;
        t["n_beams"] = p->n_beams;
        t["rpm"] = p->rpm;
        t["sample_freq"] = p->sample_freq;
        t["min_angle"] = p->min_angle;
        t["max_angle"] = p->max_angle;
        t["min_range"] = p->min_range;
        t["max_range"] = p->max_range;
    }
    return new Cloud2RangeNode(ros::NodeHandle("~"));
}
void gsref_c2r_destroy(void* h) { delete static_cast<Cloud2RangeNode*>(h); }
// what the constructor derived (cloud2range_node.cpp:305-328)
void gsref_c2r_derived(void* h, int* rows, int* cols, double* d_azimuth, double* d_altitude, double* K6) {
    Cloud2RangeNode& n = *static_cast<Cloud2RangeNode*>(h);
    *rows = n.n_beams_;
    *cols = n.n_cols_;
    *d_azimuth = n.d_azimuth_;
    *d_altitude = n.d_altitude_;
    for (int i = 0; i < 6; ++i) K6[i] = n.cinfo_.K[i];
}
// the whole callback: 16-byte XYZI (or `stride`-byte) records in, published no-ground image + CameraInfo.K out.
// start_thr_override / ground_thr_override < -1e30 keep the node's own thresholds; a start threshold of -1 rejects every
// seed, so the published image is the plain range image of the projection loop (:396-455).
int gsref_c2r_cloud_cb(void* h, const void* pts, size_t n, size_t stride, uint16_t* image_out, double* K9_out) {
    Cloud2RangeNode& node = *static_cast<Cloud2RangeNode*>(h);
    auto msg = std::make_shared<sensor_msgs::PointCloud2>();
    msg->width = static_cast<uint32_t>(n);
    msg->height = 1;
    msg->point_step = static_cast<uint32_t>(stride);
    msg->row_step = static_cast<uint32_t>(stride * n);
    msg->layout = 0;
    msg->data.assign(static_cast<const uint8_t*>(pts), static_cast<const uint8_t*>(pts) + n * stride);
    gsref::CameraFrame& cam = gsref::CameraFrame::get();
    const long before = cam.n_published;
    gsref::Topic<sensor_msgs::PointCloud2>::get().cb(msg);
    if (cam.n_published == before) return -1;
    std::memcpy(image_out, cam.image.data.data(), std::min(cam.image.data.size(), static_cast<size_t>(node.n_beams_) * node.n_cols_ * 2));
    if (K9_out) for (int i = 0; i < 9; ++i) K9_out[i] = cam.info.K[i];
    return 0;
}
void gsref_c2r_set_thresholds(void* h, double ground_thr, double start_thr) {
    Cloud2RangeNode& n = *static_cast<Cloud2RangeNode*>(h);
    n.ground_angle_threshold = ground_thr;
    n.start_angle_threshold = start_thr;
}
void gsref_c2r_repair_gaps(void* h, const uint16_t* range, int step, float depth_threshold, uint16_t* out) {
    Cloud2RangeNode& n = *static_cast<Cloud2RangeNode*>(h);
    out_u16(n.RepairGaps(mat_u16(range, n.n_beams_, n.n_cols_), step, depth_threshold), out);
}
void gsref_c2r_create_angle_img(void* h, const uint16_t* range, uint16_t* out) {
    Cloud2RangeNode& n = *static_cast<Cloud2RangeNode*>(h);
    out_u16(n.CreateAngleImg(mat_u16(range, n.n_beams_, n.n_cols_)), out);
}
int gsref_c2r_sg_smooth(void* h, const uint16_t* img, int window, uint16_t* out) {
    Cloud2RangeNode& n = *static_cast<Cloud2RangeNode*>(h);
    try {
        out_u16(n.SavitskyGolaySmoothing(mat_u16(img, n.n_beams_, n.n_cols_), window), out);
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}
int gsref_c2r_erase_ground_bfs(void* h, const uint16_t* range, const uint16_t* angle, double ground_thr, double start_thr, int kernel_size,
                               uint16_t* out) {
    Cloud2RangeNode& n = *static_cast<Cloud2RangeNode*>(h);
    try {
        out_u16(n.EraseGroundBFS(mat_u16(range, n.n_beams_, n.n_cols_), mat_u16(angle, n.n_beams_, n.n_cols_), ground_thr, start_thr, kernel_size), out);
    } catch (const std::logic_error&) {
        return 1;
    }
    return 0;
}

}  // extern "C"
