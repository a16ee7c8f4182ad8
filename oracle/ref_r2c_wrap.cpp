// ref_r2c_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// The reference's stand-alone range-image -> cloud node, src/range2cloud_node.cpp, compiled verbatim (pulled in as it lies
// under /root/reference; `main` renamed away) against the stand-in headers of oracle/shim, into oracle/_ref/libgs_ref.so.
// A mono16 image + CameraInfo go to the callback the node subscribed (Range2CloudNode::CameraCb, :36-99) and the published
// pcl::PointCloud<pcl::PointXYZ> is read back.
#include "shim/gsref_ros.h"

#define main gsref_range2cloud_main
#include "src/range2cloud_node.cpp"  // -I/root/reference
#undef main

extern "C" {

void* gsref_r2c_create(void) { return new cloud2range::Range2CloudNode(ros::NodeHandle("~")); }
void gsref_r2c_destroy(void* h) { delete static_cast<cloud2range::Range2CloudNode*>(h); }
// K9: CameraInfo.K as the cloud2range node publishes it (min/max angle, min/max range, d_azimuth, d_altitude);
// points_out: 16-byte pcl::PointXYZ records.  Returns the number of points of the published cloud (-1: nothing published).
long gsref_r2c_camera_cb(void* /*h*/, const uint16_t* image, int rows, int cols, const double* K9, void* points_out, size_t cap) {
    auto img = std::make_shared<sensor_msgs::Image>();
    img->height = static_cast<uint32_t>(rows);
    img->width = static_cast<uint32_t>(cols);
    img->encoding = "mono16";
    img->step = static_cast<uint32_t>(cols * 2);
    img->data.assign(reinterpret_cast<const uint8_t*>(image), reinterpret_cast<const uint8_t*>(image) + static_cast<size_t>(rows) * cols * 2);
    auto ci = std::make_shared<sensor_msgs::CameraInfo>();
    ci->height = static_cast<uint32_t>(rows);
    ci->width = static_cast<uint32_t>(cols);
    for (int i = 0; i < 9; ++i) ci->K[i] = K9[i];
    typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
    gsref::Topic<Cloud>& top = gsref::Topic<Cloud>::get();
    const long before = top.n_published;
    gsref::CameraFrame::get().cb(img, ci);
    if (top.n_published == before) return -1;
    const size_t np = top.last.points.size();
    if (points_out) std::memcpy(points_out, top.last.points.data(), std::min(cap, np) * sizeof(pcl::PointXYZ));
    return static_cast<long>(np);
}

}  // extern "C"
