// gsref_ros.h — stand-in for roscpp, std_msgs, sensor_msgs, image_transport, cv_bridge, pcl, pcl_conversions, alfa_msg and
// boost::thread/shared_ptr.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Just enough surface for the reference's translation units to compile UNMODIFIED and for a test harness to drive them the
// way ROS would: NodeHandle::subscribe / ImageTransport::subscribeCamera keep the callback, publishers keep the last
// message, NodeHandle::param answers from a table the harness fills, ROS_INFO lines are kept in a small log (the reference
// reports its evaluation metrics only through ROS_INFO).  No threads are started (boost::thread is inert) and /dev/mem is never
// opened (the Alfa_GS constructor would switch to its FPGA path, src/alfa_gs.cpp:79): everything runs on the caller's thread.
// All state is thread_local: one node per thread.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <unistd.h>

#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "gsref_cv.h"

namespace boost {
using std::shared_ptr;
template <typename T, std::size_t N>
using array = std::array<T, N>;
class thread {
   public:
    template <typename F, typename O>
    thread(F, O) {}  // inert: the harness calls the callbacks itself
};
}  // namespace boost

namespace gsref {
struct State {
    std::vector<std::string> log;            // last ROS_INFO / ROS_ERROR lines
    long assert_failures = 0;                // ROS_ASSERT conditions that were false (the reference would abort)
    std::map<std::string, double> params;    // NodeHandle::param answers
    std::map<std::string, std::string> sparams;
};
inline State& state() {
    static thread_local State s;
    return s;
}
inline void log_message(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
inline void log_message(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    State& s = state();
    if (s.log.size() >= 32) s.log.erase(s.log.begin());
    s.log.emplace_back(buf);
}
template <typename M>
struct Topic {
    std::function<void(const std::shared_ptr<const M>&)> cb;
    M last;
    long n_published = 0;
    static Topic& get() {
        static thread_local Topic t;
        return t;
    }
};
inline int open_guard(const char* path, int flags, ...) {
    if (std::strcmp(path, "/dev/mem") == 0) return -1;  // never map FPGA registers from a test
    return ::open(path, flags);
}
}  // namespace gsref

#define ROS_INFO(...) gsref::log_message(__VA_ARGS__)
#define ROS_ERROR(...) gsref::log_message(__VA_ARGS__)
#define ROS_DEBUG(...) ((void)0)
#define ROS_ASSERT(cond)                                   \
    do {                                                   \
        if (!(cond)) ++gsref::state().assert_failures;     \
    } while (0)

namespace ros {
struct Time {
    double sec = 0;
    static Time now() { return Time(); }
};
inline void init(int&, char**, const std::string&) {}
inline bool ok() { return false; }
inline void spin() {}
namespace master {
inline bool check() { return true; }
}  // namespace master
class Publisher {
   public:
    template <typename M>
    void publish(const M& m) const {
        gsref::Topic<M>& t = gsref::Topic<M>::get();
        t.last = m;
        ++t.n_published;
    }
};
class Subscriber {};
class ServiceServer {};
template <typename Req, typename Res>
struct Service {
    std::function<bool(Req&, Res&)> cb;
    static Service& get() {
        static thread_local Service s;
        return s;
    }
};
class NodeHandle {
    template <typename T>
    static T lookup(const std::string& name, const T& def, std::true_type /*arithmetic*/) {
        const auto& p = gsref::state().params;
        const auto it = p.find(name);
        return it == p.end() ? def : static_cast<T>(it->second);
    }
    template <typename T>
    static T lookup(const std::string& name, const T& def, std::false_type) {
        const auto& p = gsref::state().sparams;
        const auto it = p.find(name);
        return it == p.end() ? def : T(it->second);
    }

   public:
    NodeHandle(const std::string& = std::string()) {}
    template <typename M, typename T>
    Subscriber subscribe(const std::string&, uint32_t, void (T::*fp)(const boost::shared_ptr<M const>&), T* obj) {
        gsref::Topic<M>::get().cb = [obj, fp](const std::shared_ptr<const M>& m) { (obj->*fp)(m); };
        return Subscriber();
    }
    template <typename M>
    Publisher advertise(const std::string&, uint32_t) {
        return Publisher();
    }
    template <typename T, typename Req, typename Res>
    ServiceServer advertiseService(const std::string&, bool (T::*fp)(Req&, Res&), T* obj) {
        Service<Req, Res>::get().cb = [obj, fp](Req& q, Res& s) { return (obj->*fp)(q, s); };
        return ServiceServer();
    }
    template <typename T>
    T param(const std::string& name, const T& def) const {
        return lookup<T>(name, def, std::is_arithmetic<T>());
    }
};
}  // namespace ros

namespace std_msgs {
struct Header {
    uint32_t seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
}  // namespace std_msgs

namespace sensor_msgs {
struct Image {
    std_msgs::Header header;
    uint32_t height = 0, width = 0;
    std::string encoding;
    uint8_t is_bigendian = 0;
    uint32_t step = 0;
    std::vector<uint8_t> data;
};
typedef boost::shared_ptr<Image> ImagePtr;
typedef boost::shared_ptr<Image const> ImageConstPtr;
struct CameraInfo {
    std_msgs::Header header;
    uint32_t height = 0, width = 0;
    std::string distortion_model;
    boost::array<double, 9> K{};
};
typedef boost::shared_ptr<CameraInfo const> CameraInfoConstPtr;
// Records of point_step bytes: x, y, z floats at offsets 0 / 4 / 8; `layout` says what follows — 0: intensity at 12
// (16-byte XYZI records), 1: PCL's 32-byte PointXYZRGB (bgra bytes at 16)
struct PointCloud2 {
    std_msgs::Header header;
    uint32_t height = 0, width = 0;
    uint32_t point_step = 0, row_step = 0;
    int layout = 0;
    std::vector<uint8_t> data;
    bool is_dense = true;
};
typedef boost::shared_ptr<PointCloud2 const> PointCloud2ConstPtr;
}  // namespace sensor_msgs

namespace gsref {
struct CameraFrame {
    sensor_msgs::Image image;
    sensor_msgs::CameraInfo info;
    long n_published = 0;
    std::function<void(const sensor_msgs::ImageConstPtr&, const sensor_msgs::CameraInfoConstPtr&)> cb;
    static CameraFrame& get() {
        static thread_local CameraFrame f;
        return f;
    }
};
}  // namespace gsref

namespace image_transport {
class CameraPublisher {
   public:
    void publish(const sensor_msgs::Image& img, const sensor_msgs::CameraInfo& ci) const {
        gsref::CameraFrame& f = gsref::CameraFrame::get();
        f.image = img;
        f.info = ci;
        ++f.n_published;
    }
};
class CameraSubscriber {};
class ImageTransport {
   public:
    explicit ImageTransport(const ros::NodeHandle&) {}
    CameraPublisher advertiseCamera(const std::string&, uint32_t) { return CameraPublisher(); }
    template <typename T>
    CameraSubscriber subscribeCamera(const std::string&, uint32_t,
                                     void (T::*fp)(const sensor_msgs::ImageConstPtr&, const sensor_msgs::CameraInfoConstPtr&), T* obj) {
        gsref::CameraFrame::get().cb = [obj, fp](const sensor_msgs::ImageConstPtr& i, const sensor_msgs::CameraInfoConstPtr& c) { (obj->*fp)(i, c); };
        return CameraSubscriber();
    }
};
}  // namespace image_transport

namespace cv_bridge {
class Exception : public std::runtime_error {
   public:
    explicit Exception(const std::string& w) : std::runtime_error(w) {}
};
class CvImage {
   public:
    std_msgs::Header header;
    std::string encoding;
    cv::Mat image;
    CvImage() {}
    CvImage(const std_msgs::Header& h, const std::string& enc, const cv::Mat& img) : header(h), encoding(enc), image(img) {}
    sensor_msgs::ImagePtr toImageMsg() const {
        auto msg = std::make_shared<sensor_msgs::Image>();
        msg->header = header;
        msg->height = static_cast<uint32_t>(image.rows);
        msg->width = static_cast<uint32_t>(image.cols);
        msg->encoding = encoding;
        msg->step = static_cast<uint32_t>(image.cols * image.slot());
        msg->data.resize(static_cast<size_t>(msg->step) * image.rows);
        if (!msg->data.empty()) std::memcpy(msg->data.data(), image.raw(), msg->data.size());
        return msg;
    }
};
typedef boost::shared_ptr<CvImage const> CvImageConstPtr;
inline CvImageConstPtr toCvShare(const sensor_msgs::ImageConstPtr& msg, const std::string& = std::string()) {
    if (msg->encoding != "mono16") throw Exception("gsref: only mono16 images");
    auto out = std::make_shared<CvImage>();
    out->header = msg->header;
    out->encoding = msg->encoding;
    out->image = cv::Mat::zeros(static_cast<int>(msg->height), static_cast<int>(msg->width), CV_16UC1);
    if (!msg->data.empty()) std::memcpy(out->image.raw(), msg->data.data(), msg->data.size());
    return out;
}
}  // namespace cv_bridge

namespace pcl {
struct PCLHeader {
    uint32_t seq = 0;
    uint64_t stamp = 0;
    std::string frame_id;
};
// PCL's layouts: 16-byte PointXYZ, 32-byte PointXYZI (intensity at 16) and PointXYZRGB (b, g, r, a at 16)
struct PointXYZ {
    float x = 0, y = 0, z = 0, pad = 1.0f;
};
struct PointXYZI {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    float intensity = 0, pad2[3] = {0, 0, 0};
};
struct PointXYZRGB {
    float x = 0, y = 0, z = 0, pad = 1.0f;
    uint8_t b = 0, g = 0, r = 0, a = 255;
    float pad2[3] = {0, 0, 0};
};
template <typename T>
class PointCloud {
   public:
    PCLHeader header;
    std::vector<T> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    typedef boost::shared_ptr<PointCloud<T>> Ptr;
    typedef boost::shared_ptr<const PointCloud<T>> ConstPtr;
    size_t size() const { return points.size(); }
    const T& operator[](size_t i) const { return points[i]; }
    T& operator[](size_t i) { return points[i]; }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
    typename std::vector<T>::const_iterator begin() const { return points.begin(); }
    typename std::vector<T>::const_iterator end() const { return points.end(); }
    void push_back(const T& p) { points.push_back(p); }
    Ptr makeShared() const { return Ptr(new PointCloud<T>(*this)); }
};
namespace detail {
inline void extra(PointXYZ&, const uint8_t*, const sensor_msgs::PointCloud2&) {}
inline void extra(PointXYZI& p, const uint8_t* rec, const sensor_msgs::PointCloud2& m) {
    if (m.layout == 0 && m.point_step >= 16) std::memcpy(&p.intensity, rec + 12, 4);
}
inline void extra(PointXYZRGB& p, const uint8_t* rec, const sensor_msgs::PointCloud2& m) {
    if (m.layout == 1 && m.point_step >= 20) { p.b = rec[16]; p.g = rec[17]; p.r = rec[18]; p.a = rec[19]; }
}
}  // namespace detail
template <typename T>
void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<T>& cloud) {
    const size_t n = static_cast<size_t>(msg.width) * msg.height;
    cloud.points.assign(n, T());
    cloud.width = msg.width;
    cloud.height = msg.height;
    for (size_t i = 0; i < n; ++i) {
        const uint8_t* rec = msg.data.data() + i * msg.point_step;
        T& p = cloud.points[i];
        std::memcpy(&p.x, rec, 4);
        std::memcpy(&p.y, rec + 4, 4);
        std::memcpy(&p.z, rec + 8, 4);
        detail::extra(p, rec, msg);
    }
}
template <typename T>
void toROSMsg(const PointCloud<T>& cloud, sensor_msgs::PointCloud2& msg) {
    msg.width = static_cast<uint32_t>(cloud.points.size());
    msg.height = 1;
    msg.point_step = sizeof(T);
    msg.row_step = msg.point_step * msg.width;
    msg.layout = 1;
    msg.data.resize(cloud.points.size() * sizeof(T));
    if (!cloud.points.empty()) std::memcpy(msg.data.data(), cloud.points.data(), msg.data.size());
}
}  // namespace pcl

namespace pcl_conversions {
inline void toPCL(const std_msgs::Header& h, pcl::PCLHeader& out) {
    out.seq = h.seq;
    out.frame_id = h.frame_id;
}
}  // namespace pcl_conversions

namespace alfa_msg {
struct ConfigMessage {
    double config = 0;
    std::string config_name;
};
struct MetricMessage {
    double metric = 0;
    std::string units, metric_name;
};
struct AlfaMetrics {
    std::vector<MetricMessage> metrics;
    std::string message_tag;
};
struct AlfaAlivePing {
    std::string node_name, node_type, config_service_name, config_tag;
    std::vector<ConfigMessage> default_configurations;
    int current_status = 0;
};
struct AlfaConfigure {
    struct Request {
        std::vector<ConfigMessage> configurations;
    };
    struct Response {
        int return_status = 0;
    };
};
}  // namespace alfa_msg

// Alfa_GS's constructor maps FPGA registers when /dev/mem opens (src/alfa_gs.cpp:79); keep it on the software path.
// Defined last, after every standard header, so that only call expressions in the reference's own sources are affected.
#define open(...) gsref::open_guard(__VA_ARGS__)
