// host_api_test.cpp — the reference's C++ API surface (include/gs_host.hpp) exercised the way the reference's own
// process_pointcloud does (src/alfa_gs.cpp:175-244), stage by stage and fused, and compared with the CPU oracle
// (oracle/libgs_oracle.so — test infrastructure).  Built and run by tests/test_host_api.py on a GPU box.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <stdexcept>
#include <vector>

#include "../include/gs_host.hpp"

extern "C" {
int gso_process_frame(const gs_params* p, const void* points, size_t n, size_t stride, uint16_t* range_out, uint16_t* angle_out,
                      uint16_t* no_ground_out, float* xyz4_out, uint32_t* n_out, double* stage_ms);
int gso_create_labeled_range_image(const gs_params* p, const void* points, size_t n, uint16_t* img);
int gso_evaluate(const gs_params* p, const uint16_t* og, const uint16_t* seg, const uint16_t* lab, gs_eval* out);
int gso_label_one_component(const gs_params* p, int label, int row, int col, float thr, const uint16_t* range, const uint16_t* angle,
                            uint8_t* label_image);
}

static int failures = 0;
#define CHECK(cond, what)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            std::printf("FAIL: %s (%s:%d)\n", what, __FILE__, __LINE__); \
            ++failures;                                     \
        }                                                   \
    } while (0)

// a crude synthetic scan: ground plane + a wall ring, one return per beam with jitter
template <class PointT>
static std::vector<PointT> make_cloud(const gs_params& p, unsigned seed) {
    gs_derived d;
    gs_derive(&p, &d);
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> jit(-0.3, 0.3), u01(0.0, 1.0);
    std::vector<PointT> cloud;
    for (int c = 0; c < d.n_cols; ++c)
        for (int r = 0; r < d.n_rows; ++r) {
            if (u01(rng) < 0.08) continue;
            double alt = p.min_angle + (d.n_rows - 1 - r + jit(rng)) * d.d_altitude;
            alt = std::min(std::max(alt, p.min_angle + 0.02 * d.d_altitude), p.max_angle - 0.02 * d.d_altitude);
            const double az = (c + jit(rng)) * d.d_azimuth;
            double t = 40.0 / std::cos(alt);
            if (std::sin(alt) < -1e-3) t = std::min(t, -1.73 / std::sin(alt));
            t += 0.02 * (u01(rng) - 0.5);
            if (t < p.min_range + 0.05 || t > p.max_range - 0.05) continue;
            PointT q;
            std::memset(&q, 0, sizeof(q));
            q.x = static_cast<float>(std::cos(alt) * std::cos(az) * t);
            q.y = static_cast<float>(std::cos(alt) * std::sin(az) * t);
            q.z = static_cast<float>(std::sin(alt) * t);
            cloud.push_back(q);
        }
    return cloud;
}

int main() {
    // ---- Alfa_GS, the executable the reference builds: its constructor defaults (64 x 1800)
    gs::Alfa_GS node;
    const gs_params p = node.context()->params();
    const int H = node.n_rows(), W = node.n_cols();
    CHECK(H == 64 && W == 1800, "Alfa_GS default shape");
    CHECK(node.cinfo().K[4] == node.context()->derived().d_azimuth, "cinfo K[4] = d_azimuth");
    const std::vector<gs::PointXYZRGB> cloud = make_cloud<gs::PointXYZRGB>(p, 7);
    std::vector<uint16_t> o_range(H * W), o_angle(H * W), o_res(H * W);
    std::vector<float> o_xyz(static_cast<size_t>(H) * W * 4);
    uint32_t o_n = 0;
    gso_process_frame(&p, cloud.data(), cloud.size(), sizeof(gs::PointXYZRGB), o_range.data(), o_angle.data(), o_res.data(), o_xyz.data(), &o_n, nullptr);

    // stage by stage, as Alfa_GS::process_pointcloud (alfa_gs.cpp:196-239)
    gs::Image<uint16_t> range_image = node.CreateRangeImage(cloud);
    CHECK(range_image.data == o_range, "CreateRangeImage");
    gs::Image<uint16_t> angle_image = node.CreateAngleImg(range_image);
    CHECK(angle_image.data == o_angle, "CreateAngleImg");
    gs::Image<uint16_t> ma = node.MovingAverageSmoothing(angle_image, 3);  // computed and discarded by the reference
    CHECK(ma.data == angle_image.data, "MovingAverageSmoothing is the documented no-op");
    gs::Image<uint16_t> no_ground = node.EraseBFS(range_image, angle_image, p.ground_angle_threshold, p.start_angle_threshold, p.window_size);
    CHECK(no_ground.data == o_res, "EraseBFS");
    std::vector<gs::PointXYZRGB> seg = node.CameraCb(no_ground, node.cinfo());
    CHECK(seg.size() == o_n, "CameraCb point count");
    bool same = seg.size() == o_n;
    for (size_t i = 0; same && i < seg.size(); ++i)
        same = seg[i].x == o_xyz[4 * i] && seg[i].y == o_xyz[4 * i + 1] && seg[i].z == o_xyz[4 * i + 2] && seg[i].a == 255;
    CHECK(same, "CameraCb points");
    // fused
    gs::Alfa_GS::Frame f = node.process_pointcloud(cloud);
    CHECK(f.range_image.data == o_range && f.angle_image.data == o_angle && f.no_ground_image.data == o_res && f.seg_cloud.size() == o_n,
          "process_pointcloud");

    // evaluation helpers (alfa_gs.cpp:831-995): colour every third low point as ground, compare with the oracle
    {
        std::vector<gs::PointXYZRGB> labelled = cloud;
        for (size_t i = 0; i < labelled.size(); ++i)
            if (labelled[i].z < -1.5f && i % 3 == 0) { labelled[i].r = 255; labelled[i].g = 0; labelled[i].b = 255; }
        gs::Image<uint16_t> lab = node.CreateLabeledRangeImage(labelled);
        std::vector<uint16_t> o_lab(H * W);
        gso_create_labeled_range_image(&p, labelled.data(), labelled.size(), o_lab.data());
        CHECK(lab.data == o_lab, "CreateLabeledRangeImage");
        gs_eval e = node.CheckNumberOfDetectedRIdGnd(range_image, no_ground, lab), o_e;
        gso_evaluate(&p, o_range.data(), o_res.data(), o_lab.data(), &o_e);
        CHECK(e.true_positive == o_e.true_positive && e.false_positive == o_e.false_positive && e.true_negative == o_e.true_negative &&
                  e.false_negative == o_e.false_negative && e.labeled_gnd == o_e.labeled_gnd && e.seg_gnd == o_e.seg_gnd && e.iou_g == o_e.iou_g &&
                  e.recall_g == o_e.recall_g && e.recall_mo == o_e.recall_mo,
              "CheckNumberOfDetectedRIdGnd");
    }

    // error behaviour of the reference: std::logic_error for even windows / unknown Savitzky-Golay sizes
    bool threw = false;
    try { node.GetUniformKernel(6, 0); } catch (const std::logic_error&) { threw = true; }
    CHECK(threw, "GetUniformKernel(6) throws std::logic_error");
    threw = false;
    try { node.GetSavitskyGolayKernel(13); } catch (const std::logic_error&) { threw = true; }
    CHECK(threw, "GetSavitskyGolayKernel(13) throws std::logic_error");
    threw = false;
    try { node.SavitskyGolaySmoothing(angle_image, 13); } catch (const std::logic_error&) { threw = true; }
    CHECK(threw, "SavitskyGolaySmoothing(13) throws std::logic_error");
    threw = false;
    try { node.EraseBFS(range_image, angle_image, 0.1, 0.5, 8); } catch (const std::logic_error&) { threw = true; }
    CHECK(threw, "EraseBFS(kernel_size 8) throws std::logic_error");
    {
        const gs::Image<uint8_t> k = node.GetUniformKernel(9, 0);  // bytes 00 00 40 20 00 ...: SURVEY.md A.4
        CHECK(k.at(0, 0) == 0 && k.at(2, 0) == 64 && k.at(3, 0) == 32 && k.at(8, 0) == 0, "GetUniformKernel byte pattern");
    }

    // ---- the Label interface
    {
        gs::cloud2range::Label labeler(H, W, node.context());
        CHECK(labeler.WrapCols(-1) == W - 1 && labeler.WrapCols(static_cast<int16_t>(W)) == 0 && labeler.WrapCols(5) == 5, "Label::WrapCols");
        int r = H - 1;
        while (r > 0 && range_image.at(r, 10) == 0) --r;
        labeler.LabelOneComponent(1, r, 10, static_cast<float>(p.ground_angle_threshold), range_image, angle_image);
        std::vector<uint8_t> want(static_cast<size_t>(H) * W, 0);
        gso_label_one_component(&p, 1, r, 10, static_cast<float>(p.ground_angle_threshold), o_range.data(), o_angle.data(), want.data());
        CHECK(labeler.GetLabelImage().data == want, "Label::LabelOneComponent");
        CHECK(labeler.CheckLabelAt(r, 10) == 1, "Label::CheckLabelAt");
    }

    // ---- UpdateSegmentationSettings refreshes the derived values
    {
        std::vector<double> cfg(14, 0.0);
        cfg[0] = 1; cfg[1] = 32; cfg[2] = 600; cfg[3] = 18000; cfg[4] = -0.535293; cfg[5] = 0.186227; cfg[6] = 0.5; cfg[7] = 100.0;
        cfg[10] = 0.523598; cfg[11] = 0.087266; cfg[13] = 5;
        gs::Alfa_GS n2;
        n2.UpdateSegmentationSettings(cfg);
        CHECK(n2.n_rows() == 32 && n2.n_cols() == 1800 && n2.cinfo().height == 32, "UpdateSegmentationSettings");
    }

    // ---- the two-node packaging: Cloud2RangeNode::CloudCb -> Range2CloudNode::CameraCb
    {
        gs::Cloud2RangeNode c2r;
        const gs_params q = c2r.context()->params();
        const gs_derived d = c2r.context()->derived();
        const std::vector<gs::PointXYZI> pc = make_cloud<gs::PointXYZI>(q, 11);
        gs::Image<uint16_t> img = c2r.CloudCb(pc);
        std::vector<uint16_t> w_res(static_cast<size_t>(d.n_rows) * d.n_cols);
        std::vector<float> w_xyz(w_res.size() * 4);
        uint32_t w_n = 0;
        gso_process_frame(&q, pc.data(), pc.size(), sizeof(gs::PointXYZI), nullptr, nullptr, w_res.data(), w_xyz.data(), &w_n, nullptr);
        CHECK(img.data == w_res, "Cloud2RangeNode::CloudCb");
        gs::Range2CloudNode r2c;
        std::vector<gs::PointXYZ> out = r2c.CameraCb(img, c2r.cinfo());
        bool ok = out.size() == w_n;
        for (size_t i = 0; ok && i < out.size(); ++i)  // K[4] / K[5] are used as published (range2cloud_node.cpp:51-58): bit-exact
            for (int k = 0; k < 3; ++k) ok = ok && (&out[i].x)[k] == w_xyz[4 * i + k];
        CHECK(ok, "Range2CloudNode::CameraCb");
        // a foreign CameraInfo: d_azimuth / d_altitude that this library would not derive itself must be honoured
        gs::CameraInfo foreign = c2r.cinfo();
        foreign.K[4] *= 1.0 + 3e-16;
        foreign.K[5] *= 1.0 - 2e-16;
        std::vector<gs::PointXYZ> out2 = r2c.CameraCb(img, foreign);
        bool ok2 = out2.size() == out.size();
        size_t k_img = 0;
        for (int r = 0; ok2 && r < img.rows; ++r)
            for (int c = 0; ok2 && c < img.cols; ++c) {
                const uint16_t e = img.at(r, c);
                if (!e) continue;
                const double range = (static_cast<double>(e - 1) / 65534) * (foreign.K[3] - foreign.K[2]) + foreign.K[2];
                const double alt = (img.rows - r) * foreign.K[5] + foreign.K[0], az = c * foreign.K[4];
                ok2 = out2[k_img].x == static_cast<float>(std::cos(alt) * std::cos(az) * range) &&
                      out2[k_img].y == static_cast<float>(std::cos(alt) * std::sin(az) * range) &&
                      out2[k_img].z == static_cast<float>(std::sin(alt) * range);
                ++k_img;
            }
        CHECK(ok2, "Range2CloudNode::CameraCb with a foreign CameraInfo");
        bool threw = false;
        try {
            gs::Image<uint16_t> small(img.rows - 1, img.cols);
            r2c.CameraCb(small, c2r.cinfo());
        } catch (const std::logic_error&) {
            threw = true;
        }
        CHECK(threw, "Range2CloudNode::CameraCb rejects an image that is not height x width");

        // ---- gs::MultiGpu: two shards (on one GPU when only one is visible) == one context, frame by frame
        std::vector<gs::PointXYZI> batch;
        std::vector<int64_t> offs(1, 0);
        const int F = 5;
        for (int f = 0; f < F; ++f) {
            const std::vector<gs::PointXYZI> fr = make_cloud<gs::PointXYZI>(q, 100 + f);
            batch.insert(batch.end(), fr.begin(), fr.end());
            offs.push_back(static_cast<int64_t>(batch.size()));
        }
        const size_t P = static_cast<size_t>(d.n_rows) * d.n_cols;
        gs::MultiGpu gpus(q, std::vector<int>{0, 0});
        CHECK(gpus.n_gpus() == 2, "gs::MultiGpu shard count");
        std::vector<uint16_t> m_res(F * P);
        std::vector<float> m_xyz(F * P * 4);
        std::vector<int64_t> first(F);
        std::vector<uint32_t> count(F);
        gpus.process_batch_host(batch.data(), sizeof(gs::PointXYZI), offs.data(), F, m_res.data(), m_xyz.data(), F * P, first.data(), count.data());
        bool okm = true;
        for (int f = 0; okm && f < F; ++f) {
            gso_process_frame(&q, batch.data() + offs[f], static_cast<size_t>(offs[f + 1] - offs[f]), sizeof(gs::PointXYZI), nullptr, nullptr, w_res.data(),
                              w_xyz.data(), &w_n, nullptr);
            okm = std::equal(w_res.begin(), w_res.end(), m_res.begin() + f * P) && count[f] == w_n &&
                  std::equal(w_xyz.begin(), w_xyz.begin() + 4 * static_cast<size_t>(w_n), m_xyz.begin() + 4 * first[f]);
        }
        CHECK(okm, "gs::MultiGpu::process_batch_host");
    }
    if (failures == 0) std::printf("host api ok\n");
    return failures ? 1 : 0;
}
