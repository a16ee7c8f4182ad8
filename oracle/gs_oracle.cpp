// gs_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain, single-threaded C++ restatement of the reference's algorithm for the path
//   cloud -> range image -> angle image -> ground mask -> cloud
// used ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// as the checker.  The product (libgs_b200.so) never links, loads or calls anything in oracle/.
//
// Semantics: the "typed-accessor" contract of SURVEY.md appendix A — every cv::Mat::at<T>(r,c) of
// the reference means logical element (r,c); the deterministic byte-pattern quirk of
// GetUniformKernel (alfa_gs.cpp:630-639) is kept literally.
//
// Pinning status (see DESIGN.md "Oracle"): PINNED.  The reference's own translation units — src/alfa_gs.cpp,
//   src/alfa_node.cpp, Labeler/label.cpp, src/cloud2range_node.cpp, src/range2cloud_node.cpp, lib/utils.h — are
//   compiled verbatim from /root/reference into oracle/_ref/libgs_ref.so against stand-in ROS / PCL / OpenCV
//   headers (oracle/shim, recipe oracle/build.py) and every function below is compared with the reference member
//   it restates, bit for bit, stage by stage and as whole frames driven the way ROS drives the nodes
//   (tests/test_reference_verbatim.py, tests/test_oracle_vs_ref.py).  The two OpenCV calls (cv::dilate with the
//   GetUniformKernel bytes, cv::filter2D with the Savitzky-Golay kernel) — here and in the stand-in headers — are
//   pinned against OpenCV 4.13 golden vectors in tests/golden/.
//
// Build: g++ -O2 -std=c++14 -ffp-contract=off -fPIC -shared (oracle/build.py).  No FMA contraction:
// the x86-64 reference build has none either.
//
// All citations are relative to /root/reference.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <thread>
#include <vector>

#include "../include/gs_b200.h" /* POD parameter structs only */

namespace {

// ---- lib/utils.h:8-54 ---------------------------------------------------------------------------
inline double point_altitude(double x, double y, double z) { return std::atan2(z, std::hypot(x, y)); }  // utils.h:8-11
inline double point_azimuth(double x, double y) {                                                        // utils.h:24-28
    const double a = std::atan2(y, x);
    return (y >= 0 ? a : a + M_PI * 2);
}
inline double point_range(double x, double y, double z) { return std::sqrt(x * x + y * y + z * z); }    // utils.h:41-43
inline double rad_from_deg(double deg) { return deg / 180 * M_PI; }                                       // utils.h:54

struct Dims {
    int H, W;
    double d_az, d_alt;
};

// src/alfa_gs.cpp:48-50 (== src/cloud2range_node.cpp:308-312)
Dims dims_of(const gs_params& p) {
    Dims d;
    d.H = p.n_beams;
    d.W = p.sample_freq * 60 / p.rpm;  // integer division
    d.d_az = rad_from_deg(360.0 / d.W);
    d.d_alt = (p.max_angle - p.min_angle) / (p.n_beams - 1);
    return d;
}

inline bool finite3(float x, float y, float z) { return std::isfinite(x) && std::isfinite(y) && std::isfinite(z); }

// Edge test cut-off: label.cpp:99 compares an integer-valued double with `float thr * 100` (float product).
int k_of_threshold(double thr) {
    const float t = static_cast<float>(thr);  // label.h:41 takes the threshold as float
    const float T = t * 100;                  // label.cpp:99, float * int -> float
    return static_cast<int>(std::ceil(static_cast<double>(T))) - 1;  // diff < T  <=>  diff <= ceil(T)-1
}
// Seed gate: alfa_gs.cpp:733 `angle > start*100` in double  ->  accepted <=> angle <= floor(start*100)
int s_of_threshold(double start) { return static_cast<int>(std::floor(start * 100)); }

// The label image + BFS of Labeler/label.{h,cpp}, logical indexing, wide labels (Q9).
struct LabelImage {
    int rows, cols;
    std::vector<int32_t> lab;
    LabelImage(int r, int c) : rows(r), cols(c), lab(static_cast<size_t>(r) * c, 0) {}  // label.cpp:10
    int32_t& at(int r, int c) { return lab[static_cast<size_t>(r) * cols + c]; }
    int wrap_col(int c) const {  // label.cpp:33-45
        if (c < 0) return c + cols;
        if (c >= cols) return c - cols;
        return c;
    }
    // label.cpp:48-104.  Neighbour order (label.cpp:13-20): up, down, left, right.
    void label_one_component(int32_t label, int row, int col, float thr, const uint16_t* range,
                             const uint16_t* angle) {
        static const int dr[4] = {-1, +1, 0, 0};
        static const int dc[4] = {0, 0, -1, +1};
        std::deque<std::pair<int, int>> q;
        q.emplace_back(row, col);
        const float T = thr * 100;  // label.cpp:99
        while (!q.empty()) {
            const std::pair<int, int> cur = q.front();
            q.pop_front();
            if (at(cur.first, cur.second) > 0) continue;  // label.cpp:61-65
            at(cur.first, cur.second) = label;            // label.cpp:67
            const size_t ci = static_cast<size_t>(cur.first) * cols + cur.second;
            if (range[ci] == 0) continue;  // label.cpp:70-74 (ushort < 0.001f <=> == 0)
            for (int k = 0; k < 4; ++k) {
                const int nr = cur.first + dr[k];
                if (nr < 0 || nr >= rows) continue;  // label.cpp:77 (uint16 wrap makes -1 >= rows)
                const int nc = wrap_col(cur.second + dc[k]);  // label.cpp:82
                if (at(nr, nc) > 0) continue;                 // label.cpp:83-88
                const size_t ni = static_cast<size_t>(nr) * cols + nc;
                const double diff = std::fabs(static_cast<int>(angle[ci]) - static_cast<int>(angle[ni]));  // label.cpp:90
                if (diff < T) q.emplace_back(nr, nc);  // label.cpp:99-101
            }
        }
    }
};

// GetUniformKernel (alfa_gs.cpp:630-639) executed literally on bytes: two float 1.0 stores into a
// ks-byte u8 column (the second overruns by 3 bytes), then `/= 2` with OpenCV's u8 rounding
// (round-half-even, saturate).  Returns the ks kernel bytes.
std::vector<uint8_t> uniform_kernel_bytes(int ks) {
    std::vector<uint8_t> buf(static_cast<size_t>(ks) + 3, 0);  // +3: the overrun lands in slack
    const float one = 1.0f;
    std::memcpy(&buf[0], &one, 4);       // kernel.at<float>(0,0) = 1            (:635)
    std::memcpy(&buf[ks - 1], &one, 4);  // kernel.at<float>(window_size-1,0)=1  (:636), step = 1 byte
    std::vector<uint8_t> k(ks);
    for (int i = 0; i < ks; ++i) k[i] = static_cast<uint8_t>(std::nearbyint(buf[i] / 2.0));  // kernel /= 2 (:637)
    return k;
}

// Row offsets of the effective structuring element of cv::dilate(label, dilated, kernel) (alfa_gs.cpp:750):
// non-zero kernel bytes, anchor = centre; an all-zero kernel degenerates to the single tap at kernel
// coordinate (0,0) (OpenCV preprocess2DKernel: nz==0 -> one default Point) — pinned by tests/golden.
std::vector<int> dilate_offsets_ks(int ks) {
    const std::vector<uint8_t> k = uniform_kernel_bytes(ks);
    std::vector<int> off;
    for (int i = 0; i < ks; ++i)
        if (k[i]) off.push_back(i - ks / 2);
    if (off.empty()) off.push_back(0 - ks / 2);
    return off;
}
std::vector<int> dilate_offsets(int kernel_size_arg) {
    return dilate_offsets_ks(std::max(kernel_size_arg - 2, 3));  // alfa_gs.cpp:747
}

// Savitzky-Golay taps, alfa_gs.cpp:653-699: integer numerators, then `kernel /= norm`, which OpenCV executes
// as Mat::convertTo(kernel, -1, alpha = 1./norm): every tap becomes numerator * (float)(1.0 / norm).
bool sg_taps(int ws, std::vector<float>& taps) {
    static const float k5[] = {-3, 12, 17, 12, -3};
    static const float k7[] = {-2, 3, 6, 7, 6, 3, -2};
    static const float k9[] = {-21, 14, 39, 54, 59, 54, 39, 14, -21};
    static const float k11[] = {-36, 9, 44, 69, 84, 89, 84, 69, 44, 9, -36};
    const float* src;
    float norm;
    switch (ws) {
        case 5: src = k5; norm = 35.0f; break;
        case 7: src = k7; norm = 21.0f; break;
        case 9: src = k9; norm = 231.0f; break;
        case 11: src = k11; norm = 429.0f; break;
        default: return false;  // alfa_gs.cpp:645-649 throws std::logic_error
    }
    taps.resize(ws);
    const float alpha = static_cast<float>(1.0 / static_cast<double>(norm));
    for (int i = 0; i < ws; ++i) taps[i] = src[i] * alpha;
    return true;
}

inline int reflect101(int i, int n) {  // cv::BORDER_REFLECT101: gfedcb|abcdefgh|gfedcba
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * n - 2 - i;
    }
    return i;
}

}  // namespace

extern "C" {

int gso_derive(const gs_params* p, gs_derived* out) {
    if (!p || !out || p->n_beams < 2 || p->rpm <= 0 || p->sample_freq <= 0) return GS_EINVAL;
    const Dims d = dims_of(*p);
    out->n_rows = d.H;
    out->n_cols = d.W;
    out->d_azimuth = d.d_az;
    out->d_altitude = d.d_alt;
    out->k_ground = k_of_threshold(p->ground_angle_threshold);
    out->s_start = s_of_threshold(p->start_angle_threshold);
    if (std::max(p->window_size - 2, 3) % 2 == 0) return GS_EINVAL;  // alfa_gs.cpp:631-633
    const std::vector<int> off = dilate_offsets(p->window_size);
    out->n_dilate = static_cast<int>(off.size());
    out->dilate_off[0] = off[0];
    out->dilate_off[1] = off.size() > 1 ? off[1] : off[0];
    return GS_OK;
}

int gso_k_of_threshold(double thr) { return k_of_threshold(thr); }
int gso_s_of_threshold(double thr) { return s_of_threshold(thr); }

// GetUniformKernel bytes and effective offsets, exposed for the golden-vector tests.
int gso_uniform_kernel(int ks, uint8_t* bytes_out) {
    if (ks % 2 == 0) return GS_EINVAL;  // alfa_gs.cpp:631-633
    const std::vector<uint8_t> k = uniform_kernel_bytes(ks);
    std::memcpy(bytes_out, k.data(), k.size());
    return GS_OK;
}

// cv::dilate(in, out, GetUniformKernel(ks, CV_8U)) on a u8 image (alfa_gs.cpp:748-750), for the golden-vector test.
int gso_dilate(int rows, int cols, int ks, const uint8_t* in, uint8_t* out) {
    if (ks % 2 == 0) return GS_EINVAL;
    const std::vector<int> off = dilate_offsets_ks(ks);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            uint8_t m = 0;
            for (size_t k = 0; k < off.size(); ++k) {
                const int rr = r + off[k];
                if (rr >= 0 && rr < rows) m = std::max(m, in[static_cast<size_t>(rr) * cols + c]);
            }
            out[static_cast<size_t>(r) * cols + c] = m;
        }
    return GS_OK;
}

int gso_sg_kernel(int ws, float* taps_out) {
    std::vector<float> t;
    if (!sg_taps(ws, t)) return GS_EINVAL;
    std::memcpy(taps_out, t.data(), t.size() * sizeof(float));
    return GS_OK;
}

// ---- R1 / R1' : Alfa_GS::CreateRangeImage (alfa_gs.cpp:247-298) and the projection loop of
// Cloud2RangeNode::CloudCb (cloud2range_node.cpp:396-455) ---------------------------------------
// near_edge (may be NULL) counts accepted points whose altitude or azimuth lies within 1e-5 rad of
// a cell boundary (the north-star tolerance set).
int gso_create_range_image(const gs_params* p, const void* points, size_t n, size_t stride,
                           uint16_t* img, uint64_t* near_edge, uint64_t* dropped) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    const size_t P = static_cast<size_t>(H) * W;
    std::memset(img, 0, P * sizeof(uint16_t));  // Mat::zeros (alfa_gs.cpp:253)
    const double half_d_alt = d.d_alt / 2, half_d_az = d.d_az / 2;  // :249-250
    uint64_t ne = 0, drop = 0;
    const unsigned char* base = static_cast<const unsigned char*>(points);
    for (size_t i = 0; i < n; ++i) {
        float fx, fy, fz;
        std::memcpy(&fx, base + i * stride + 0, 4);
        std::memcpy(&fy, base + i * stride + 4, 4);
        std::memcpy(&fz, base + i * stride + 8, 4);
        if (!finite3(fx, fy, fz)) { ++drop; continue; }  // contract Q3: non-finite points are skipped
        const double x = fx, y = fy, z = fz;
        const double range = point_range(x, y, z);                        // :259
        if (range < p->min_range || range > p->max_range) { ++drop; continue; }  // :261
        const double azimuth = point_azimuth(x, y);                       // :266
        const double altitude = point_altitude(x, y, z);                  // :267
        if (altitude < p->min_angle || altitude > p->max_angle) { ++drop; continue; }  // :269
        double rowf, colf;
        if (p->variant == GS_VARIANT_CLOUD2RANGE) {
            rowf = ((p->max_angle - altitude) / (p->max_angle - p->min_angle)) * H;  // cloud2range_node.cpp:436
            colf = azimuth / d.d_az;                                                 // :438
        } else {
            rowf = (H - 1) - ((altitude - half_d_alt - p->min_angle) / d.d_alt);     // alfa_gs.cpp:275
            colf = (azimuth + half_d_az) / d.d_az;                                   // :277
        }
        if (!(colf >= 0.0)) { ++drop; continue; }  // contract Q3 (y=-0.0, x<0 gives a negative column)
        const int row = static_cast<int>(rowf);
        int col = static_cast<int>(colf);
        if (near_edge) {
            const double fr = rowf - std::floor(rowf), fc = colf - std::floor(colf);
            const double cell_alt = (p->variant == GS_VARIANT_CLOUD2RANGE) ? (p->max_angle - p->min_angle) / H : d.d_alt;
            if (std::min(fr, 1 - fr) * cell_alt < 1e-5 || std::min(fc, 1 - fc) * d.d_az < 1e-5) ++ne;
        }
        if (row < 0 || row >= H) { ++drop; continue; }  // Q4: cloud2range asserts (:442); contract: drop
        size_t flat;
        if (p->variant == GS_VARIANT_CLOUD2RANGE) {
            if (col >= W) { ++drop; continue; }  // cloud2range_node.cpp:443 asserts; contract: drop
            flat = static_cast<size_t>(row) * W + col;
        } else if (p->col_policy == GS_COL_WRAP) {
            col %= W;
            flat = static_cast<size_t>(row) * W + col;
        } else {
            // Q2: col may equal W; at<ushort>(row, W) of a continuous Mat is flat index row*W+W
            flat = static_cast<size_t>(row) * W + col;
            if (flat >= P) { ++drop; continue; }  // out-of-bounds store in the reference; contract: drop
        }
        const double range_norm = (range - p->min_range) / (p->max_range - p->min_range);  // :288
        const uint16_t enc = static_cast<uint16_t>(range_norm * (std::numeric_limits<uint16_t>::max() - 1) + 1);  // :291
        if (p->pixel_policy == GS_PIXEL_NEAREST) {
            if (img[flat] == 0 || enc < img[flat]) img[flat] = enc;
        } else {
            img[flat] = enc;  // Q1: later points overwrite earlier ones
        }
    }
    if (near_edge) *near_edge = ne;
    if (dropped) *dropped = drop;
    return GS_OK;
}

// ---- A1 : Alfa_GS::CreateAngleImg (alfa_gs.cpp:335-386 == cloud2range_node.cpp:517-568) -------
int gso_create_angle_image(const gs_params* p, const uint16_t* range_image, uint16_t* angle_image,
                           uint64_t* near_edge) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    std::memset(angle_image, 0, static_cast<size_t>(H) * W * sizeof(uint16_t));  // :338
    int prev_col = 1;                                                            // :340
    float prev_r = 0, prev_angle_y = 0, angle_y = 0, delta_z = 0, delta_x = 0, alfa = 0;  // :342-343
    const float vertical_amplitude = static_cast<float>(p->max_angle - p->min_angle);     // :344
    uint64_t ne = 0;
    for (int col = 0; col < W; ++col) {            // :347
        for (int row = H - 1; row >= 0; --row) {   // :349
            const uint16_t enc = range_image[static_cast<size_t>(row) * W + col];
            if (enc == 0) continue;                // :355
            angle_y = static_cast<float>(row * d.d_alt - vertical_amplitude / 2);  // :357 (float/int, then double math)
            const double range_norm = static_cast<double>(enc - 1) / (std::numeric_limits<uint16_t>::max() - 1);  // :359
            const double range = range_norm * (p->max_range - p->min_range) + p->min_range;                        // :360
            if (col == prev_col) {  // :362
                // :366-367  float*float product minus double*float product, |.| in double, stored as float
                delta_z = static_cast<float>(std::fabs(static_cast<double>(prev_r * sinf(prev_angle_y)) - range * static_cast<double>(sinf(angle_y))));
                delta_x = static_cast<float>(std::fabs(static_cast<double>(prev_r * cosf(prev_angle_y)) - range * static_cast<double>(cosf(angle_y))));
                alfa = atan2f(delta_z, delta_x);  // :368
                const float scaled = alfa * 100;  // :370 float product
                angle_image[static_cast<size_t>(row) * W + col] = static_cast<uint16_t>(scaled);
                if (near_edge) {
                    const double fr = scaled - std::floor(static_cast<double>(scaled));
                    if (std::min(fr, 1 - fr) < 1e-3) ++ne;  // 1e-5 rad of a centi-radian edge
                }
            }
            prev_col = col;                        // :377
            prev_r = static_cast<float>(range);    // :379
            prev_angle_y = angle_y;                // :380
        }
    }
    if (near_edge) *near_edge = ne;
    return GS_OK;
}

// ---- S2 : Alfa_GS::SavitskyGolaySmoothing (alfa_gs.cpp:705-713): cv::filter2D(image, out, -1, ws x 1 float
// kernel, anchor centre, delta 0, BORDER_REFLECT101) on a 16U image.  OpenCV's direct filter engine:
// float32 accumulation over the non-zero taps in kernel order, then cvRound (half-to-even), saturate to u16.
int gso_sg_smooth(const gs_params* p, const uint16_t* in, int ws, uint16_t* out) {
    std::vector<float> taps;
    if (ws % 2 == 0 || !sg_taps(ws, taps)) return GS_EINVAL;
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W, half = ws / 2;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            float acc = 0.0f;
            for (int k = 0; k < ws; ++k) {
                const int rr = reflect101(r + k - half, H);
                acc = acc + taps[k] * static_cast<float>(in[static_cast<size_t>(rr) * W + c]);
            }
            long v = std::lrintf(acc);  // round-half-even under the default rounding mode
            if (v < 0) v = 0;
            if (v > 65535) v = 65535;
            out[static_cast<size_t>(r) * W + c] = static_cast<uint16_t>(v);
        }
    return GS_OK;
}

// ---- N2 : typed RepairGaps (alfa_gs.cpp:300-333 / cloud2range_node.cpp:482-515).  The literal code reads
// the 16U image through at<float> (Q22, out of bounds); this is the typed reading: depths are the
// encoded u16 values taken as float, the result is written back truncated to u16.
int gso_repair_gaps(const gs_params* p, const uint16_t* in, int step, float depth_threshold, uint16_t* out) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    std::memcpy(out, in, static_cast<size_t>(H) * W * sizeof(uint16_t));  // clone (:302)
    for (int c = 0; c < W; ++c)
        for (int r = 0; r < H; ++r) {  // in place, top to bottom: repaired pixels feed later ones (:304-305)
            const float curr = out[static_cast<size_t>(r) * W + c];
            if (!(curr < 0.001f)) continue;  // :307
            int counter = 0;
            float sum = 0.0f;
            for (int i = 1; i < step; ++i) {
                if (r - i < 0) continue;
                for (int j = 1; j < step; ++j) {
                    if (r + j > H - 1) continue;
                    const float prev = out[static_cast<size_t>(r - i) * W + c];
                    const float next = out[static_cast<size_t>(r + j) * W + c];
                    if (prev > 0.001f && next > 0.001f && std::fabs(prev - next) < depth_threshold) {  // :320
                        sum += prev + next;
                        counter += 2;
                    }
                }
            }
            if (counter > 0) out[static_cast<size_t>(r) * W + c] = static_cast<uint16_t>(sum / counter);  // :326-328
        }
    return GS_OK;
}

// ---- G1 : Alfa_GS::EraseBFS (alfa_gs.cpp:715-766) == Cloud2RangeNode::EraseGroundBFS (cloud2range_node.cpp:701-744)
// labels_out / dilated_out (u8, may be NULL) expose the Labeler image and its dilation.
int gso_erase_bfs(const gs_params* p, const uint16_t* range, const uint16_t* angle, double ground_thr,
                  double start_thr, int kernel_size, uint16_t* res, uint8_t* labels_out, uint8_t* dilated_out) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    const size_t P = static_cast<size_t>(H) * W;
    const int ks = std::max(kernel_size - 2, 3);  // :747
    if (ks % 2 == 0) return GS_EINVAL;            // :631-633 throws
    LabelImage labeler(H, W);                     // :717
    std::memset(res, 0, P * sizeof(uint16_t));    // :719
    for (int c = 0; c < W; ++c) {                 // :720
        int r = H - 1;
        while (r > 0 && range[static_cast<size_t>(r) * W + c] < 1) --r;  // :722-725
        if (labeler.at(r, c) > 0) continue;                              // :727-731
        if (angle[static_cast<size_t>(r) * W + c] > start_thr * 100) continue;  // :733 (double compare)
        labeler.label_one_component(1, r, c, static_cast<float>(ground_thr), range, angle);  // :738 (double -> float param)
    }
    const std::vector<int> off = dilate_offsets(kernel_size);  // :747-750
    const int erase_ground = (p->variant == GS_VARIANT_CLOUD2RANGE) ? 1 : p->erase_ground;  // EraseGroundBFS has no switch
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            int dil = 0;
            for (size_t k = 0; k < off.size(); ++k) {
                const int rr = r + off[k];
                if (rr < 0 || rr >= H) continue;  // outside the image: ignored by cv::dilate
                dil |= labeler.at(rr, c) != 0;
            }
            const size_t i = static_cast<size_t>(r) * W + c;
            if (labels_out) labels_out[i] = static_cast<uint8_t>(labeler.at(r, c) != 0);
            if (dilated_out) dilated_out[i] = static_cast<uint8_t>(dil);
            if (erase_ground) {
                if (!dil) res[i] = range[i];  // :755-756
            } else {
                if (dil) res[i] = range[i];   // :760-761
            }
        }
    return GS_OK;
}

// Label::LabelOneComponent through the Label interface (Labeler/label.h:41) on a caller-held u8 label image.
int gso_label_one_component(const gs_params* p, int label, int row, int col, float thr, const uint16_t* range,
                            const uint16_t* angle, uint8_t* label_image) {
    const Dims d = dims_of(*p);
    if (row < 0 || row >= d.H || col < 0 || col >= d.W) return GS_EINVAL;
    if ((label & 0xff) == 0) return GS_EINVAL;  // label 0 would never terminate in the reference
    LabelImage li(d.H, d.W);
    const size_t P = static_cast<size_t>(d.H) * d.W;
    for (size_t i = 0; i < P; ++i) li.lab[i] = label_image[i];
    li.label_one_component(static_cast<uint8_t>(label), row, col, thr, range, angle);  // uint8_t label (label.h:41)
    for (size_t i = 0; i < P; ++i) label_image[i] = static_cast<uint8_t>(li.lab[i]);
    return GS_OK;
}

// The Labeler driven as an all-components CCL (SURVEY.md A.4 "Component labels"): raster scan over the
// unlabelled NON-EMPTY pixels of `image`, one LabelOneComponent per new component with a wide label.
// labels_out = smallest flat index of the component over its non-empty pixels, -1 for empty pixels.
int gso_label_components(const gs_params* p, const uint16_t* image, const uint16_t* angle, double beta,
                         int32_t* labels_out, int32_t* n_components) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    const size_t P = static_cast<size_t>(H) * W;
    LabelImage li(H, W);
    int32_t next = 0;
    std::vector<int32_t> first_index;  // label -> raster-first (= smallest) flat index of a non-empty pixel
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const size_t i = static_cast<size_t>(r) * W + c;
            if (image[i] == 0 || li.lab[i] > 0) continue;
            ++next;
            first_index.push_back(static_cast<int32_t>(i));
            li.label_one_component(next, r, c, static_cast<float>(beta), image, angle);
        }
    for (size_t i = 0; i < P; ++i) labels_out[i] = (image[i] != 0 && li.lab[i] > 0) ? first_index[li.lab[i] - 1] : -1;
    if (n_components) *n_components = next;
    return GS_OK;
}

// ---- B1 / B1' : Alfa_GS::CameraCb (alfa_gs.cpp:997-1074) / Range2CloudNode::CameraCb (range2cloud_node.cpp:36-99)
// xyz4: records {x, y, z, 1.0f} (the PCL point padding), row-major order of the non-zero pixels.
int gso_camera_cb(const gs_params* p, const uint16_t* image, float* xyz4, uint32_t* n_out) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    // cinfo_.K[0..5] (alfa_gs.cpp:58-63 -> :1010-1015)
    const double min_angle = p->min_angle, min_range = p->min_range, max_range = p->max_range;
    const double d_azimuth = d.d_az, d_altitude = d.d_alt;
    uint32_t n = 0;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const uint16_t enc = image[static_cast<size_t>(r) * W + c];
            if (enc == 0) continue;  // :1030
            const double range_norm = static_cast<double>(enc - 1) / (std::numeric_limits<uint16_t>::max() - 1);  // :1042
            const double altitude = (H - r) * d_altitude + min_angle;  // :1050 (Q15: one beam step high)
            const double azimuth = c * d_azimuth;                      // :1052
            float x, y, z;
            if (p->variant == GS_VARIANT_CLOUD2RANGE) {
                const double range = range_norm * (max_range - min_range) + min_range;  // range2cloud_node.cpp:76
                x = static_cast<float>(std::cos(altitude) * std::cos(azimuth) * range);  // :85-87
                y = static_cast<float>(std::cos(altitude) * std::sin(azimuth) * range);
                z = static_cast<float>(std::sin(altitude) * range);
            } else {
                const float range = static_cast<float>(range_norm * (max_range - min_range) + min_range);  // alfa_gs.cpp:1027,1045
                x = static_cast<float>(std::cos(altitude) * std::cos(azimuth) * range);  // :1055-1057
                y = static_cast<float>(std::cos(altitude) * std::sin(azimuth) * range);
                z = static_cast<float>(std::sin(altitude) * range);
            }
            xyz4[4 * static_cast<size_t>(n) + 0] = x;
            xyz4[4 * static_cast<size_t>(n) + 1] = y;
            xyz4[4 * static_cast<size_t>(n) + 2] = z;
            xyz4[4 * static_cast<size_t>(n) + 3] = 1.0f;
            ++n;
        }
    *n_out = n;
    return GS_OK;
}

// ---- O1 / O1' : Alfa_GS::process_pointcloud hw==0 (alfa_gs.cpp:175-244) and Cloud2RangeNode::CloudCb
// (cloud2range_node.cpp:381-480) followed by Range2CloudNode::CameraCb.  MovingAverageSmoothing
// (alfa_gs.cpp:212) is not executed: its result is discarded (:220) and it indexes out of bounds (Q6).
// Any output may be NULL.  stage_ms (may be NULL): 4 doubles RI / AI / GS / back-projection, the
// reference's own timer boundaries (alfa_gs.cpp:195-229).
int gso_process_frame(const gs_params* p, const void* points, size_t n, size_t stride, uint16_t* range_out,
                      uint16_t* angle_out, uint16_t* no_ground_out, float* xyz4_out, uint32_t* n_out,
                      double* stage_ms) {
    using clk = std::chrono::high_resolution_clock;
    const Dims d = dims_of(*p);
    const size_t P = static_cast<size_t>(d.H) * d.W;
    std::vector<uint16_t> range(P), angle(P), work(P), res(P);
    const auto t0 = clk::now();
    int rc = gso_create_range_image(p, points, n, stride, range.data(), nullptr, nullptr);  // :196
    if (rc) return rc;
    const auto t1 = clk::now();
    const uint16_t* angle_src = range.data();
    if (p->repair_gaps) {  // cloud2range_node.cpp:468 (typed)
        gso_repair_gaps(p, range.data(), 5, 1.0f, work.data());
        angle_src = work.data();
    }
    gso_create_angle_image(p, angle_src, angle.data(), nullptr);  // :211
    if (p->smoothing == GS_SMOOTH_SG) {                           // cloud2range_node.cpp:470
        rc = gso_sg_smooth(p, angle.data(), p->window_size, work.data());
        if (rc) return rc;
        angle.swap(work);
    }
    const auto t2 = clk::now();
    rc = gso_erase_bfs(p, range.data(), angle.data(), p->ground_angle_threshold, p->start_angle_threshold,
                       p->window_size, res.data(), nullptr, nullptr);  // :220
    if (rc) return rc;
    const auto t3 = clk::now();
    uint32_t cnt = 0;
    std::vector<float> tmp;
    float* xyz = xyz4_out;
    if (!xyz) { tmp.resize(4 * P); xyz = tmp.data(); }
    gso_camera_cb(p, res.data(), xyz, &cnt);  // :239
    const auto t4 = clk::now();
    if (range_out) std::memcpy(range_out, range.data(), P * 2);
    if (angle_out) std::memcpy(angle_out, angle.data(), P * 2);
    if (no_ground_out) std::memcpy(no_ground_out, res.data(), P * 2);
    if (n_out) *n_out = cnt;
    if (stage_ms) {
        stage_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        stage_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
        stage_ms[2] = std::chrono::duration<double, std::milli>(t3 - t2).count();
        stage_ms[3] = std::chrono::duration<double, std::milli>(t4 - t3).count();
    }
    return GS_OK;
}

// Batch driver for the CPU baseline: frames [0, n_frames) (16-byte XYZI records, offsets in points) processed
// as independent streams on `n_threads` host threads (1 = the reference's execution model: one
// ros::spin thread, alfa_node.cpp:387).  Outputs may be NULL.  Returns wall seconds in *seconds_out and the
// summed per-stage milliseconds in stage_ms_sum[4].
int gso_process_batch(const gs_params* p, const float* points_xyzi, const int64_t* offsets, int n_frames,
                      int n_threads, uint16_t* no_ground_out, float* xyz4_out, uint32_t* counts_out,
                      double* seconds_out, double* stage_ms_sum) {
    const Dims d = dims_of(*p);
    const size_t P = static_cast<size_t>(d.H) * d.W;
    if (n_threads < 1) n_threads = 1;
    std::atomic<int> next(0);
    std::atomic<int> err(0);
    std::vector<std::vector<double>> ms(n_threads, std::vector<double>(4, 0.0));
    auto worker = [&](int tid) {
        for (;;) {
            const int f = next.fetch_add(1);
            if (f >= n_frames) break;
            double st[4];
            uint32_t cnt = 0;
            const int rc = gso_process_frame(p, points_xyzi + 4 * offsets[f], static_cast<size_t>(offsets[f + 1] - offsets[f]), 16,
                                             nullptr, nullptr, no_ground_out ? no_ground_out + f * P : nullptr,
                                             xyz4_out ? xyz4_out + 4 * f * P : nullptr, &cnt, st);
            if (rc) err = rc;
            if (counts_out) counts_out[f] = cnt;
            for (int k = 0; k < 4; ++k) ms[tid][k] += st[k];
        }
    };
    const auto t0 = std::chrono::high_resolution_clock::now();
    if (n_threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    if (stage_ms_sum)
        for (int k = 0; k < 4; ++k) {
            stage_ms_sum[k] = 0;
            for (int t = 0; t < n_threads; ++t) stage_ms_sum[k] += ms[t][k];
        }
    return err.load();
}

// lib/utils.h helpers as restated above, for the cross-check against the verbatim build in oracle/_ref.
double gso_point_range(double x, double y, double z) { return point_range(x, y, z); }
double gso_point_azimuth(double x, double y) { return point_azimuth(x, y); }
double gso_point_altitude(double x, double y, double z) { return point_altitude(x, y, z); }
double gso_rad_deg(double deg) { return rad_from_deg(deg); }

// libm entry points as seen by this oracle build, for checking device-side clones bit by bit.
float gso_atan2f(float y, float x) { return atan2f(y, x); }
float gso_sinf(float x) { return sinf(x); }
float gso_cosf(float x) { return cosf(x); }

// ---- N3 : Alfa_GS::CreateLabeledRangeImage (alfa_gs.cpp:945-995) ------------------------------
// points: 32-byte pcl::PointXYZRGB records (x y z pad | b g r a | pad).  gnd_flag is raised before the gates (:961-965)
// and lowered only at the end of a kept point's iteration (:992): it sticks across skipped points (kept literally).
// Projection = CreateRangeImage's (the same lines, :968-985); the flat-index alias of column W and non-finite points
// follow the same contract (Q2, Q3).
int gso_create_labeled_range_image(const gs_params* p, const void* points, size_t n, uint16_t* img) {
    const Dims d = dims_of(*p);
    const int H = d.H, W = d.W;
    const size_t P = static_cast<size_t>(H) * W;
    for (size_t i = 0; i < P; ++i) img[i] = 1;  // Mat(n_beams_, n_cols_, CV_16UC1, 1)  :947
    const double half_d_alt = d.d_alt / 2, half_d_az = d.d_az / 2;  // :953-954
    int gnd_flag = 0;
    const unsigned char* base = static_cast<const unsigned char*>(points);
    for (size_t i = 0; i < n; ++i) {
        float fx, fy, fz;
        std::memcpy(&fx, base + i * 32 + 0, 4);
        std::memcpy(&fy, base + i * 32 + 4, 4);
        std::memcpy(&fz, base + i * 32 + 8, 4);
        const unsigned b = base[i * 32 + 16], g = base[i * 32 + 17], r = base[i * 32 + 18];
        if ((r == 75 && g == 0 && b == 75) || (r == 255 && g == 0 && b == 255) || (r == 255 && g == 150 && b == 255)) gnd_flag = 1;  // :961-965
        if (!finite3(fx, fy, fz)) continue;  // contract Q3
        const double x = fx, y = fy, z = fz;
        const double range = point_range(x, y, z);
        if (range < p->min_range || range > p->max_range) continue;  // :967-971
        const double azimuth = point_azimuth(x, y);
        const double altitude = point_altitude(x, y, z);
        if (altitude < p->min_angle || altitude > p->max_angle) continue;  // :975-979
        double rowf, colf;
        if (p->variant == GS_VARIANT_CLOUD2RANGE) {
            rowf = ((p->max_angle - altitude) / (p->max_angle - p->min_angle)) * H;
            colf = azimuth / d.d_az;
        } else {
            rowf = (H - 1) - ((altitude - half_d_alt - p->min_angle) / d.d_alt);  // :981
            colf = (azimuth + half_d_az) / d.d_az;                                // :982
        }
        if (!(colf >= 0.0)) continue;
        const int row = static_cast<int>(rowf);
        int col = static_cast<int>(colf);
        if (row < 0 || row >= H) continue;
        size_t flat;
        if (p->variant == GS_VARIANT_CLOUD2RANGE) {
            if (col >= W) continue;
            flat = static_cast<size_t>(row) * W + col;
        } else if (p->col_policy == GS_COL_WRAP) {
            col %= W;
            flat = static_cast<size_t>(row) * W + col;
        } else {
            flat = static_cast<size_t>(row) * W + col;
            if (flat >= P) continue;
        }
        if (gnd_flag == 1) img[flat] = 0;  // :988-991
        gnd_flag = 0;                      // :992
    }
    return GS_OK;
}

// ---- N3 : Alfa_GS::CheckNumberOfDetectedRIdGnd (alfa_gs.cpp:831-889), per-frame part ------------
int gso_evaluate(const gs_params* p, const uint16_t* og, const uint16_t* seg, const uint16_t* lab, gs_eval* out) {
    const Dims d = dims_of(*p);
    int points_cntr = 0, seg_gnd_cntr = 0, labeled_gnd_cntr = 0, labeled_obst_cntr = 0, true_detected_obst_cntr = 0;  // :833
    float true_positive = 0, false_positive = 0, true_negative = 0, false_negative = 0;                              // :834
    for (int col = 0; col < d.W; ++col) {      // :838
        for (int row = 0; row < d.H; ++row) {  // :840
            const size_t i = static_cast<size_t>(row) * d.W + col;
            points_cntr++;
            if (lab[i] == 0) labeled_gnd_cntr++;
            else if (og[i] != 0) labeled_obst_cntr++;
            if (seg[i] == 0 && og[i] != 0) seg_gnd_cntr++;
            if (seg[i] == lab[i] && og[i] != 0) true_positive++;
            if (seg[i] == 0 && lab[i] != 0 && og[i] != 0) false_positive++;
            if (seg[i] != 0 && lab[i] != 0) true_negative++;
            if (seg[i] != 0 && lab[i] == 0) false_negative++;
            if (seg[i] != 0 && lab[i] != 0) true_detected_obst_cntr++;
        }
    }
    out->points = points_cntr;
    out->labeled_gnd = labeled_gnd_cntr;
    out->labeled_obst = labeled_obst_cntr;
    out->seg_gnd = seg_gnd_cntr;
    out->true_positive = static_cast<uint64_t>(true_positive);
    out->false_positive = static_cast<uint64_t>(false_positive);
    out->true_negative = static_cast<uint64_t>(true_negative);
    out->false_negative = static_cast<uint64_t>(false_negative);
    out->true_detected_obst = true_detected_obst_cntr;
    out->iou_g = ((true_positive) / (true_positive + false_positive + false_negative)) * 100;  // :885
    out->recall_g = (true_positive / (true_positive + false_negative)) * 100;                   // :886
    out->recall_mo = (true_negative / (true_negative + false_positive)) * 100;                  // :887
    return GS_OK;
}

}  // extern "C"
