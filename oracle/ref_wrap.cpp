// ref_wrap.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN sources, compiled verbatim from /root/reference:
//   Labeler/label.{h,cpp}  (class cloud2range::Label, the BFS)        and
//   lib/utils.h            (PointRange / PointAzimuth / PointAltitude / Deg_Rad / Rad_Deg).
// Output goes to oracle/_ref/libgs_ref.so (git-ignored); used by tests/ to validate the oracle's
// restatement of those two files.  Nothing of the reference is copied into this repository.
#include <cstdint>
#include <cstring>

#include "Labeler/label.h"  // -I/root/reference
#include "utils.h"          // -I/root/reference/lib

using cloud2range::Label;

namespace {
cv::Mat mat_from_u16(const uint16_t* src, int rows, int cols) {
    cv::Mat m = cv::Mat::zeros(rows, cols, CV_16UC1);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) m.at<uint16_t>(r, c) = src[static_cast<size_t>(r) * cols + c];
    return m;
}
}  // namespace

extern "C" {

double gsref_point_range(double x, double y, double z) { return cloud2range::PointRange(x, y, z); }
double gsref_point_azimuth(double x, double y) { return cloud2range::PointAzimuth(x, y); }
double gsref_point_altitude(double x, double y, double z) { return cloud2range::PointAltitude(x, y, z); }
double gsref_deg_rad(double rad) { return cloud2range::Deg_Rad(rad); }
double gsref_rad_deg(double deg) { return cloud2range::Rad_Deg(deg); }

int gsref_wrap_cols(int rows, int cols, int col) {
    Label l(rows, cols);
    return l.WrapCols(static_cast<int16_t>(col));
}

// One Label::LabelOneComponent call on a fresh Label pre-loaded with `label_image` (u8, in/out).
void gsref_label_one_component(int rows, int cols, int label, int row, int col, float thr, const uint16_t* range,
                               const uint16_t* angle, uint8_t* label_image) {
    Label l(rows, cols);
    cv::Mat img = l.GetLabelImage();  // shares storage with the Label's image, like a cv::Mat header copy
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) img.at<uint16_t>(r, c) = label_image[static_cast<size_t>(r) * cols + c];
    l.LabelOneComponent(static_cast<uint8_t>(label), row, col, thr, mat_from_u16(range, rows, cols),
                        mat_from_u16(angle, rows, cols));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) label_image[static_cast<size_t>(r) * cols + c] = l.CheckLabelAt(r, c);
}

// The seed loop of Alfa_GS::EraseBFS (src/alfa_gs.cpp:720-739, restated) driving the VERBATIM Label class.
void gsref_ground_labels(int rows, int cols, double ground_thr, double start_thr, const uint16_t* range,
                         const uint16_t* angle, uint8_t* label_image) {
    Label l(rows, cols);
    const cv::Mat range_m = mat_from_u16(range, rows, cols), angle_m = mat_from_u16(angle, rows, cols);
    for (int c = 0; c < cols; ++c) {
        int r = rows - 1;
        while (r > 0 && range[static_cast<size_t>(r) * cols + c] < 1) --r;
        if (l.CheckLabelAt(r, c) > 0) continue;
        if (angle[static_cast<size_t>(r) * cols + c] > start_thr * 100) continue;
        l.LabelOneComponent(1, r, c, ground_thr, range_m, angle_m);
    }
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) label_image[static_cast<size_t>(r) * cols + c] = l.CheckLabelAt(r, c);
}

}  // extern "C"
