// gs_internal.h — declarations shared by gs_kernels.cu (device code + launchers) and gs_capi.cu (context, C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gs_b200.h"

namespace gs {

// Everything a kernel needs to know about the sensor, passed by value.
struct KParams {
    int H, W, P;
    int variant, pixel_policy, col_policy;
    double min_range, max_range, span;          // span = max_range - min_range
    double min_angle, max_angle, ang_span;      // ang_span = max_angle - min_angle
    double d_alt, d_az, half_d_alt, half_d_az;  // reference src/alfa_gs.cpp:48-50,249-250
    int k_ground, s_start;                      // integer cut-offs
    int n_dil, dil0, dil1, erase_ground;        // effective dilation taps (A.4)
};

// Host-computed lookup tables (device pointers).  Computing them with the host's libm is what makes
// the angle image and the back-projection bit-exact.
struct Luts {
    const float* sin_y;   // [H] sinf(angle_y(row))   alfa_gs.cpp:357,366
    const float* cos_y;   // [H] cosf(angle_y(row))   alfa_gs.cpp:357,367
    const double* cos_alt;  // [H] cos((H-r)*d_alt+min_angle)   alfa_gs.cpp:1050,1055
    const double* sin_alt;  // [H]
    const double* cos_az;   // [W] cos(c*d_az)                  alfa_gs.cpp:1052
    const double* sin_az;   // [W]
    const double* sin_y_d;  // [H] (double)sinf(angle_y(row)): the value the reference widens in alfa_gs.cpp:366
    const double* cos_y_d;  // [H]
};

struct SgTaps {
    float t[11];
    int ws;
};

struct Counters {
    unsigned long long near_edge_project, near_edge_angle, dropped, slow_path, components, fast_mismatch, remote;
    unsigned long long phase[20];  // fused kernel: clock64() cycles per phase, summed over CTAs (statistics runs only)
};

// Device scratch for one wave ("chunk") of frames; every array holds `frames` slabs.
struct Scratch {
    int frames = 0;
    size_t P = 0;
    int W = 0, H = 0;
    unsigned long long* keys = nullptr;  // [frames][P] packed (index, range) pixel owners, kept zeroed between waves
    uint16_t* range = nullptr;           // [frames][P]
    uint16_t* angle = nullptr;           // [frames][P]
    uint16_t* angle2 = nullptr;          // [frames][P] smoothed / repaired
    int* parent = nullptr;               // [frames][P] union-find forest, root = smallest flat index, bit 30 = ground flag
    uint8_t* label = nullptr;            // [frames][P] the Labeler's image L
    int* seed_row = nullptr;             // [frames][W]
    uint8_t* eseed = nullptr;            // [frames][W] accepted seeds that are empty pixels
    unsigned* rowcount = nullptr;        // [frames][H]
    long long* offsets = nullptr;        // [frames+1] point offsets of the wave (device copy)
    uint16_t* res = nullptr;             // [frames][P] (only when the caller gives no output buffer)
    unsigned* counts = nullptr;          // [frames]
};

struct Launch {
    cudaStream_t stream;
    unsigned long long* n_launches;  // host counter
};

// ---- launchers (gs_kernels.cu).  `frames` slabs are processed per call; arrays are [frames][P] unless noted. ----
void launch_unpack_points(const Launch& L, const void* data, size_t n, unsigned step, unsigned ox, unsigned oy, unsigned oz, int oi, int big_endian,
                          float* xyzi_out);
void launch_pack_xyzrgb(const Launch& L, const float* xyz4, unsigned n, void* out);
void launch_project(const Launch& L, const KParams& kp, const void* points, size_t stride, const long long* d_offsets,
                    int frames, long long max_points, unsigned long long* keys, Counters* cnt, bool stats);
void launch_finalize(const Launch& L, const KParams& kp, int frames, unsigned long long* keys, uint16_t* range);
void launch_angle(const Launch& L, const KParams& kp, const Luts& lut, int frames, const uint16_t* range, uint16_t* angle,
                  Counters* cnt, bool stats);
void launch_sg(const Launch& L, const KParams& kp, const SgTaps& taps, int frames, const uint16_t* in, uint16_t* out);
void launch_repair(const Launch& L, const KParams& kp, int frames, const uint16_t* in, uint16_t* out, int step, float thr);
void launch_ccl(const Launch& L, const KParams& kp, int frames, const uint16_t* range, const uint16_t* angle,
                const uint8_t* blocked, int k_edge, int* parent);
void launch_ground(const Launch& L, const KParams& kp, int frames, const uint16_t* range, const uint16_t* angle,
                   int* parent, int* seed_row, uint8_t* eseed, uint8_t* label);
void launch_one_component(const Launch& L, const KParams& kp, const uint16_t* range, const uint16_t* angle, int k_edge,
                          int* parent, int row, int col, int label, uint8_t* label_image);
void launch_components(const Launch& L, const KParams& kp, int frames, const uint16_t* image, int* parent,
                       int32_t* labels, Counters* cnt);
// the Labeler as an all-components CCL on `frames` images (shared-memory tiles + seam unions + flatten); labels: [frames][P] int32,
// n_components: [frames] or NULL, cnt->components accumulates the total when cnt is given
void launch_components_batch(const Launch& L, const KParams& kp, int frames, const uint16_t* image, const uint16_t* angle, int k_edge,
                             int32_t* labels, unsigned* n_components, Counters* cnt);
// N4: AlfaNode's FPGA wire formats and the coloured angle image
void launch_fpga_store(const Launch& L, const void* points, size_t n, size_t stride, unsigned long long* words);
void launch_fpga_read_range(const Launch& L, const unsigned long long* words, unsigned n_pixels, uint16_t* image64x1800);
void launch_fpga_read_angle(const Launch& L, const unsigned long long* words, unsigned n_pixels, unsigned cols, uint16_t* image);
void launch_colored_angle(const Launch& L, const uint16_t* angle, unsigned n_pixels, unsigned char* rgb);
void launch_erase(const Launch& L, const KParams& kp, int frames, const uint16_t* range, const uint8_t* label, uint16_t* res,
                  unsigned* rowcount);
void launch_backproject(const Launch& L, const KParams& kp, const Luts& lut, int frames, const uint16_t* image,
                        const unsigned* rowcount, float* xyz4, unsigned* counts, bool range_as_double);
void launch_rowcount(const Launch& L, const KParams& kp, int frames, const uint16_t* image, unsigned* rowcount);
// host batch path: packs the per-frame cloud slabs [frames][P] back to back (prefix: frames + 1 entries of device scratch)
void launch_pack_clouds(const Launch& L, const float4* slabs, const unsigned* counts, int frames, size_t P, unsigned long long* prefix,
                        float4* packed);
// N3 evaluation: Alfa_GS::CreateLabeledRangeImage (32-byte PointXYZRGB records) and the counters of CheckNumberOfDetectedRIdGnd
size_t label_image_scratch_bytes(size_t n);
void launch_label_image(const Launch& L, const KParams& kp, const void* points32, size_t n, uint8_t* flags, int* flat, uint8_t* summary,
                        uint16_t* labeled);
void launch_evaluate(const Launch& L, const KParams& kp, const uint16_t* og, const uint16_t* seg, const uint16_t* lab, unsigned long long* out9);


enum { FUSED_MODE_HW = 0, FUSED_MODE_HYBRID = 1, FUSED_MODE_SOFT = 2 };
// ---- fused single-kernel pipeline (gs_fused.cu): one thread-block cluster per frame, images resident in (distributed)
// shared memory from the projection to the back-projection.
struct FusedState {
    bool disabled = false;   // GS_OPT_FORCE_STAGED
    bool configured = false; // the configuration fits the fused kernel
    int cluster = 0;         // CTAs per frame
    int threads = 1024;      // threads per CTA
    const void* kernel = nullptr;  // the k_frame instantiation chosen for this configuration
    int cols_per_cta = 0;
    size_t smem_bytes = 0;
    int n_sms = 0;
    int max_clusters = 0;    // co-resident clusters (persistent grid size)
    int mode = 0;            // FUSED_MODE_*
    const void* kernel_soft = nullptr;  // the software-cluster instantiation of the same shape
    int soft_clusters = 0;   // software clusters launched beside (hybrid) or instead of (soft) the cluster kernel
    int lone_cluster = 0;    // CTAs per frame when a call brings ONE frame (the widest cluster that fits), 0 = none
    const void* lone_kernel = nullptr;
    bool no_lone = false;    // GS_OPT_NO_LONE_PLAN: single frames take the batch plan's cluster (tests)
    void* d_mail = nullptr;  // frame queue + software-cluster mailboxes
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    void* d_const = nullptr; // device copy of the per-configuration constants
    unsigned* d_rowbase = nullptr;
    size_t rowbase_cap = 0;
    KParams kp;
    Luts lut;
    SgTaps sg;
    bool smooth = false, repair = false;
};
// return 0 or a cudaError_t
int fused_configure(FusedState& fs, const KParams& kp, const Luts& lut, const SgTaps& sg, bool smooth, bool repair, int device);
bool fused_supported(const FusedState& fs);
long long fused_max_points();  // frames with more points than this run on the staged kernels
int fused_run(FusedState& fs, const Launch& L, const float4* points, const long long* d_offsets, int frames, long long max_points,
              uint16_t* range_out, uint16_t* angle_out, uint16_t* res_out, float4* xyz_out, unsigned* counts_out, Counters* cnt,
              int stats);
void fused_release(FusedState& fs);

}  // namespace gs
