// gs_capi.cu — context and C ABI (include/gs_b200.h) over the kernels of gs_kernels.cu / gs_fused.cu.
//
// Nothing here computes on the CPU except parameter derivation and the per-row / per-column lookup
// tables (the reference's constructor work, src/alfa_gs.cpp:16-63).  There is no CPU fallback: every
// entry point that produces image or cloud data launches CUDA kernels and fails with GS_ECUDA when no
// device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "gs_internal.h"

using namespace gs;

// ------------------------------------------------------------------------------------------------
// parameter derivation (host; reference constructor arithmetic)
// ------------------------------------------------------------------------------------------------
namespace {

// Label::LabelOneComponent compares an integer-valued double with `float thr * 100` evaluated in float
// (Labeler/label.cpp:99, label.h:41):  diff < T  <=>  diff <= ceil(T) - 1.
int edge_cutoff(double thr) {
    const float t = static_cast<float>(thr);
    const float T = t * 100;
    return static_cast<int>(std::ceil(static_cast<double>(T))) - 1;
}
// EraseBFS seed gate (alfa_gs.cpp:733): rejected when angle > start * 100 (double).
int seed_cutoff(double start) { return static_cast<int>(std::floor(start * 100)); }

// Effective structuring element of cv::dilate(label, dilated, GetUniformKernel(ks, CV_8U)) (alfa_gs.cpp:630-639,
// 747-750).  GetUniformKernel stores two `float 1.0f` (bytes 00 00 80 3F) at byte offsets 0 and ks-1 of a ks-byte
// column and halves it with u8 rounding, so kernel byte i is non-zero exactly where the overlaid float bytes are:
// offsets 2,3 (from the first store; the second store's non-zero bytes fall at ks+1, ks+2, outside the kernel).
// For ks == 3 the second store (offset 2) overwrites byte 2 with 0x00 and byte 3 is outside: the kernel is all
// zero and OpenCV dilates with the single tap at kernel coordinate (0,0).  Pinned against OpenCV 4.13 in tests/golden.
int dilation_taps(int kernel_size_arg, int off[2]) {
    const int ks = std::max(kernel_size_arg - 2, 3);
    std::vector<unsigned char> buf(static_cast<size_t>(ks) + 3, 0);
    const float one = 1.0f;
    std::memcpy(&buf[0], &one, 4);
    std::memcpy(&buf[ks - 1], &one, 4);
    int n = 0;
    for (int i = 0; i < ks; ++i) {
        const int halved = static_cast<int>(std::nearbyint(buf[i] / 2.0));
        if (halved != 0 && n < 2) off[n++] = i - ks / 2;
    }
    if (n == 0) off[n++] = -(ks / 2);
    if (n == 1) off[1] = off[0];
    return n;
}

bool sg_taps(int ws, SgTaps* out) {
    static const float k5[] = {-3, 12, 17, 12, -3};
    static const float k7[] = {-2, 3, 6, 7, 6, 3, -2};
    static const float k9[] = {-21, 14, 39, 54, 59, 54, 39, 14, -21};
    static const float k11[] = {-36, 9, 44, 69, 84, 89, 84, 69, 44, 9, -36};
    const float* src;
    float norm;
    switch (ws) {  // alfa_gs.cpp:653-699
        case 5: src = k5; norm = 35.0f; break;
        case 7: src = k7; norm = 21.0f; break;
        case 9: src = k9; norm = 231.0f; break;
        case 11: src = k11; norm = 429.0f; break;
        default: return false;  // alfa_gs.cpp:645-649: std::logic_error("bad window size")
    }
    // `kernel /= norm` on a CV_32F Mat = convertTo(alpha = 1/norm): tap = numerator * (float)(1.0 / norm)
    const float alpha = static_cast<float>(1.0 / static_cast<double>(norm));
    for (int i = 0; i < ws; ++i) out->t[i] = src[i] * alpha;
    out->ws = ws;
    return true;
}

int derive(const gs_params* p, gs_derived* d) {
    if (!p || !d) return GS_EINVAL;
    if (p->n_beams < 2 || p->n_beams > 1024 || p->rpm <= 0 || p->sample_freq <= 0) return GS_EINVAL;
    const long long w = static_cast<long long>(p->sample_freq) * 60 / p->rpm;  // alfa_gs.cpp:48 (int division)
    if (w < 1 || w > 32767) return GS_EINVAL;                                   // Label::WrapCols takes int16_t (label.h:40)
    if (!(p->max_angle > p->min_angle) || !(p->max_range > p->min_range)) return GS_EINVAL;
    if (p->window_size < 3 || std::max(p->window_size - 2, 3) % 2 == 0) return GS_EINVAL;  // alfa_gs.cpp:631-633
    if (p->smoothing == GS_SMOOTH_SG) {
        SgTaps t;
        if (!sg_taps(p->window_size, &t)) return GS_EINVAL;
    }
    if (p->variant != GS_VARIANT_ALFA_GS && p->variant != GS_VARIANT_CLOUD2RANGE) return GS_EINVAL;
    d->n_rows = p->n_beams;
    d->n_cols = static_cast<int>(w);
    d->d_azimuth = 360.0 / d->n_cols / 180 * M_PI;                       // alfa_gs.cpp:49, utils.h:54
    d->d_altitude = (p->max_angle - p->min_angle) / (p->n_beams - 1);    // alfa_gs.cpp:50
    d->k_ground = edge_cutoff(p->ground_angle_threshold);
    d->s_start = seed_cutoff(p->start_angle_threshold);
    d->n_dilate = dilation_taps(p->window_size, d->dilate_off);
    return GS_OK;
}

gs_params make_params(int n_beams, int sample_freq, double min_angle, double max_angle, double min_range, double max_range,
                      double ground, int window, int variant, int smoothing) {
    gs_params p;
    std::memset(&p, 0, sizeof(p));
    p.n_beams = n_beams;
    p.rpm = 600;
    p.sample_freq = sample_freq;
    p.min_angle = min_angle;
    p.max_angle = max_angle;
    p.min_range = min_range;
    p.max_range = max_range;
    p.ground_angle_threshold = ground;
    p.start_angle_threshold = 0.523598;
    p.window_size = window;
    p.erase_ground = 1;
    p.variant = variant;
    p.pixel_policy = GS_PIXEL_LAST;
    p.col_policy = GS_COL_FLAT_ALIAS;
    p.smoothing = smoothing;
    return p;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct gs_ctx {
    int device = 0;
    gs_params params;
    gs_derived der;
    KParams kp;
    Luts lut;
    SgTaps sg;
    void* lut_block = nullptr;  // one device allocation behind all LUTs
    void* ci_block = nullptr;   // cos / sin tables of gs_backproject_camera_info for the CameraInfo.K values in ci_K
    double ci_K[6] = {0, 0, 0, 0, 0, 0};
    bool ci_valid = false;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    Scratch sc;
    int chunk_frames = 0;  // 0 = automatic
    Counters* d_cnt = nullptr;
    Counters* h_cnt = nullptr;  // pinned
    unsigned long long n_launches = 0;
    bool collect_stats = false;
    bool phase_timing = false;
    bool cas_unions = false;
    // single-frame staging (stage API)
    void* d_points = nullptr;
    size_t d_points_cap = 0;
    uint16_t *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;  // P u16 each
    uint8_t* d_lab = nullptr;                                   // P u8
    int32_t* d_lab32 = nullptr;                                 // P i32
    float* d_xyz = nullptr;                                     // P float4
    // batch staging (host batch API): two slots
    float* d_in[2] = {nullptr, nullptr};
    size_t d_in_cap[2] = {0, 0};
    uint16_t* d_res[2] = {nullptr, nullptr};
    float* d_cloud[2] = {nullptr, nullptr};
    float* d_packed[2] = {nullptr, nullptr};              // the clouds of a wave packed back to back
    unsigned long long* d_prefix[2] = {nullptr, nullptr};  // [slot_frames + 1]
    unsigned* d_counts[2] = {nullptr, nullptr};
    unsigned* h_counts[2] = {nullptr, nullptr};  // pinned
    int slot_frames = 0;
    long long* d_offsets = nullptr;
    size_t d_offsets_cap = 0;
    unsigned* h_count_mapped = nullptr;  // zero-copy word the fused kernel writes a lone frame's point count into
    unsigned* d_count_mapped = nullptr;  // its device alias
    // wave scratch of gs_process_batch_device_labels: no-ground and angle images the CCL kernels read
    uint16_t *lab_res = nullptr, *lab_ang = nullptr;
    int lab_frames = 0;
    FusedState fused;
    std::string err;
};

namespace {

int fail(gs_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        c->err = buf;
    }
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        const cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? GS_ENOMEM : GS_ECUDA, \
                                           "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

Launch launch_of(gs_ctx* c) { return Launch{c->stream, &c->n_launches}; }

void free_scratch(Scratch& s) {
    cudaFree(s.keys); cudaFree(s.range); cudaFree(s.angle); cudaFree(s.angle2); cudaFree(s.parent); cudaFree(s.label);
    cudaFree(s.seed_row); cudaFree(s.eseed); cudaFree(s.rowcount); cudaFree(s.res); cudaFree(s.counts);
    s = Scratch();
}

int ensure_scratch(gs_ctx* ctx, int frames) {
    Scratch& s = ctx->sc;
    const size_t P = ctx->kp.P;
    if (s.frames >= frames && s.P == P && s.W == ctx->kp.W && s.H == ctx->kp.H) return GS_OK;
    free_scratch(s);
    const size_t n = static_cast<size_t>(frames) * P;
    CU(cudaMalloc(&s.keys, n * 8));
    CU(cudaMemsetAsync(s.keys, 0, n * 8, ctx->stream));
    CU(cudaMalloc(&s.range, n * 2));
    CU(cudaMalloc(&s.angle, n * 2));
    CU(cudaMalloc(&s.angle2, n * 2));
    CU(cudaMalloc(&s.parent, (n + 4) * 4));
    CU(cudaMalloc(&s.label, n));
    CU(cudaMalloc(&s.res, n * 2));
    CU(cudaMalloc(&s.seed_row, static_cast<size_t>(frames) * ctx->kp.W * 4));
    CU(cudaMalloc(&s.eseed, static_cast<size_t>(frames) * ctx->kp.W));
    CU(cudaMalloc(&s.rowcount, static_cast<size_t>(frames) * ctx->kp.H * 4));
    CU(cudaMalloc(&s.counts, static_cast<size_t>(frames) * 4));
    s.frames = frames;
    s.P = P;
    s.W = ctx->kp.W;
    s.H = ctx->kp.H;
    return GS_OK;
}

void free_stage(gs_ctx* c) {
    cudaFree(c->d_points); cudaFree(c->d_a); cudaFree(c->d_b); cudaFree(c->d_c); cudaFree(c->d_lab); cudaFree(c->d_lab32);
    cudaFree(c->d_xyz);
    c->d_points = nullptr; c->d_points_cap = 0;
    c->d_a = c->d_b = c->d_c = nullptr; c->d_lab = nullptr; c->d_lab32 = nullptr; c->d_xyz = nullptr;
    cudaFree(c->lab_res); cudaFree(c->lab_ang);
    c->lab_res = c->lab_ang = nullptr; c->lab_frames = 0;
    for (int k = 0; k < 2; ++k) {
        cudaFree(c->d_in[k]); cudaFree(c->d_res[k]); cudaFree(c->d_cloud[k]); cudaFree(c->d_counts[k]);
        cudaFree(c->d_packed[k]); cudaFree(c->d_prefix[k]);
        c->d_packed[k] = nullptr; c->d_prefix[k] = nullptr;
        if (c->h_counts[k]) cudaFreeHost(c->h_counts[k]);
        c->d_in[k] = nullptr; c->d_in_cap[k] = 0; c->d_res[k] = nullptr; c->d_cloud[k] = nullptr; c->d_counts[k] = nullptr;
        c->h_counts[k] = nullptr;
    }
    c->slot_frames = 0;
}

int ensure_stage(gs_ctx* ctx) {
    if (ctx->d_a) return GS_OK;
    const size_t P = ctx->kp.P;
    CU(cudaMalloc(&ctx->d_a, P * 2));
    CU(cudaMalloc(&ctx->d_b, P * 2));
    CU(cudaMalloc(&ctx->d_c, P * 2));
    CU(cudaMalloc(&ctx->d_lab, P));
    CU(cudaMalloc(&ctx->d_lab32, P * 4));
    CU(cudaMalloc(&ctx->d_xyz, P * 16));
    return GS_OK;
}

int ensure_points(gs_ctx* ctx, size_t bytes) {
    if (ctx->d_points_cap >= bytes) return GS_OK;
    cudaFree(ctx->d_points);
    ctx->d_points = nullptr;
    ctx->d_points_cap = 0;
    const size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    CU(cudaMalloc(&ctx->d_points, cap));
    ctx->d_points_cap = cap;
    return GS_OK;
}

int ensure_offsets(gs_ctx* ctx, size_t n) {
    if (ctx->d_offsets_cap >= n) return GS_OK;
    cudaFree(ctx->d_offsets);
    ctx->d_offsets = nullptr;
    CU(cudaMalloc(&ctx->d_offsets, n * sizeof(long long)));
    ctx->d_offsets_cap = n;
    return GS_OK;
}

// Parameters -> kernel constants and device lookup tables.
int configure(gs_ctx* ctx, const gs_params* p) {
    gs_derived d;
    const int rc = derive(p, &d);
    if (rc) return fail(ctx, rc, "invalid parameters (beams %d, rpm %d, sample_freq %d, window %d)", p->n_beams, p->rpm,
                        p->sample_freq, p->window_size);
    const bool shape_changed = !ctx->lut_block || d.n_rows != ctx->kp.H || d.n_cols != ctx->kp.W;
    ctx->params = *p;
    ctx->der = d;
    KParams& k = ctx->kp;
    k.H = d.n_rows;
    k.W = d.n_cols;
    k.P = k.H * k.W;
    k.variant = p->variant;
    k.pixel_policy = p->pixel_policy;
    k.col_policy = p->col_policy;
    k.min_range = p->min_range;
    k.max_range = p->max_range;
    k.span = p->max_range - p->min_range;
    k.min_angle = p->min_angle;
    k.max_angle = p->max_angle;
    k.ang_span = p->max_angle - p->min_angle;
    k.d_alt = d.d_altitude;
    k.d_az = d.d_azimuth;
    k.half_d_alt = d.d_altitude / 2;  // alfa_gs.cpp:249
    k.half_d_az = d.d_azimuth / 2;    // alfa_gs.cpp:250
    k.k_ground = d.k_ground;
    k.s_start = d.s_start;
    k.n_dil = d.n_dilate;
    k.dil0 = d.dilate_off[0];
    k.dil1 = d.dilate_off[1];
    k.erase_ground = (p->variant == GS_VARIANT_CLOUD2RANGE) ? 1 : (p->erase_ground ? 1 : 0);  // EraseGroundBFS has no switch
    ctx->sg.ws = 0;
    if (p->smoothing == GS_SMOOTH_SG) sg_taps(p->window_size, &ctx->sg);

    if (shape_changed) {
        free_scratch(ctx->sc);
        free_stage(ctx);
        fused_release(ctx->fused);
        cudaFree(ctx->lut_block);
        ctx->lut_block = nullptr;
        cudaFree(ctx->ci_block);
        ctx->ci_block = nullptr;
        ctx->ci_valid = false;
    }
    // LUT block: [H] sin_y, [H] cos_y (float) | [H] cos_alt, [H] sin_alt, [W] cos_az, [W] sin_az, [H] sin_y, [H] cos_y (double)
    const int H = k.H, W = k.W;
    const size_t f_bytes = ((2 * static_cast<size_t>(H) * sizeof(float)) + 15) & ~static_cast<size_t>(15);
    const size_t d_count = 4 * static_cast<size_t>(H) + 2 * static_cast<size_t>(W);
    const size_t total = f_bytes + d_count * sizeof(double);
    std::vector<unsigned char> host(total);
    float* hf = reinterpret_cast<float*>(host.data());
    double* hd = reinterpret_cast<double*>(host.data() + f_bytes);
    const float vertical_amplitude = static_cast<float>(p->max_angle - p->min_angle);  // alfa_gs.cpp:344
    for (int r = 0; r < H; ++r) {
        const float angle_y = static_cast<float>(r * d.d_altitude - vertical_amplitude / 2);  // alfa_gs.cpp:357
        hf[r] = sinf(angle_y);      // :366
        hf[H + r] = cosf(angle_y);  // :367
        const double altitude = (H - r) * d.d_altitude + p->min_angle;  // alfa_gs.cpp:1050 / range2cloud_node.cpp:81
        hd[r] = std::cos(altitude);
        hd[H + r] = std::sin(altitude);
        hd[2 * H + 2 * W + r] = static_cast<double>(hf[r]);
        hd[3 * H + 2 * W + r] = static_cast<double>(hf[H + r]);
    }
    for (int c = 0; c < W; ++c) {
        const double azimuth = c * d.d_azimuth;  // alfa_gs.cpp:1052
        hd[2 * H + c] = std::cos(azimuth);
        hd[2 * H + W + c] = std::sin(azimuth);
    }
    if (!ctx->lut_block) CU(cudaMalloc(&ctx->lut_block, total));
    CU(cudaMemcpyAsync(ctx->lut_block, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned char* base = static_cast<unsigned char*>(ctx->lut_block);
    ctx->lut.sin_y = reinterpret_cast<const float*>(base);
    ctx->lut.cos_y = reinterpret_cast<const float*>(base) + H;
    const double* dd = reinterpret_cast<const double*>(base + f_bytes);
    ctx->lut.cos_alt = dd;
    ctx->lut.sin_alt = dd + H;
    ctx->lut.cos_az = dd + 2 * H;
    ctx->lut.sin_az = dd + 2 * H + W;
    ctx->lut.sin_y_d = dd + 2 * H + 2 * W;
    ctx->lut.cos_y_d = dd + 3 * H + 2 * W;
    const int frc = fused_configure(ctx->fused, k, ctx->lut, ctx->sg, p->smoothing == GS_SMOOTH_SG, p->repair_gaps != 0, ctx->device);
    if (frc) return fail(ctx, GS_ECUDA, "fused kernel configuration failed: %s", cudaGetErrorString(static_cast<cudaError_t>(frc)));
    return GS_OK;
}

int auto_chunk(const gs_ctx* c) {
    if (c->chunk_frames > 0) return c->chunk_frames;
    // staged path: keep the ~19 B/pixel of wave scratch inside the 126 MB L2
    const size_t per_frame = static_cast<size_t>(c->kp.P) * 19;
    const int n = static_cast<int>((96ull << 20) / per_frame);
    return std::max(1, std::min(n, 256));
}

int reset_counters(gs_ctx* ctx) {
    CU(cudaMemsetAsync(ctx->d_cnt, 0, sizeof(Counters), ctx->stream));
    return GS_OK;
}

int fetch_counters(gs_ctx* ctx) {
    CU(cudaMemcpyAsync(ctx->h_cnt, ctx->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

// ---- the staged pipeline on a wave of frames resident on the device -----------------------------
// range/angle/res may point at caller buffers (frames slabs); NULL -> context scratch.
int run_wave_staged(gs_ctx* ctx, const void* d_points, size_t stride, const long long* d_offsets, int frames, long long max_points,
                    uint16_t* range_out, uint16_t* angle_out, uint16_t* res_out, float* xyz_out, unsigned* counts_out) {
    const Launch L = launch_of(ctx);
    const KParams& kp = ctx->kp;
    Scratch& s = ctx->sc;
    uint16_t* range = range_out ? range_out : s.range;
    uint16_t* angle = angle_out ? angle_out : s.angle;
    uint16_t* res = res_out ? res_out : s.res;
    launch_project(L, kp, d_points, stride, d_offsets, frames, max_points, s.keys, ctx->d_cnt, ctx->collect_stats);
    launch_finalize(L, kp, frames, s.keys, range);
    const uint16_t* angle_src = range;
    if (ctx->params.repair_gaps) {  // cloud2range_node.cpp:468: RepairGaps(range, 5, 1.0f) feeds the angle image only
        launch_repair(L, kp, frames, range, s.angle2, 5, 1.0f);
        angle_src = s.angle2;
    }
    if (ctx->params.smoothing == GS_SMOOTH_SG) {  // cloud2range_node.cpp:470
        // raw angles go to scratch; the smoothed image is what the ground search (and the caller) sees.  angle2 may hold
        // the repaired range image: k_angle has consumed it before k_sg overwrites it (same stream).
        launch_angle(L, kp, ctx->lut, frames, angle_src, s.angle, ctx->d_cnt, ctx->collect_stats);
        angle = angle_out ? angle_out : s.angle2;
        launch_sg(L, kp, ctx->sg, frames, s.angle, angle);
    } else {
        launch_angle(L, kp, ctx->lut, frames, angle_src, angle, ctx->d_cnt, ctx->collect_stats);
    }
    launch_ccl(L, kp, frames, range, angle, nullptr, kp.k_ground, s.parent);
    launch_ground(L, kp, frames, range, angle, s.parent, s.seed_row, s.eseed, s.label);
    launch_erase(L, kp, frames, range, s.label, res, s.rowcount);
    if (xyz_out || counts_out)  // the counts come out of the back-projection's slot computation (fused kernel: always written)
        launch_backproject(L, kp, ctx->lut, frames, res, s.rowcount, xyz_out, counts_out ? counts_out : s.counts,
                           ctx->params.variant == GS_VARIANT_CLOUD2RANGE);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, GS_ECUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return GS_OK;
}

// Picks the fused single-kernel pipeline when it supports the configuration, else the staged kernels.
int run_wave(gs_ctx* ctx, const void* d_points, size_t stride, const long long* d_offsets, int frames, long long max_points,
             uint16_t* range_out, uint16_t* angle_out, uint16_t* res_out, float* xyz_out, unsigned* counts_out) {
    if (stride == 16 && (reinterpret_cast<uintptr_t>(d_points) & 15) == 0 && fused_supported(ctx->fused) && max_points <= fused_max_points()) {
        const Launch L = launch_of(ctx);
        const int rc = fused_run(ctx->fused, L, static_cast<const float4*>(d_points), d_offsets, frames, max_points, range_out, angle_out, res_out,
                                 reinterpret_cast<float4*>(xyz_out), counts_out, ctx->d_cnt, (ctx->collect_stats ? 1 : 0) | (ctx->phase_timing ? 2 : 0) | (ctx->cas_unions ? 4 : 0));
        if (rc) return fail(ctx, GS_ECUDA, "fused kernel launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(rc)));
        return GS_OK;
    }
    const int rc = ensure_scratch(ctx, std::max(frames, ctx->sc.frames));
    if (rc) return rc;
    return run_wave_staged(ctx, d_points, stride, d_offsets, frames, max_points, range_out, angle_out, res_out, xyz_out, counts_out);
}

bool bad_image_args(const gs_ctx* c, const void* a, const void* b) { return !c || !a || !b; }

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI (the only symbols the library exports: everything else is built with -fvisibility=hidden)
// ------------------------------------------------------------------------------------------------
#pragma GCC visibility push(default)
extern "C" {

int gs_abi_version(void) { return GS_ABI_VERSION; }

int gs_default_params(int preset, gs_params* out) {
    if (!out) return GS_EINVAL;
    switch (preset) {
        case GS_PRESET_ALFA_GS_DEFAULT:  // alfa_gs.cpp:16-38
            *out = make_params(64, 18000, -0.453785, 0.09599, 0.5, 100.0, 0.115, 11, GS_VARIANT_ALFA_GS, GS_SMOOTH_NONE);
            return GS_OK;
        case GS_PRESET_VLP16:  // cloud2range_node.cpp:66-86
            *out = make_params(16, 18000, -0.2617993878, 0.2617993878, 0.5, 100.0, 0.087266, 5, GS_VARIANT_ALFA_GS, GS_SMOOTH_NONE);
            return GS_OK;
        case GS_PRESET_HDL64E_2048:  // alfa_gs.cpp:16-38 with sample_freq 20480
            *out = make_params(64, 20480, -0.453785, 0.09599, 0.5, 100.0, 0.115, 11, GS_VARIANT_ALFA_GS, GS_SMOOTH_NONE);
            return GS_OK;
        case GS_PRESET_OS1_128:  // cloud2range_node.cpp:281-300, beams doubled, sample_freq 20480
            *out = make_params(128, 20480, -0.392699, 0.392699, 0.3, 120.0, 0.09, 7, GS_VARIANT_ALFA_GS, GS_SMOOTH_NONE);
            return GS_OK;
        case GS_PRESET_CLOUD2RANGE_DEFAULT:  // cloud2range_node.cpp:281-300 as shipped; CloudCb runs RepairGaps + SG (:468-470)
            *out = make_params(64, 20000, -0.392699, 0.392699, 0.3, 120.0, 0.09, 7, GS_VARIANT_CLOUD2RANGE, GS_SMOOTH_SG);
            out->repair_gaps = 1;
            return GS_OK;
        case GS_PRESET_HDL32:  // cloud2range_node.cpp:122-135
            *out = make_params(32, 18000, -0.535293, 0.186227, 0.5, 100.0, 0.087266, 5, GS_VARIANT_ALFA_GS, GS_SMOOTH_NONE);
            return GS_OK;
        default: return GS_EINVAL;
    }
}

int gs_derive(const gs_params* p, gs_derived* out) { return derive(p, out); }

int gs_create(const gs_params* p, int device, gs_ctx** out) {
    if (!p || !out) return GS_EINVAL;
    *out = nullptr;
    gs_derived d;
    if (derive(p, &d) != GS_OK) return GS_EINVAL;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return GS_ECUDA;  // no CPU fallback
    if (cudaSetDevice(device) != cudaSuccess) return GS_ECUDA;
    gs_ctx* ctx = new gs_ctx();
    ctx->device = device;
    std::memset(&ctx->kp, 0, sizeof(ctx->kp));
    bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int k = 0; ok && k < 2; ++k)
        ok = cudaEventCreateWithFlags(&ctx->ev_in[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_done[k], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_out[k], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc(&ctx->d_cnt, sizeof(Counters)) == cudaSuccess &&
         cudaMallocHost(&ctx->h_cnt, sizeof(Counters)) == cudaSuccess;
    if (!ok) {
        gs_destroy(ctx);
        return GS_ECUDA;
    }
    ctx->stream = ctx->own_stream;
    std::memset(ctx->h_cnt, 0, sizeof(Counters));
    cudaMemsetAsync(ctx->d_cnt, 0, sizeof(Counters), ctx->stream);
    const int rc = configure(ctx, p);
    if (rc) {
        gs_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return GS_OK;
}

void gs_destroy(gs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    free_scratch(ctx->sc);
    free_stage(ctx);
    fused_release(ctx->fused);
    cudaFree(ctx->lut_block);
    cudaFree(ctx->ci_block);
    if (ctx->h_count_mapped) cudaFreeHost(ctx->h_count_mapped);
    cudaFree(ctx->d_cnt);
    cudaFree(ctx->d_offsets);
    if (ctx->h_cnt) cudaFreeHost(ctx->h_cnt);
    for (int k = 0; k < 2; ++k) {
        if (ctx->ev_in[k]) cudaEventDestroy(ctx->ev_in[k]);
        if (ctx->ev_done[k]) cudaEventDestroy(ctx->ev_done[k]);
        if (ctx->ev_out[k]) cudaEventDestroy(ctx->ev_out[k]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    delete ctx;
}

int gs_set_params(gs_ctx* ctx, const gs_params* p) {
    if (!ctx || !p) return GS_EINVAL;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return configure(ctx, p);
}

int gs_get_params(const gs_ctx* ctx, gs_params* out) {
    if (!ctx || !out) return GS_EINVAL;
    *out = ctx->params;
    return GS_OK;
}

int gs_get_derived(const gs_ctx* ctx, gs_derived* out) {
    if (!ctx || !out) return GS_EINVAL;
    *out = ctx->der;
    return GS_OK;
}

int gs_set_stream(gs_ctx* ctx, void* cuda_stream) {
    if (!ctx) return GS_EINVAL;
    cudaSetDevice(ctx->device);
    // work queued on the old stream uses the context's scratch (remote lists, offsets): let it finish before another stream may touch it
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return GS_OK;
}

int gs_set_chunk_frames(gs_ctx* ctx, int frames) {
    if (!ctx || frames < 0 || frames > 4096) return GS_EINVAL;
    ctx->chunk_frames = frames;
    return GS_OK;
}

int gs_set_option(gs_ctx* ctx, int option, int value) {
    if (!ctx) return GS_EINVAL;
    switch (option) {
        case GS_OPT_COLLECT_STATS: ctx->collect_stats = value != 0; return GS_OK;
        case GS_OPT_FORCE_STAGED: ctx->fused.disabled = value != 0; return GS_OK;
        case GS_OPT_PHASE_TIMING: ctx->phase_timing = value != 0; return GS_OK;
        case GS_OPT_CAS_UNIONS: ctx->cas_unions = value != 0; return GS_OK;
        case GS_OPT_NO_LONE_PLAN: ctx->fused.no_lone = value != 0; return GS_OK;
        default: return fail(ctx, GS_EINVAL, "unknown option %d", option);
    }
}

int gs_get_stats(gs_ctx* ctx, gs_stats* out) {
    if (!ctx || !out) return GS_EINVAL;
    cudaSetDevice(ctx->device);
    const int rc = fetch_counters(ctx);
    if (rc) return rc;
    out->kernel_launches = ctx->n_launches;
    out->near_edge_project = ctx->h_cnt->near_edge_project;
    out->near_edge_angle = ctx->h_cnt->near_edge_angle;
    out->dropped_points = ctx->h_cnt->dropped;
    out->slow_path_points = ctx->h_cnt->slow_path;
    out->fast_path_mismatches = ctx->h_cnt->fast_mismatch;
    out->remote_points = ctx->h_cnt->remote;
    return GS_OK;
}

int gs_get_phase_cycles(gs_ctx* ctx, uint64_t* out16) {
    if (!ctx || !out16) return GS_EINVAL;
    cudaSetDevice(ctx->device);
    const int rc = fetch_counters(ctx);
    if (rc) return rc;
    for (int k = 0; k < 16; ++k) out16[k] = ctx->h_cnt->phase[k];

    return GS_OK;
}

int gs_reset_stats(gs_ctx* ctx) {
    if (!ctx) return GS_EINVAL;
    cudaSetDevice(ctx->device);
    return reset_counters(ctx);
}

const char* gs_last_error(const gs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gs_create_labeled_range_image(gs_ctx* ctx, const void* points_xyzrgb32, size_t n_points, uint16_t* labeled_out) {
    if (!ctx || !labeled_out || (!points_xyzrgb32 && n_points)) return fail(ctx, GS_EINVAL, "gs_create_labeled_range_image: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    const size_t raw = (n_points * 32 + 15) / 16 * 16;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_points(ctx, raw + label_image_scratch_bytes(n_points)))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    uint8_t* flags = base + raw;
    const size_t chunks = (n_points + 255) / 256;
    uint8_t* summary = flags + (n_points + 15) / 16 * 16;
    int* flat = reinterpret_cast<int*>(summary + (chunks + 15) / 16 * 16);
    if (n_points) CU(cudaMemcpyAsync(base, points_xyzrgb32, n_points * 32, cudaMemcpyHostToDevice, ctx->stream));
    launch_label_image(launch_of(ctx), ctx->kp, base, n_points, flags, flat, summary, ctx->d_a);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(labeled_out, ctx->d_a, static_cast<size_t>(ctx->kp.P) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_evaluate(gs_ctx* ctx, const uint16_t* og_range, const uint16_t* seg_range, const uint16_t* labeled_range, gs_eval* out) {
    if (!ctx || !og_range || !seg_range || !labeled_range || !out) return fail(ctx, GS_EINVAL, "gs_evaluate: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx))) return rc;
    const size_t bytes = static_cast<size_t>(ctx->kp.P) * 2;
    unsigned long long* d9 = reinterpret_cast<unsigned long long*>(ctx->d_lab32);  // P * 4 bytes >= 72
    CU(cudaMemcpyAsync(ctx->d_a, og_range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_b, seg_range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_c, labeled_range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d9, 0, 9 * sizeof(unsigned long long), ctx->stream));
    launch_evaluate(launch_of(ctx), ctx->kp, ctx->d_a, ctx->d_b, ctx->d_c, d9);
    CU(cudaGetLastError());
    unsigned long long h9[9];
    CU(cudaMemcpyAsync(h9, d9, sizeof(h9), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    out->points = h9[0];
    out->labeled_gnd = h9[1];
    out->labeled_obst = h9[2];
    out->seg_gnd = h9[3];
    out->true_positive = h9[4];
    out->false_positive = h9[5];
    out->true_negative = h9[6];
    out->false_negative = h9[7];
    out->true_detected_obst = h9[8];
    // the reference counts in float (alfa_gs.cpp:834): increments stop at 2^24
    auto as_float_counter = [](unsigned long long v) { return static_cast<float>(std::min<unsigned long long>(v, 16777216ull)); };
    const float tp = as_float_counter(h9[4]), fp = as_float_counter(h9[5]), tn = as_float_counter(h9[6]), fn = as_float_counter(h9[7]);
    out->iou_g = (tp / (tp + fp + fn)) * 100;   // :885
    out->recall_g = (tp / (tp + fn)) * 100;     // :886
    out->recall_mo = (tn / (tn + fp)) * 100;    // :887
    return GS_OK;
}

int gs_synchronize(gs_ctx* ctx) {
    if (!ctx) return GS_EINVAL;
    cudaSetDevice(ctx->device);
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

// ---- stage API -------------------------------------------------------------------------------
int gs_project(gs_ctx* ctx, const void* points, size_t n_points, size_t stride_bytes, uint16_t* range_out) {
    if (!ctx || !range_out || (n_points && !points) || stride_bytes < 12 || (stride_bytes & 3)) return fail(ctx, GS_EINVAL, "gs_project: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_scratch(ctx, 1)) || (rc = ensure_points(ctx, n_points * stride_bytes)) ||
        (rc = ensure_offsets(ctx, 2)))
        return rc;
    const long long offs[2] = {0, static_cast<long long>(n_points)};
    CU(cudaMemcpyAsync(ctx->d_offsets, offs, sizeof(offs), cudaMemcpyHostToDevice, ctx->stream));
    if (n_points) CU(cudaMemcpyAsync(ctx->d_points, points, n_points * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
    const Launch L = launch_of(ctx);
    launch_project(L, ctx->kp, ctx->d_points, stride_bytes, ctx->d_offsets, 1, static_cast<long long>(n_points), ctx->sc.keys, ctx->d_cnt,
                   ctx->collect_stats);
    launch_finalize(L, ctx->kp, 1, ctx->sc.keys, ctx->d_a);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(range_out, ctx->d_a, static_cast<size_t>(ctx->kp.P) * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_angle_image(gs_ctx* ctx, const uint16_t* range, uint16_t* angle_out) {
    if (bad_image_args(ctx, range, angle_out)) return fail(ctx, GS_EINVAL, "gs_angle_image: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx))) return rc;
    const size_t bytes = static_cast<size_t>(ctx->kp.P) * 2;
    CU(cudaMemcpyAsync(ctx->d_a, range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_angle(launch_of(ctx), ctx->kp, ctx->lut, 1, ctx->d_a, ctx->d_b, ctx->d_cnt, ctx->collect_stats);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(angle_out, ctx->d_b, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_smooth(gs_ctx* ctx, const uint16_t* angle, int window, uint16_t* smoothed_out) {
    if (bad_image_args(ctx, angle, smoothed_out)) return fail(ctx, GS_EINVAL, "gs_smooth: bad arguments");
    SgTaps taps;
    if (window % 2 == 0 || !sg_taps(window, &taps)) return fail(ctx, GS_EINVAL, "bad window size %d", window);  // alfa_gs.cpp:642-649
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx))) return rc;
    const size_t bytes = static_cast<size_t>(ctx->kp.P) * 2;
    CU(cudaMemcpyAsync(ctx->d_a, angle, bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_sg(launch_of(ctx), ctx->kp, taps, 1, ctx->d_a, ctx->d_b);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(smoothed_out, ctx->d_b, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_repair_gaps(gs_ctx* ctx, const uint16_t* range, int step, float depth_threshold, uint16_t* repaired_out) {
    if (bad_image_args(ctx, range, repaired_out) || step < 1) return fail(ctx, GS_EINVAL, "gs_repair_gaps: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx))) return rc;
    const size_t bytes = static_cast<size_t>(ctx->kp.P) * 2;
    CU(cudaMemcpyAsync(ctx->d_a, range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_repair(launch_of(ctx), ctx->kp, 1, ctx->d_a, ctx->d_b, step, depth_threshold);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(repaired_out, ctx->d_b, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_erase_ground(gs_ctx* ctx, const uint16_t* range, const uint16_t* angle, double ground_angle_threshold,
                    double start_angle_threshold, int kernel_size, uint16_t* res_out, uint8_t* labels_out) {
    if (!ctx || !range || !angle || !res_out) return fail(ctx, GS_EINVAL, "gs_erase_ground: bad arguments");
    if (kernel_size < 3 || std::max(kernel_size - 2, 3) % 2 == 0)
        return fail(ctx, GS_EINVAL, "Window size must be odd (kernel_size %d)", kernel_size);  // alfa_gs.cpp:631-633
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_scratch(ctx, 1))) return rc;
    KParams kp = ctx->kp;
    kp.k_ground = edge_cutoff(ground_angle_threshold);
    kp.s_start = seed_cutoff(start_angle_threshold);
    int off[2];
    kp.n_dil = dilation_taps(kernel_size, off);
    kp.dil0 = off[0];
    kp.dil1 = off[1];
    const size_t bytes = static_cast<size_t>(kp.P) * 2;
    CU(cudaMemcpyAsync(ctx->d_a, range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_b, angle, bytes, cudaMemcpyHostToDevice, ctx->stream));
    const Launch L = launch_of(ctx);
    launch_ccl(L, kp, 1, ctx->d_a, ctx->d_b, nullptr, kp.k_ground, ctx->sc.parent);
    launch_ground(L, kp, 1, ctx->d_a, ctx->d_b, ctx->sc.parent, ctx->sc.seed_row, ctx->sc.eseed, ctx->sc.label);
    launch_erase(L, kp, 1, ctx->d_a, ctx->sc.label, ctx->d_c, ctx->sc.rowcount);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(res_out, ctx->d_c, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (labels_out) CU(cudaMemcpyAsync(labels_out, ctx->sc.label, static_cast<size_t>(kp.P), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_label_one_component(gs_ctx* ctx, int label, int row, int col, float angle_threshold, const uint16_t* range,
                           const uint16_t* angle, uint8_t* label_image) {
    if (!ctx || !range || !angle || !label_image) return fail(ctx, GS_EINVAL, "gs_label_one_component: bad arguments");
    if (row < 0 || row >= ctx->kp.H || col < 0 || col >= ctx->kp.W) return fail(ctx, GS_EINVAL, "seed (%d,%d) outside the image", row, col);
    if ((label & 0xff) == 0) return fail(ctx, GS_EINVAL, "label 0 never terminates in the reference (label.cpp:61-67)");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_scratch(ctx, 1))) return rc;
    const KParams& kp = ctx->kp;
    const size_t bytes = static_cast<size_t>(kp.P) * 2;
    const int k_edge = edge_cutoff(static_cast<double>(angle_threshold));
    CU(cudaMemcpyAsync(ctx->d_a, range, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_b, angle, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_lab, label_image, static_cast<size_t>(kp.P), cudaMemcpyHostToDevice, ctx->stream));
    const Launch L = launch_of(ctx);
    launch_ccl(L, kp, 1, ctx->d_a, ctx->d_b, ctx->d_lab, k_edge, ctx->sc.parent);
    launch_one_component(L, kp, ctx->d_a, ctx->d_b, k_edge, ctx->sc.parent, row, col, label & 0xff, ctx->d_lab);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(label_image, ctx->d_lab, static_cast<size_t>(kp.P), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_label_components(gs_ctx* ctx, const uint16_t* image, const uint16_t* angle, double beta, int32_t* labels_out,
                        int32_t* n_components_out) {
    if (!ctx || !image || !angle || !labels_out) return fail(ctx, GS_EINVAL, "gs_label_components: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_scratch(ctx, 1))) return rc;
    const KParams& kp = ctx->kp;
    const size_t bytes = static_cast<size_t>(kp.P) * 2;
    CU(cudaMemcpyAsync(ctx->d_a, image, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_b, angle, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(&ctx->d_cnt->components, 0, sizeof(unsigned long long), ctx->stream));
    const Launch L = launch_of(ctx);
    launch_components_batch(L, kp, 1, ctx->d_a, ctx->d_b, edge_cutoff(beta), ctx->d_lab32, nullptr, ctx->d_cnt);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(labels_out, ctx->d_lab32, static_cast<size_t>(kp.P) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = fetch_counters(ctx))) return rc;
    if (n_components_out) *n_components_out = static_cast<int32_t>(ctx->h_cnt->components);
    return GS_OK;
}

int gs_backproject(gs_ctx* ctx, const uint16_t* image, float* xyz4_out, uint32_t* n_out) {
    if (!ctx || !image || !xyz4_out || !n_out) return fail(ctx, GS_EINVAL, "gs_backproject: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_scratch(ctx, 1))) return rc;
    const KParams& kp = ctx->kp;
    CU(cudaMemcpyAsync(ctx->d_a, image, static_cast<size_t>(kp.P) * 2, cudaMemcpyHostToDevice, ctx->stream));
    const Launch L = launch_of(ctx);
    launch_rowcount(L, kp, 1, ctx->d_a, ctx->sc.rowcount);
    launch_backproject(L, kp, ctx->lut, 1, ctx->d_a, ctx->sc.rowcount, ctx->d_xyz, ctx->sc.counts,
                       ctx->params.variant == GS_VARIANT_CLOUD2RANGE);
    CU(cudaGetLastError());
    unsigned n = 0;
    CU(cudaMemcpyAsync(&n, ctx->sc.counts, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n) CU(cudaMemcpyAsync(xyz4_out, ctx->d_xyz, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return GS_OK;
}

int gs_backproject_camera_info(gs_ctx* ctx, const uint16_t* image, int32_t height, int32_t width, const double* K, int range_as_double,
                               float* xyz4_out, uint32_t* n_out) {
    if (!ctx || !image || !K || !xyz4_out || !n_out) return fail(ctx, GS_EINVAL, "gs_backproject_camera_info: bad arguments");
    if (height != ctx->kp.H || width != ctx->kp.W)
        return fail(ctx, GS_EINVAL, "gs_backproject_camera_info: image is %d x %d, the context %d x %d", height, width, ctx->kp.H, ctx->kp.W);
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_scratch(ctx, 1))) return rc;
    const int H = ctx->kp.H, W = ctx->kp.W;
    if (!ctx->ci_valid || std::memcmp(ctx->ci_K, K, sizeof(ctx->ci_K)) != 0) {
        // the decoding parameters are the ones the message carries (range2cloud_node.cpp:51-58, alfa_gs.cpp:1010-1015):
        // altitude = (rows - r) * K[5] + K[0], azimuth = c * K[4]; cos / sin with the host's libm like the reference
        std::vector<double> host(2 * static_cast<size_t>(H) + 2 * static_cast<size_t>(W));
        for (int r = 0; r < H; ++r) {
            const double altitude = (H - r) * K[5] + K[0];
            host[r] = std::cos(altitude);
            host[H + r] = std::sin(altitude);
        }
        for (int c = 0; c < W; ++c) {
            const double azimuth = c * K[4];
            host[2 * H + c] = std::cos(azimuth);
            host[2 * H + W + c] = std::sin(azimuth);
        }
        if (!ctx->ci_block) CU(cudaMalloc(&ctx->ci_block, host.size() * sizeof(double)));
        CU(cudaMemcpyAsync(ctx->ci_block, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        std::memcpy(ctx->ci_K, K, sizeof(ctx->ci_K));
        ctx->ci_valid = true;
    }
    KParams kp = ctx->kp;
    kp.min_range = K[2];
    kp.max_range = K[3];
    kp.span = K[3] - K[2];
    Luts lut = ctx->lut;
    const double* dd = static_cast<const double*>(ctx->ci_block);
    lut.cos_alt = dd;
    lut.sin_alt = dd + H;
    lut.cos_az = dd + 2 * H;
    lut.sin_az = dd + 2 * H + W;
    CU(cudaMemcpyAsync(ctx->d_a, image, static_cast<size_t>(kp.P) * 2, cudaMemcpyHostToDevice, ctx->stream));
    const Launch L = launch_of(ctx);
    launch_rowcount(L, kp, 1, ctx->d_a, ctx->sc.rowcount);
    launch_backproject(L, kp, lut, 1, ctx->d_a, ctx->sc.rowcount, ctx->d_xyz, ctx->sc.counts, range_as_double != 0);
    CU(cudaGetLastError());
    unsigned n = 0;
    CU(cudaMemcpyAsync(&n, ctx->sc.counts, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (n) CU(cudaMemcpyAsync(xyz4_out, ctx->d_xyz, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return GS_OK;
}

int gs_process_frame(gs_ctx* ctx, const void* points, size_t n_points, size_t stride_bytes, uint16_t* range_out,
                     uint16_t* angle_out, uint16_t* no_ground_out, float* xyz4_out, uint32_t* n_out) {
    if (!ctx || (n_points && !points) || stride_bytes < 12 || (stride_bytes & 3)) return fail(ctx, GS_EINVAL, "gs_process_frame: bad arguments");
    // records wider than 16 bytes (pcl::PointXYZRGB, alfa_gs.h:25) are repacked on the device and take the same fused kernel
    if (stride_bytes != 16)
        return gs_process_cloud2(ctx, points, n_points, static_cast<uint32_t>(stride_bytes), 0, 4, 8, 0, range_out, angle_out, no_ground_out,
                                 xyz4_out, n_out);
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_points(ctx, n_points * stride_bytes)) || (rc = ensure_offsets(ctx, 2))) return rc;
    const KParams& kp = ctx->kp;
    const size_t bytes = static_cast<size_t>(kp.P) * 2;
    const long long offs[2] = {0, static_cast<long long>(n_points)};
    CU(cudaMemcpyAsync(ctx->d_offsets, offs, sizeof(offs), cudaMemcpyHostToDevice, ctx->stream));
    if (n_points) CU(cudaMemcpyAsync(ctx->d_points, points, n_points * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
    // Latency path (the reference's execution model is one frame per callback, alfa_node.cpp:298-319): when the caller's output
    // buffers are page-locked host memory the device can address, the fused kernel stores the image, the cloud and the point
    // count straight into them while it computes — no device-to-host copies behind the kernel and ONE synchronisation.
    {
        auto visible = [](const void* p, void** d) -> bool {
            if (!p) { *d = nullptr; return true; }
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
            if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
            *d = a.devicePointer;
            return true;
        };
        void *d_rng = nullptr, *d_ang = nullptr, *d_res = nullptr, *d_xyz = nullptr;
        const bool fused_ok = fused_supported(ctx->fused) && static_cast<long long>(n_points) <= fused_max_points();
        if (fused_ok && (no_ground_out || xyz4_out) && visible(range_out, &d_rng) && visible(angle_out, &d_ang) && visible(no_ground_out, &d_res) &&
            visible(xyz4_out, &d_xyz)) {
            if (!ctx->h_count_mapped) {
                CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_count_mapped), 64, cudaHostAllocMapped));
                CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->d_count_mapped), ctx->h_count_mapped, 0));
            }
            *ctx->h_count_mapped = 0;
            if ((rc = run_wave(ctx, ctx->d_points, stride_bytes, ctx->d_offsets, 1, static_cast<long long>(n_points), static_cast<uint16_t*>(d_rng),
                               static_cast<uint16_t*>(d_ang), static_cast<uint16_t*>(d_res), static_cast<float*>(d_xyz), ctx->d_count_mapped)))
                return rc;
            CU(cudaStreamSynchronize(ctx->stream));
            if (n_out) *n_out = *ctx->h_count_mapped;
            return GS_OK;
        }
    }
    if ((rc = ensure_scratch(ctx, 1))) return rc;
    if ((rc = run_wave(ctx, ctx->d_points, stride_bytes, ctx->d_offsets, 1, static_cast<long long>(n_points), ctx->d_a, ctx->d_b, ctx->d_c,
                       ctx->d_xyz, ctx->sc.counts)))
        return rc;
    if (range_out) CU(cudaMemcpyAsync(range_out, ctx->d_a, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (angle_out) CU(cudaMemcpyAsync(angle_out, ctx->d_b, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (no_ground_out) CU(cudaMemcpyAsync(no_ground_out, ctx->d_c, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    unsigned n = 0;
    CU(cudaMemcpyAsync(&n, ctx->sc.counts, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (xyz4_out && n) {
        CU(cudaMemcpyAsync(xyz4_out, ctx->d_xyz, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (n_out) *n_out = n;
    return GS_OK;
}

// ---- wire formats (N1) -------------------------------------------------------------------------
namespace {
// device-side unpack of a PointCloud2-style host buffer into ctx->d_points as packed float4 (staging behind it)
int unpack_to_device(gs_ctx* ctx, const void* data, size_t n, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, int32_t oi, int big_endian) {
    const size_t raw = n * static_cast<size_t>(step), packed = n * 16;
    const size_t raw_off = (packed + 255) & ~static_cast<size_t>(255);
    int rc;
    if ((rc = ensure_points(ctx, raw_off + raw))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    if (n) CU(cudaMemcpyAsync(base + raw_off, data, raw, cudaMemcpyHostToDevice, ctx->stream));
    launch_unpack_points(launch_of(ctx), base + raw_off, n, step, ox, oy, oz, oi, big_endian, reinterpret_cast<float*>(base));
    CU(cudaGetLastError());
    return GS_OK;
}
bool bad_fields(size_t n, const void* data, uint32_t step, uint32_t ox, uint32_t oy, uint32_t oz, int32_t oi) {
    if (n && !data) return true;
    if (step < 12) return true;
    if (ox + 4 > step || oy + 4 > step || oz + 4 > step) return true;
    if (oi >= 0 && static_cast<uint32_t>(oi) + 4 > step) return true;
    return false;
}
}  // namespace

int gs_unpack_points(gs_ctx* ctx, const void* data, size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                     int32_t off_intensity, int is_bigendian, float* xyzi_out) {
    if (!ctx || !xyzi_out || bad_fields(n_points, data, point_step, off_x, off_y, off_z, off_intensity)) return fail(ctx, GS_EINVAL, "gs_unpack_points: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = unpack_to_device(ctx, data, n_points, point_step, off_x, off_y, off_z, off_intensity, is_bigendian))) return rc;
    if (n_points) CU(cudaMemcpyAsync(xyzi_out, ctx->d_points, n_points * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_process_cloud2(gs_ctx* ctx, const void* data, size_t n_points, uint32_t point_step, uint32_t off_x, uint32_t off_y, uint32_t off_z,
                      int is_bigendian, uint16_t* range_out, uint16_t* angle_out, uint16_t* no_ground_out, float* xyz4_out, uint32_t* n_out) {
    if (!ctx || bad_fields(n_points, data, point_step, off_x, off_y, off_z, -1)) return fail(ctx, GS_EINVAL, "gs_process_cloud2: bad arguments");
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_stage(ctx)) || (rc = ensure_offsets(ctx, 2))) return rc;
    const KParams& kp = ctx->kp;
    const size_t bytes = static_cast<size_t>(kp.P) * 2;
    const long long offs[2] = {0, static_cast<long long>(n_points)};
    CU(cudaMemcpyAsync(ctx->d_offsets, offs, sizeof(offs), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = unpack_to_device(ctx, data, n_points, point_step, off_x, off_y, off_z, -1, is_bigendian))) return rc;
    if ((rc = ensure_scratch(ctx, 1))) return rc;
    if ((rc = run_wave(ctx, ctx->d_points, 16, ctx->d_offsets, 1, static_cast<long long>(n_points), ctx->d_a, ctx->d_b, ctx->d_c, ctx->d_xyz,
                       ctx->sc.counts)))
        return rc;
    if (range_out) CU(cudaMemcpyAsync(range_out, ctx->d_a, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (angle_out) CU(cudaMemcpyAsync(angle_out, ctx->d_b, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (no_ground_out) CU(cudaMemcpyAsync(no_ground_out, ctx->d_c, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    unsigned n = 0;
    CU(cudaMemcpyAsync(&n, ctx->sc.counts, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (xyz4_out && n) {
        CU(cudaMemcpyAsync(xyz4_out, ctx->d_xyz, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    if (n_out) *n_out = n;
    return GS_OK;
}

int gs_pack_cloud_xyzrgb(gs_ctx* ctx, const float* xyz4, uint32_t n_points, void* out32) {
    if (!ctx || (n_points && (!xyz4 || !out32))) return fail(ctx, GS_EINVAL, "gs_pack_cloud_xyzrgb: bad arguments");
    if (n_points == 0) return GS_OK;
    cudaSetDevice(ctx->device);
    int rc;
    const size_t in_bytes = static_cast<size_t>(n_points) * 16, out_off = (in_bytes + 255) & ~static_cast<size_t>(255);
    if ((rc = ensure_points(ctx, out_off + static_cast<size_t>(n_points) * 32))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    CU(cudaMemcpyAsync(base, xyz4, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_pack_xyzrgb(launch_of(ctx), reinterpret_cast<const float*>(base), n_points, base + out_off);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out32, base + out_off, static_cast<size_t>(n_points) * 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

// ---- N4: AlfaNode's FPGA wire formats and the coloured angle image (host buffers; the context lends its stream and staging) ----
namespace {
int stage_bytes(gs_ctx* ctx, size_t bytes) { return ensure_points(ctx, bytes); }
}  // namespace

int gs_fpga_store_pointcloud(gs_ctx* ctx, const void* points, size_t n_points, size_t stride_bytes, uint64_t* words_out) {
    if (!ctx || (n_points && (!points || !words_out)) || stride_bytes < 12 || (stride_bytes & 3)) return fail(ctx, GS_EINVAL, "gs_fpga_store_pointcloud: bad arguments");
    if (n_points == 0) return GS_OK;
    cudaSetDevice(ctx->device);
    const size_t in_bytes = (n_points * stride_bytes + 255) & ~static_cast<size_t>(255);
    int rc;
    if ((rc = stage_bytes(ctx, in_bytes + n_points * 8))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    CU(cudaMemcpyAsync(base, points, n_points * stride_bytes, cudaMemcpyHostToDevice, ctx->stream));
    launch_fpga_store(launch_of(ctx), base, n_points, stride_bytes, reinterpret_cast<unsigned long long*>(base + in_bytes));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(words_out, base + in_bytes, n_points * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_fpga_read_range_image(gs_ctx* ctx, const uint64_t* words, uint32_t rows, uint32_t cols, uint16_t* image64x1800_out) {
    if (!ctx || !words || !image64x1800_out) return fail(ctx, GS_EINVAL, "gs_fpga_read_range_image: bad arguments");
    const size_t n = static_cast<size_t>(rows) * cols / 4 * 4;  // whole words only, like the reference's loop (alfa_node.cpp:165,170)
    if (n > static_cast<size_t>(64) * 1800) return fail(ctx, GS_EINVAL, "gs_fpga_read_range_image: %u x %u pixels do not fit the 64 x 1800 image the reference hard-codes", rows, cols);
    cudaSetDevice(ctx->device);
    const size_t img_bytes = static_cast<size_t>(64) * 1800 * 2;
    int rc;
    if ((rc = stage_bytes(ctx, img_bytes + n * 2 + 256))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    CU(cudaMemsetAsync(base, 0, img_bytes, ctx->stream));  // Mat::zeros (:162)
    if (n) {
        CU(cudaMemcpyAsync(base + img_bytes, words, n * 2, cudaMemcpyHostToDevice, ctx->stream));
        launch_fpga_read_range(launch_of(ctx), reinterpret_cast<const unsigned long long*>(base + img_bytes), static_cast<unsigned>(n),
                               reinterpret_cast<uint16_t*>(base));
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(image64x1800_out, base, img_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_fpga_read_angle_image(gs_ctx* ctx, const uint64_t* words, uint32_t rows, uint32_t cols, uint16_t* image_out) {
    if (!ctx || !words || !image_out) return fail(ctx, GS_EINVAL, "gs_fpga_read_angle_image: bad arguments");
    if (rows != 64) return fail(ctx, GS_EINVAL, "gs_fpga_read_angle_image: the reference's unpacking counts 64 rows per column (alfa_node.cpp:219,240), got %u", rows);
    const size_t n = static_cast<size_t>(rows) * cols;
    if (n == 0) return GS_OK;
    const size_t n_words = (n >> 8) * 43 + (((n & 255) < 252) ? ((n & 255) + 5) / 6 : 43);
    cudaSetDevice(ctx->device);
    const size_t img_bytes = (n * 2 + 255) & ~static_cast<size_t>(255);
    int rc;
    if ((rc = stage_bytes(ctx, img_bytes + n_words * 8))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    CU(cudaMemcpyAsync(base + img_bytes, words, n_words * 8, cudaMemcpyHostToDevice, ctx->stream));
    launch_fpga_read_angle(launch_of(ctx), reinterpret_cast<const unsigned long long*>(base + img_bytes), static_cast<unsigned>(n), cols,
                           reinterpret_cast<uint16_t*>(base));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(image_out, base, n * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_colored_angle_image(gs_ctx* ctx, const uint16_t* angle, int32_t rows, int32_t cols, uint8_t* rgb_out) {
    if (!ctx || !angle || !rgb_out || rows < 0 || cols < 0) return fail(ctx, GS_EINVAL, "gs_colored_angle_image: bad arguments");
    const size_t n = static_cast<size_t>(rows) * cols;
    if (n == 0) return GS_OK;
    cudaSetDevice(ctx->device);
    const size_t in_bytes = (n * 2 + 255) & ~static_cast<size_t>(255);
    int rc;
    if ((rc = stage_bytes(ctx, in_bytes + n * 3))) return rc;
    unsigned char* base = static_cast<unsigned char*>(ctx->d_points);
    CU(cudaMemcpyAsync(base, angle, n * 2, cudaMemcpyHostToDevice, ctx->stream));
    launch_colored_angle(launch_of(ctx), reinterpret_cast<const uint16_t*>(base), static_cast<unsigned>(n), base + in_bytes);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rgb_out, base + in_bytes, n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
}

int gs_get_camera_info(const gs_ctx* ctx, int32_t* height, int32_t* width, double* K9) {
    if (!ctx || !height || !width || !K9) return GS_EINVAL;
    *height = ctx->der.n_rows;
    *width = ctx->der.n_cols;
    const double k[9] = {ctx->params.min_angle, ctx->params.max_angle, ctx->params.min_range, ctx->params.max_range, ctx->der.d_azimuth,
                         ctx->der.d_altitude, 0, 0, 0};
    for (int i = 0; i < 9; ++i) K9[i] = k[i];
    return GS_OK;
}

// ---- batch API -------------------------------------------------------------------------------
int gs_process_batch_device(gs_ctx* ctx, const float* d_points_xyzi, const int64_t* h_offsets, int n_frames, uint16_t* d_range_out,
                            uint16_t* d_angle_out, uint16_t* d_no_ground_out, float* d_xyz4_out, uint32_t* d_counts_out) {
    if (!ctx || !h_offsets || n_frames < 0 || (n_frames && !d_points_xyzi)) return fail(ctx, GS_EINVAL, "gs_process_batch_device: bad arguments");
    if (n_frames == 0) return GS_OK;
    cudaSetDevice(ctx->device);
    int rc;
    if ((rc = ensure_offsets(ctx, static_cast<size_t>(n_frames) + 1))) return rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    CU(cudaMemcpyAsync(ctx->d_offsets, h_offsets, (static_cast<size_t>(n_frames) + 1) * sizeof(long long), cudaMemcpyHostToDevice,
                       ctx->stream));
    const size_t P = ctx->kp.P;
    long long batch_max = 0;
    for (int f = 0; f < n_frames; ++f) batch_max = std::max<long long>(batch_max, h_offsets[f + 1] - h_offsets[f]);
    const bool use_fused = (reinterpret_cast<uintptr_t>(d_points_xyzi) & 15) == 0 && fused_supported(ctx->fused) && batch_max <= fused_max_points();
    const int chunk = use_fused ? n_frames : std::min(n_frames, auto_chunk(ctx));
    if (!use_fused && (rc = ensure_scratch(ctx, chunk))) return rc;
    for (int f0 = 0; f0 < n_frames; f0 += chunk) {
        const int nf = std::min(chunk, n_frames - f0);
        long long max_points = 0;
        for (int f = f0; f < f0 + nf; ++f) {
            const long long n = h_offsets[f + 1] - h_offsets[f];
            if (n < 0) return fail(ctx, GS_EINVAL, "offsets must be non-decreasing (frame %d)", f);
            max_points = std::max(max_points, n);
        }
        const size_t o = static_cast<size_t>(f0) * P;
        rc = run_wave(ctx, d_points_xyzi, 16, ctx->d_offsets + f0, nf, max_points, d_range_out ? d_range_out + o : nullptr,
                      d_angle_out ? d_angle_out + o : nullptr, d_no_ground_out ? d_no_ground_out + o : nullptr,
                      d_xyz4_out ? d_xyz4_out + 4 * o : nullptr, d_counts_out ? d_counts_out + f0 : nullptr);
        if (rc) return rc;
    }
    return GS_OK;
}

int gs_label_components_batch_device(gs_ctx* ctx, const uint16_t* d_images, const uint16_t* d_angles, int n_frames, double beta,
                                     int32_t* d_labels_out, uint32_t* d_n_components_out) {
    if (!ctx || n_frames < 0 || (n_frames && (!d_images || !d_angles || !d_labels_out))) return fail(ctx, GS_EINVAL, "gs_label_components_batch_device: bad arguments");
    if (n_frames == 0) return GS_OK;
    cudaSetDevice(ctx->device);
    launch_components_batch(launch_of(ctx), ctx->kp, n_frames, d_images, d_angles, edge_cutoff(beta), d_labels_out, d_n_components_out, nullptr);
    CU(cudaGetLastError());
    return GS_OK;
}

int gs_process_batch_device_labels(gs_ctx* ctx, const float* d_points_xyzi, const int64_t* h_offsets, int n_frames, uint16_t* d_range_out,
                                   uint16_t* d_angle_out, uint16_t* d_no_ground_out, float* d_xyz4_out, uint32_t* d_counts_out,
                                   int32_t* d_labels_out, uint32_t* d_n_components_out, double beta) {
    if (!d_labels_out)
        return gs_process_batch_device(ctx, d_points_xyzi, h_offsets, n_frames, d_range_out, d_angle_out, d_no_ground_out, d_xyz4_out, d_counts_out);
    if (!ctx || !h_offsets || n_frames < 0 || (n_frames && !d_points_xyzi)) return fail(ctx, GS_EINVAL, "gs_process_batch_device_labels: bad arguments");
    if (n_frames == 0) return GS_OK;
    for (int f = 0; f < n_frames; ++f)
        if (h_offsets[f + 1] < h_offsets[f]) return fail(ctx, GS_EINVAL, "offsets must be non-decreasing (frame %d)", f);
    cudaSetDevice(ctx->device);
    const size_t P = ctx->kp.P;
    // Waves: the CCL kernels read the no-ground and angle images of the wave the fused kernel has just written; a wave is sized
    // so that those two images stay in L2 (126 MB) between the two, i.e. the labels cost their own 4 P bytes per frame and
    // little else.  Images the caller does not ask for live in wave-sized scratch.
    const size_t l2_budget = static_cast<size_t>(80) << 20;
    int wave = static_cast<int>(std::max<size_t>(8, std::min<size_t>(static_cast<size_t>(n_frames), l2_budget / (P * 4))));
    if (fused_supported(ctx->fused)) {  // whole rounds of the persistent grid: no cluster idles through the last frames of a wave
        const int grid = ctx->fused.max_clusters + ctx->fused.soft_clusters;
        if (wave > grid) wave -= wave % grid;
    }
    const bool need_res = d_no_ground_out == nullptr, need_ang = d_angle_out == nullptr;
    if ((need_res || need_ang) && ctx->lab_frames < wave) {
        CU(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->lab_res);
        cudaFree(ctx->lab_ang);
        ctx->lab_res = ctx->lab_ang = nullptr;
        ctx->lab_frames = 0;
        CU(cudaMalloc(&ctx->lab_res, static_cast<size_t>(wave) * P * 2));
        CU(cudaMalloc(&ctx->lab_ang, static_cast<size_t>(wave) * P * 2));
        ctx->lab_frames = wave;
    }
    int rc;
    if ((rc = ensure_offsets(ctx, static_cast<size_t>(n_frames) + 1))) return rc;
    CU(cudaMemcpyAsync(ctx->d_offsets, h_offsets, (static_cast<size_t>(n_frames) + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx->stream));
    const int k_beta = edge_cutoff(beta);
    for (int f0 = 0; f0 < n_frames; f0 += wave) {
        const int nf = std::min(wave, n_frames - f0);
        long long max_points = 0;
        for (int f = f0; f < f0 + nf; ++f) max_points = std::max<long long>(max_points, h_offsets[f + 1] - h_offsets[f]);
        const size_t o = static_cast<size_t>(f0) * P;
        uint16_t* res = need_res ? ctx->lab_res : d_no_ground_out + o;
        uint16_t* ang = need_ang ? ctx->lab_ang : d_angle_out + o;
        rc = run_wave(ctx, d_points_xyzi, 16, ctx->d_offsets + f0, nf, max_points, d_range_out ? d_range_out + o : nullptr, ang, res,
                      d_xyz4_out ? d_xyz4_out + 4 * o : nullptr, d_counts_out ? d_counts_out + f0 : nullptr);
        if (rc) return rc;
        launch_components_batch(launch_of(ctx), ctx->kp, nf, res, ang, k_beta, d_labels_out + o, d_n_components_out ? d_n_components_out + f0 : nullptr, nullptr);
        CU(cudaGetLastError());
    }
    return GS_OK;
}

int gs_process_batch_host_strided(gs_ctx* ctx, const void* h_points, size_t stride_bytes, const int64_t* h_offsets, int n_frames,
                                  uint16_t* h_no_ground_out, float* h_xyz4_out, uint64_t xyz4_capacity, int64_t* h_out_offsets) {
    if (!ctx || !h_offsets || n_frames < 0 || (n_frames && !h_points)) return fail(ctx, GS_EINVAL, "gs_process_batch_host: bad arguments");
    if (stride_bytes < 12 || (stride_bytes & 3)) return fail(ctx, GS_EINVAL, "gs_process_batch_host: records are x, y, z floats plus optional payload, %zu bytes is not a record size", stride_bytes);
    if (h_xyz4_out && !h_out_offsets) return fail(ctx, GS_EINVAL, "gs_process_batch_host: h_out_offsets is required with h_xyz4_out");
    if (n_frames == 0) return GS_OK;
    // every offset is checked before anything is copied: a bad table must not leave DMA in flight on the caller's buffers
    if (h_offsets[0] < 0) return fail(ctx, GS_EINVAL, "offsets must be non-negative");
    long long batch_max = 0;
    for (int f = 0; f < n_frames; ++f) {
        if (h_offsets[f + 1] < h_offsets[f]) return fail(ctx, GS_EINVAL, "offsets must be non-decreasing (frame %d)", f);
        batch_max = std::max<long long>(batch_max, h_offsets[f + 1] - h_offsets[f]);
    }
    cudaSetDevice(ctx->device);
    const size_t P = ctx->kp.P;
    const bool want_cloud = h_xyz4_out != nullptr;
    const bool repack = stride_bytes != 16;  // e.g. 12-byte XYZ records: a quarter less PCIe traffic than XYZI; repacked to 16 bytes on the device
    int rc;
    // Wave size.  The copies dominate (PCIe moves a 64 x 2048 frame in ~35 us, the GPU computes it in ~3 us), so waves are kept
    // small: the first H2D and the last D2H overlap nothing, and that fill / drain time is proportional to the wave.  Two frames
    // per co-resident cluster of the fused kernel keep its persistent grid busy.  Large images are bounded by a byte budget
    // per slot (outputs + worst-case cloud slabs) instead.
    const bool fused_ok = fused_supported(ctx->fused) && batch_max <= fused_max_points();
    const int auto_wave = fused_ok ? std::max(2 * (ctx->fused.max_clusters + ctx->fused.soft_clusters), 16) : 148;
    const size_t per_frame = P * 2 + (want_cloud ? P * 32 : 0) + static_cast<size_t>(batch_max) * (repack ? stride_bytes + 16 : 16);
    const int budget_wave = static_cast<int>(std::max<size_t>(1, (static_cast<size_t>(1) << 30) / std::max<size_t>(per_frame, 1)));
    const int wave = std::max(1, std::min(std::min(n_frames, budget_wave), ctx->chunk_frames > 0 ? ctx->chunk_frames : auto_wave));
    if (ctx->slot_frames < wave || (want_cloud && !ctx->d_cloud[0])) {
        CU(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k < 2; ++k) {
            cudaFree(ctx->d_res[k]); cudaFree(ctx->d_cloud[k]); cudaFree(ctx->d_counts[k]);
            cudaFree(ctx->d_packed[k]); cudaFree(ctx->d_prefix[k]);
            ctx->d_packed[k] = nullptr; ctx->d_prefix[k] = nullptr;
            if (ctx->h_counts[k]) cudaFreeHost(ctx->h_counts[k]);
            ctx->d_res[k] = nullptr; ctx->d_cloud[k] = nullptr; ctx->d_counts[k] = nullptr; ctx->h_counts[k] = nullptr;
        }
        const int nw = std::max(wave, ctx->slot_frames);
        ctx->slot_frames = 0;
        for (int k = 0; k < 2; ++k) {
            CU(cudaMalloc(&ctx->d_res[k], static_cast<size_t>(nw) * P * 2));
            if (want_cloud) {  // the cloud slabs are only allocated for callers that ask for clouds
                CU(cudaMalloc(&ctx->d_cloud[k], static_cast<size_t>(nw) * P * 16));
                CU(cudaMalloc(&ctx->d_packed[k], static_cast<size_t>(nw) * P * 16));
                CU(cudaMalloc(&ctx->d_prefix[k], (static_cast<size_t>(nw) + 1) * sizeof(unsigned long long)));
            }
            CU(cudaMalloc(&ctx->d_counts[k], static_cast<size_t>(nw) * 4));
            CU(cudaMallocHost(&ctx->h_counts[k], static_cast<size_t>(nw) * 4));
        }
        ctx->slot_frames = nw;
    }
    if ((rc = ensure_offsets(ctx, static_cast<size_t>(n_frames) + 1))) return rc;
    CU(cudaMemcpyAsync(ctx->d_offsets, h_offsets, (static_cast<size_t>(n_frames) + 1) * sizeof(long long), cudaMemcpyHostToDevice,
                       ctx->stream));
    // run_wave allocates the staged scratch on demand when a wave cannot take the fused kernel
    const int n_waves = (n_frames + wave - 1) / wave;
    int64_t out_cursor = 0;
    if (h_out_offsets) h_out_offsets[0] = 0;
    const unsigned char* h_bytes = static_cast<const unsigned char*>(h_points);
    // software pipeline over waves: H2D(w+1) overlaps compute(w) overlaps D2H(w-1).
    // Slot layout of d_in[k]: [packed float4 records of the wave][raw records as they came, when they are not 16 bytes wide]
    auto issue_in = [&](int w) -> int {
        const int k = w & 1;
        const int f0 = w * wave, nf = std::min(wave, n_frames - f0);
        const long long p0 = h_offsets[f0], p1 = h_offsets[f0 + nf];
        const size_t n = static_cast<size_t>(p1 - p0);
        const size_t packed = (n * 16 + 255) & ~static_cast<size_t>(255);
        const size_t bytes = repack ? packed + n * stride_bytes : n * 16;
        if (ctx->d_in_cap[k] < bytes) {
            CU(cudaStreamSynchronize(ctx->stream));  // the slot may still be read by an earlier wave
            CU(cudaStreamSynchronize(ctx->copy_in));
            cudaFree(ctx->d_in[k]);
            ctx->d_in[k] = nullptr;
            ctx->d_in_cap[k] = 0;
            CU(cudaMalloc(&ctx->d_in[k], bytes + bytes / 8 + 256));
            ctx->d_in_cap[k] = bytes + bytes / 8 + 256;
        }
        if (w >= 2) CU(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_done[k], 0));  // compute of wave w-2 released the slot
        unsigned char* dst = reinterpret_cast<unsigned char*>(ctx->d_in[k]) + (repack ? packed : 0);
        if (n) CU(cudaMemcpyAsync(dst, h_bytes + static_cast<size_t>(p0) * stride_bytes, n * stride_bytes, cudaMemcpyHostToDevice, ctx->copy_in));
        CU(cudaEventRecord(ctx->ev_in[k], ctx->copy_in));
        return GS_OK;
    };
    auto issue_compute = [&](int w) -> int {
        const int k = w & 1;
        const int f0 = w * wave, nf = std::min(wave, n_frames - f0);
        long long max_points = 0;
        for (int f = f0; f < f0 + nf; ++f) max_points = std::max<long long>(max_points, h_offsets[f + 1] - h_offsets[f]);
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[k], 0));
        if (w >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[k], 0));  // D2H of wave w-2 drained the slot's outputs
        if (repack) {
            const size_t n = static_cast<size_t>(h_offsets[f0 + nf] - h_offsets[f0]);
            const size_t packed = (n * 16 + 255) & ~static_cast<size_t>(255);
            launch_unpack_points(launch_of(ctx), reinterpret_cast<unsigned char*>(ctx->d_in[k]) + packed, n, static_cast<unsigned>(stride_bytes), 0, 4, 8,
                                 stride_bytes >= 16 ? 12 : -1, 0, ctx->d_in[k]);
        }
        // offsets inside the slot are relative to the wave's first point: pass a shifted base pointer instead
        const float* base = ctx->d_in[k] - 4 * h_offsets[f0];
        const int rc2 = run_wave(ctx, base, 16, ctx->d_offsets + f0, nf, max_points, nullptr, nullptr, ctx->d_res[k], want_cloud ? ctx->d_cloud[k] : nullptr,
                                 ctx->d_counts[k]);
        if (rc2) return rc2;
        if (want_cloud)
            launch_pack_clouds(launch_of(ctx), reinterpret_cast<const float4*>(ctx->d_cloud[k]), ctx->d_counts[k], nf, P, ctx->d_prefix[k],
                               reinterpret_cast<float4*>(ctx->d_packed[k]));
        CU(cudaMemcpyAsync(ctx->h_counts[k], ctx->d_counts[k], static_cast<size_t>(nf) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaEventRecord(ctx->ev_done[k], ctx->stream));
        return GS_OK;
    };
    auto issue_out = [&](int w) -> int {
        const int k = w & 1;
        const int f0 = w * wave, nf = std::min(wave, n_frames - f0);
        CU(cudaEventSynchronize(ctx->ev_done[k]));  // counts are on the host now
        // the capacity is checked for the whole wave before its copy is issued
        const int64_t wave_first = out_cursor;
        int64_t cursor = out_cursor;
        for (int f = 0; f < nf; ++f) {
            cursor += ctx->h_counts[k][f];
            if (want_cloud && static_cast<uint64_t>(cursor) > xyz4_capacity) return fail(ctx, GS_EINVAL, "xyz4 output capacity exceeded at frame %d", f0 + f);
        }
        CU(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_done[k], 0));
        if (h_no_ground_out)
            CU(cudaMemcpyAsync(h_no_ground_out + static_cast<size_t>(f0) * P, ctx->d_res[k], static_cast<size_t>(nf) * P * 2,
                               cudaMemcpyDeviceToHost, ctx->copy_out));
        for (int f = 0; f < nf; ++f) {
            out_cursor += ctx->h_counts[k][f];
            if (h_out_offsets) h_out_offsets[f0 + f + 1] = out_cursor;
        }
        if (want_cloud && out_cursor > wave_first)  // the wave's clouds, packed on the device, in one copy
            CU(cudaMemcpyAsync(h_xyz4_out + 4 * wave_first, ctx->d_packed[k], static_cast<size_t>(out_cursor - wave_first) * 16, cudaMemcpyDeviceToHost,
                               ctx->copy_out));
        CU(cudaEventRecord(ctx->ev_out[k], ctx->copy_out));
        return GS_OK;
    };
    // one exit: whatever happened, no copy into or out of the caller's buffers and no kernel on the slots is still in flight on return
    auto pipeline = [&]() -> int {
        int r;
        if ((r = issue_in(0))) return r;
        for (int w = 0; w < n_waves; ++w) {
            if (w + 1 < n_waves && (r = issue_in(w + 1))) return r;
            if ((r = issue_compute(w))) return r;
            if (w >= 1 && (r = issue_out(w - 1))) return r;
        }
        return issue_out(n_waves - 1);
    };
    rc = pipeline();
    const cudaError_t e1 = cudaStreamSynchronize(ctx->copy_in), e2 = cudaStreamSynchronize(ctx->stream), e3 = cudaStreamSynchronize(ctx->copy_out);
    if (rc) return rc;
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
        const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        return fail(ctx, GS_ECUDA, "gs_process_batch_host: %s", cudaGetErrorString(e));
    }
    return GS_OK;
}

int gs_process_batch_host(gs_ctx* ctx, const float* h_points_xyzi, const int64_t* h_offsets, int n_frames, uint16_t* h_no_ground_out,
                          float* h_xyz4_out, uint64_t xyz4_capacity, int64_t* h_out_offsets) {
    return gs_process_batch_host_strided(ctx, h_points_xyzi, 16, h_offsets, n_frames, h_no_ground_out, h_xyz4_out, xyz4_capacity, h_out_offsets);
}

}  // extern "C"
#pragma GCC visibility pop
