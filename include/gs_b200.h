/*
 * gs_b200.h — C ABI of the B200 (sm_100a) ground-segmentation hot path.
 *
 * This is the drop-in boundary for ONE path of matigolias/GroundSegmentationROSNode:
 *     point cloud -> range image -> angle image -> ground mask -> point cloud
 * (reference call stack: src/alfa_gs.cpp:175-244, the hw==0 branch of
 * Alfa_GS::process_pointcloud; two-node packaging: src/cloud2range_node.cpp:381-480 and
 * src/range2cloud_node.cpp:36-99).
 *
 * The reference has no FFI for this path: its only accelerator seam is the `hw` switch in
 * Alfa_GS::process_pointcloud (src/alfa_gs.cpp:97,175) that routes stages to an FPGA through
 * AlfaNode's store/read helpers (lib/alfa_node.h:102-149).  The entry points below are what a
 * GPU back end bound at that seam needs; each cites the reference member it replaces.  Plain
 * pointers and sizes only — no C++/torch/OpenCV/PCL types cross this boundary.
 *
 * Conventions
 *   - images are row-major, row 0 = highest beam (src/alfa_gs.cpp:275); u16 range pixels,
 *     0 = empty, otherwise 1 + 65534 * (range - min_range) / (max_range - min_range);
 *   - a point record is `stride` bytes with float x,y,z at byte offsets 0,4,8
 *     (pcl::PointXYZI = 16 B, pcl::PointXYZRGB = 32 B);
 *   - every function returns GS_OK or an error code and never throws / aborts;
 *     gs_last_error() gives the message of the last failure on that context;
 *   - a context is used by one host thread at a time; one context per GPU; contexts are
 *     independent (the reference runs the path on one ros::spin thread, src/alfa_node.cpp:387).
 *   - there is NO CPU fallback: without a CUDA device gs_create() fails with GS_ECUDA.
 */
#ifndef GS_B200_H
#define GS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------------------------- */
#define GS_OK 0
#define GS_EINVAL 1 /* bad argument / window size / shape (reference: std::logic_error, ROS_ASSERT) */
#define GS_ECUDA 2  /* CUDA runtime failure, message in gs_last_error()                           */
#define GS_ENOMEM 3
#define GS_ESTATE 4

/* ---- policies ------------------------------------------------------- */
/* which packaging of the pipeline the formulas follow */
#define GS_VARIANT_ALFA_GS 0     /* src/alfa_gs.cpp (the executable CMake builds)              */
#define GS_VARIANT_CLOUD2RANGE 1 /* src/cloud2range_node.cpp + src/range2cloud_node.cpp        */
/* which return owns a pixel when several points project to it */
#define GS_PIXEL_LAST 0    /* reference: plain overwrite in input order (alfa_gs.cpp:291)      */
#define GS_PIXEL_NEAREST 1 /* north-star variant: smallest range, ties -> lowest index         */
/* what happens to col == n_cols (alfa_gs.cpp:277 leaves it un-wrapped) */
#define GS_COL_FLAT_ALIAS 0 /* reference: lands on flat index row*W+W = (row+1, 0); dropped past the end */
#define GS_COL_WRAP 1       /* intended behaviour: col % n_cols                                */
/* angle-image smoothing before the ground search */
#define GS_SMOOTH_NONE 0 /* alfa_gs.cpp:212-220: the moving average result is discarded        */
#define GS_SMOOTH_SG 1   /* cloud2range_node.cpp:470: Savitzky-Golay, window = window_size     */

/*
 * Parameters — the constants of Alfa_GS::Alfa_GS (src/alfa_gs.cpp:16-38) and
 * Cloud2RangeNode::Cloud2RangeNode (src/cloud2range_node.cpp:281-300).
 * Derived values (n_cols, d_azimuth, d_altitude; alfa_gs.cpp:48-50) are recomputed by
 * gs_create()/gs_set_params() — unlike Alfa_GS::UpdateSegmentationSettings
 * (alfa_gs.cpp:1122-1136), which forgets to.
 */
typedef struct gs_params {
    int32_t n_beams;     /* H, image rows                                   */
    int32_t rpm;         /* rotations per minute                            */
    int32_t sample_freq; /* firings per second; n_cols = sample_freq*60/rpm */
    double min_angle;    /* rad, lowest beam altitude                       */
    double max_angle;    /* rad, highest beam altitude                      */
    double min_range;    /* m                                               */
    double max_range;    /* m                                               */
    double ground_angle_threshold; /* rad (alfa_gs.cpp:18)                  */
    double start_angle_threshold;  /* rad (alfa_gs.cpp:20)                  */
    int32_t window_size;  /* 5,7,9,11: SG window and dilation kernel source (alfa_gs.cpp:16) */
    int32_t erase_ground; /* 1: output keeps non-ground; 0: keeps ground (alfa_gs.cpp:38,753) */
    int32_t variant;      /* GS_VARIANT_*      */
    int32_t pixel_policy; /* GS_PIXEL_*        */
    int32_t col_policy;   /* GS_COL_*          */
    int32_t smoothing;    /* GS_SMOOTH_*       */
    int32_t repair_gaps;  /* 0 (default) / 1: typed RepairGaps before the angle image (cloud2range_node.cpp:468) */
    int32_t reserved;
} gs_params;

/* values the reference derives in its constructors, plus the integer cut-offs of appendix A.3/A.4 */
typedef struct gs_derived {
    int32_t n_rows, n_cols;
    double d_azimuth, d_altitude;
    int32_t k_ground;     /* edge  <=> |angle difference| <= k_ground (label.cpp:90,99)       */
    int32_t s_start;      /* seed accepted <=> angle <= s_start       (alfa_gs.cpp:733)       */
    int32_t n_dilate;     /* number of vertical taps of the effective structuring element      */
    int32_t dilate_off[2];/* dilated(r,c) = OR_k L(r + dilate_off[k], c) (alfa_gs.cpp:747-750) */
} gs_derived;

/* sensor presets taken from the reference's own parameter blocks */
#define GS_PRESET_ALFA_GS_DEFAULT 0 /* alfa_gs.cpp:16-38: HDL-64E, 64 x 1800                  */
#define GS_PRESET_VLP16 1           /* cloud2range_node.cpp:66-86: 16 x 1800                  */
#define GS_PRESET_HDL64E_2048 2     /* alfa_gs.cpp:16-38 with sample_freq 20480: 64 x 2048    */
#define GS_PRESET_OS1_128 3         /* cloud2range_node.cpp:281-300, beams x2: 128 x 2048     */
#define GS_PRESET_CLOUD2RANGE_DEFAULT 4 /* cloud2range_node.cpp:281-300: OS1-64, 64 x 2000, cloud2range variant, SG-7 */
#define GS_PRESET_HDL32 5           /* cloud2range_node.cpp:122-135: 32 x 1800                */

typedef struct gs_ctx gs_ctx;

/* counters filled by the last call on a context */
typedef struct gs_stats {
    uint64_t kernel_launches;   /* CUDA kernels launched by this context since creation           */
    uint64_t near_edge_project; /* points whose altitude/azimuth lies within 1e-5 rad of a cell edge */
    uint64_t near_edge_angle;   /* pixels whose inclination angle lies within 1e-5 rad of a centi-radian edge */
    uint64_t dropped_points;    /* points skipped by range/altitude gates or falling outside the image */
    uint64_t slow_path_points;  /* points / pixels whose fast (float) evaluation fell inside its certified error margin of a
                                   quantisation edge and were re-evaluated with the reference's exact arithmetic */
    uint64_t fast_path_mismatches; /* with GS_OPT_COLLECT_STATS the fused kernel evaluates BOTH paths for every point and pixel and
                                      counts accepted fast results that differ from the exact ones: must stay 0 */
    uint64_t remote_points;     /* points whose pixel lives in another CTA of the frame's cluster (distributed shared memory path) */
} gs_stats;

/* ---- parameters --------------------------------------------------------------------------- */
int gs_abi_version(void);
int gs_default_params(int preset, gs_params* out);
/* alfa_gs.cpp:48-50 (+ thresholds, dilation taps).  GS_EINVAL on a bad window / shape. */
int gs_derive(const gs_params* p, gs_derived* out);

/* ---- context ------------------------------------------------------------------------------ */
/* replaces Alfa_GS::Alfa_GS (alfa_gs.cpp:12-93): parameters, derived values, LUTs, device scratch */
int gs_create(const gs_params* p, int device, gs_ctx** out);
void gs_destroy(gs_ctx* ctx);
/* replaces Alfa_GS::UpdateSegmentationSettings (alfa_gs.cpp:1122-1136) */
int gs_set_params(gs_ctx* ctx, const gs_params* p);
int gs_get_params(const gs_ctx* ctx, gs_params* out);
int gs_get_derived(const gs_ctx* ctx, gs_derived* out);
/* all work of the context is enqueued on this cudaStream_t (0 = the context's own stream); the old stream is drained first */
int gs_set_stream(gs_ctx* ctx, void* cuda_stream);
/* frames processed per wave by the batch calls (scratch = chunk * ~20 B/pixel, kept L2 resident) */
int gs_set_chunk_frames(gs_ctx* ctx, int frames);
/* options: GS_OPT_COLLECT_STATS (default 0) makes the kernels fill the near-edge / dropped counters of gs_stats;
 * GS_OPT_FORCE_STAGED (default 0) routes the frame/batch calls through the per-stage kernels instead of the fused one
 * (both are CUDA paths; used by the parity tests to check one against the other). */
#define GS_OPT_COLLECT_STATS 1
#define GS_OPT_FORCE_STAGED 2
#define GS_OPT_PHASE_TIMING 3 /* fused kernel accumulates clock64() cycles per internal phase (gs_get_phase_cycles) */
#define GS_OPT_CAS_UNIONS 4   /* fused kernel: take the compare-and-swap union path (the one used by frames whose link list
                               * does not fit in shared memory) for every frame; a test aid */
#define GS_OPT_NO_LONE_PLAN 5 /* fused kernel: a call that brings ONE frame normally spreads it over the widest cluster that fits
                               * (lowest latency); 1 = use the batch plan's cluster for single frames too.  A test aid. */
int gs_set_option(gs_ctx* ctx, int option, int value);
int gs_get_stats(gs_ctx* ctx, gs_stats* out);
int gs_reset_stats(gs_ctx* ctx);
/* fused kernel, with GS_OPT_PHASE_TIMING: SM clock cycles per internal phase summed over all CTAs since the last reset
 * (16 entries; the phase names are in DESIGN.md).  A profiling aid, not part of the reference's surface. */
int gs_get_phase_cycles(gs_ctx* ctx, uint64_t* out16);
const char* gs_last_error(const gs_ctx* ctx);

/* ---- stage API: one frame, HOST buffers (copies included) ---------------------------------- */
/* Alfa_GS::CreateRangeImage (alfa_gs.cpp:247-298) / Cloud2RangeNode::CloudCb projection
 * (cloud2range_node.cpp:396-455), chosen by params.variant.  range_out: n_rows*n_cols u16. */
int gs_project(gs_ctx* ctx, const void* points, size_t n_points, size_t stride_bytes,
               uint16_t* range_out);
/* Alfa_GS::CreateAngleImg (alfa_gs.cpp:335-386) */
int gs_angle_image(gs_ctx* ctx, const uint16_t* range, uint16_t* angle_out);
/* Alfa_GS::SavitskyGolaySmoothing (alfa_gs.cpp:705-713); window in {5,7,9,11} else GS_EINVAL
 * (reference: std::logic_error, alfa_gs.cpp:642-649) */
int gs_smooth(gs_ctx* ctx, const uint16_t* angle, int window, uint16_t* smoothed_out);
/* typed Alfa_GS::RepairGaps (alfa_gs.cpp:300-333) on encoded u16 depths */
int gs_repair_gaps(gs_ctx* ctx, const uint16_t* range, int step, float depth_threshold,
                   uint16_t* repaired_out);
/* Alfa_GS::EraseBFS (alfa_gs.cpp:715-766) incl. Label::LabelOneComponent (Labeler/label.cpp:48-104),
 * GetUniformKernel + cv::dilate (alfa_gs.cpp:747-750).  labels_out (n_rows*n_cols u8, may be NULL)
 * receives the Labeler's label image before dilation.  Even kernel_size-2 -> GS_EINVAL
 * (reference: std::logic_error, alfa_gs.cpp:631-633). */
int gs_erase_ground(gs_ctx* ctx, const uint16_t* range, const uint16_t* angle,
                    double ground_angle_threshold, double start_angle_threshold, int kernel_size,
                    uint16_t* res_out, uint8_t* labels_out);
/* Label::LabelOneComponent (Labeler/label.cpp:48-104) on an existing label image (in/out, u8):
 * labels the 4-connected component of (row,col) under the angle threshold; already labelled
 * pixels are barriers, empty pixels are labelled but not expanded. */
int gs_label_one_component(gs_ctx* ctx, int label, int row, int col, float angle_threshold,
                           const uint16_t* range, const uint16_t* angle, uint8_t* label_image);
/* The Labeler driven as an all-components CCL (north-star kernel 4): components of the non-zero
 * pixels of `image` under the angle threshold `beta` (rad).  labels_out: int32 per pixel = smallest
 * flat index of the pixel's component (canonical relabelling), -1 for empty pixels. */
int gs_label_components(gs_ctx* ctx, const uint16_t* image, const uint16_t* angle, double beta,
                        int32_t* labels_out, int32_t* n_components_out);
/* Alfa_GS::CameraCb (alfa_gs.cpp:997-1074) / Range2CloudNode::CameraCb (range2cloud_node.cpp:36-99,
 * double range) chosen by params.variant.  xyz4_out: up to n_rows*n_cols records {x,y,z,1.0f},
 * row-major order of the surviving pixels. */
int gs_backproject(gs_ctx* ctx, const uint16_t* image, float* xyz4_out, uint32_t* n_out);
/* The same with the decoding parameters taken from the sensor_msgs/CameraInfo that travels with the image, as
 * Range2CloudNode::CameraCb does (range2cloud_node.cpp:51-58: min_angle = K[0], min/max range = K[2] / K[3], d_azimuth = K[4],
 * d_altitude = K[5]; range kept in double, :76) and Alfa_GS::CameraCb with its cinfo argument (alfa_gs.cpp:1010-1015; range
 * rounded to float, :1027,1045: range_as_double = 0).  K: at least 6 values.  height x width must be the context's shape
 * (GS_EINVAL otherwise): the context only lends its device buffers, none of its own derived values are used. */
int gs_backproject_camera_info(gs_ctx* ctx, const uint16_t* image, int32_t height, int32_t width, const double* K,
                               int range_as_double, float* xyz4_out, uint32_t* n_out);
/* Alfa_GS::process_pointcloud, hw==0 branch (alfa_gs.cpp:175-244) / Cloud2RangeNode::CloudCb +
 * Range2CloudNode::CameraCb.  Any output pointer may be NULL. */
int gs_process_frame(gs_ctx* ctx, const void* points, size_t n_points, size_t stride_bytes,
                     uint16_t* range_out, uint16_t* angle_out, uint16_t* no_ground_out,
                     float* xyz4_out, uint32_t* n_out);

/* ---- wire formats either side of the path ----------------------------------- */
/* sensor_msgs/PointCloud2 -> packed {x, y, z, intensity} float records: what pcl::fromROSMsg does for the node
 * (alfa_node.cpp:311).  data: n_points records of point_step bytes with float32 fields at the given byte offsets
 * (off_intensity < 0: none, 0 is stored).  HOST buffers. */
int gs_unpack_points(gs_ctx* ctx, const void* data, size_t n_points, uint32_t point_step, uint32_t off_x,
                     uint32_t off_y, uint32_t off_z, int32_t off_intensity, int is_bigendian, float* xyzi_out);
/* the whole path on a PointCloud2-style buffer (fields unpacked on the device, then the fused kernel) */
int gs_process_cloud2(gs_ctx* ctx, const void* data, size_t n_points, uint32_t point_step, uint32_t off_x,
                      uint32_t off_y, uint32_t off_z, int is_bigendian, uint16_t* range_out, uint16_t* angle_out,
                      uint16_t* no_ground_out, float* xyz4_out, uint32_t* n_out);
/* {x, y, z, 1} records -> 32-byte pcl::PointXYZRGB records (x y z 1.0 | b g r a = 0 0 0 255 | padding), the cloud
 * Alfa_GS::CameraCb builds for pcl::toROSMsg (alfa_gs.cpp:1055-1058, alfa_node.cpp:35-41).  The {x, y, z, 1} records
 * themselves are already pcl::PointXYZ (range2cloud_node.cpp:89). */
int gs_pack_cloud_xyzrgb(gs_ctx* ctx, const float* xyz4, uint32_t n_points, void* out32);
/* sensor_msgs/CameraInfo fields the nodes publish with the mono16 image (alfa_gs.cpp:55-63, cloud2range_node.cpp:319-328):
 * K[0..5] = min_angle, max_angle, min_range, max_range, d_azimuth, d_altitude.  The image itself needs no conversion:
 * the u16 row-major buffers of this API are the mono16 payload (step = 2 * width). */
int gs_get_camera_info(const gs_ctx* ctx, int32_t* height, int32_t* width, double* K9);

/* ---- FPGA wire formats and the coloured angle image.  HOST buffers. ----
 * gs_fpga_store_pointcloud = AlfaNode::store_pointcloud_hardware (src/alfa_node.cpp:105-131): one 64-bit word per point, four u16
 * {elevation, azimuth in centi-degrees, range in centimetres, 0}, in the reference's float arithmetic (atan2f / hypotf / sqrtf).
 * gs_fpga_read_range_image = AlfaNode::read_hardware_pointcloud (:157-194): rows * cols / 4 words of four u16 pixels, column by
 * column (64 rows each), into the 64 x 1800 image the reference hard-codes (:162); more pixels than that -> GS_EINVAL.
 * gs_fpga_read_angle_image = AlfaNode::read_hardware_filtered_angle_image (:196-271): 10-bit fields, six per word and four in every
 * 43rd word, column by column into a rows x cols image; rows must be 64 (the reference's row counter, :219,240).  The input words
 * are left untouched (the reference shifts them while reading).
 * gs_colored_angle_image = Alfa_GS::CreateColoredAngleImage (src/alfa_gs.cpp:768-829): rows x cols x 3 bytes; every angle is
 * scaled by 255 / (largest angle seen so far in raster order) before it is coloured (:781-793) — kept literally. */
int gs_fpga_store_pointcloud(gs_ctx* ctx, const void* points, size_t n_points, size_t stride_bytes, uint64_t* words_out);
int gs_fpga_read_range_image(gs_ctx* ctx, const uint64_t* words, uint32_t rows, uint32_t cols, uint16_t* image64x1800_out);
int gs_fpga_read_angle_image(gs_ctx* ctx, const uint64_t* words, uint32_t rows, uint32_t cols, uint16_t* image_out);
int gs_colored_angle_image(gs_ctx* ctx, const uint16_t* angle, int32_t rows, int32_t cols, uint8_t* rgb_out);

/* ---- evaluation against labelled data ----
 * gs_create_labeled_range_image = Alfa_GS::CreateLabeledRangeImage (src/alfa_gs.cpp:945-995): points are 32-byte
 * pcl::PointXYZRGB records (x y z pad | b g r a | pad); the image is 1 everywhere and 0 where a kept point lands while the
 * reference's gnd_flag is up — the flag is raised by every point coloured (75,0,75), (255,0,255) or (255,150,255) and
 * lowered only at the end of a kept point's iteration, so it sticks across points skipped by the gates (kept literally).
 * gs_evaluate = the per-frame part of Alfa_GS::CheckNumberOfDetectedRIdGnd (src/alfa_gs.cpp:831-889): the counters of
 * :838-870 over (original range image, no-ground image, labelled image) and IoU_g / Recall_g / Recall_mo in the
 * reference's float arithmetic (:885-887; NaN when a denominator is 0, as there).  The reference's integer percentages
 * (:873-877) divide by counters that may be 0 (and true_detected_gnd_cntr is never incremented) and its running means
 * over frames are logging state: neither is reproduced. */
typedef struct gs_eval {
    uint64_t points, labeled_gnd, labeled_obst, seg_gnd;
    uint64_t true_positive, false_positive, true_negative, false_negative, true_detected_obst;
    float iou_g, recall_g, recall_mo;
} gs_eval;
int gs_create_labeled_range_image(gs_ctx* ctx, const void* points_xyzrgb32, size_t n_points, uint16_t* labeled_out);
int gs_evaluate(gs_ctx* ctx, const uint16_t* og_range, const uint16_t* seg_range, const uint16_t* labeled_range, gs_eval* out);

/* ---- batch API: frames are independent (alfa_gs.cpp:196-239 holds no cross-frame state) ---- */
/* DEVICE buffers.  points: packed 16-byte XYZI records of all frames; offsets (HOST, n_frames+1
 * entries, in points) delimit frames.  no_ground_out: n_frames * P u16;  xyz4_out: n_frames * P
 * float4 slabs (frame f starts at f*P records);  counts_out: n_frames u32 (device).
 * range_out / angle_out may be NULL.  Asynchronous on the context's stream. */
int gs_process_batch_device(gs_ctx* ctx, const float* d_points_xyzi, const int64_t* h_offsets,
                            int n_frames, uint16_t* d_range_out, uint16_t* d_angle_out,
                            uint16_t* d_no_ground_out, float* d_xyz4_out, uint32_t* d_counts_out);
/* The same with the Labeler driven as an all-components CCL on every frame's no-ground image (north-star kernel 4; BASELINE
 * configs C3 "full pipeline including Labeler CCL" and C5): d_labels_out, n_frames * P int32 = smallest flat index of the pixel's
 * component, -1 for erased / empty pixels (canonical min-index relabelling); components join 4-neighbours (column
 * wrap, Label::WrapCols) whose angles (the image EraseBFS saw: smoothed when smoothing is on) differ by at most the integer
 * cut-off of `beta` (Labeler/label.cpp:90-99, label.h:41).  d_n_components_out: n_frames u32, may be NULL.  d_labels_out NULL =
 * gs_process_batch_device.  Reference interface: Label::LabelOneComponent, Labeler/label.cpp:48-104. */
int gs_process_batch_device_labels(gs_ctx* ctx, const float* d_points_xyzi, const int64_t* h_offsets,
                                   int n_frames, uint16_t* d_range_out, uint16_t* d_angle_out,
                                   uint16_t* d_no_ground_out, float* d_xyz4_out, uint32_t* d_counts_out,
                                   int32_t* d_labels_out, uint32_t* d_n_components_out, double beta);
/* the CCL alone on device-resident images (n_frames * P u16 each) */
int gs_label_components_batch_device(gs_ctx* ctx, const uint16_t* d_images, const uint16_t* d_angles,
                                     int n_frames, double beta, int32_t* d_labels_out,
                                     uint32_t* d_n_components_out);
/* HOST buffers (pinned recommended): the same pipeline with the host<->device copies done inside, overlapped with
 * compute over waves of frames.  h_no_ground_out: n_frames * P u16 (may be NULL).  h_xyz4_out (may be NULL): the
 * clouds of all frames packed back to back, xyz4_capacity records long; h_out_offsets (n_frames + 1 entries, in
 * records; required with h_xyz4_out, else optional) delimits them.  Synchronous: outputs are complete on return. */
int gs_process_batch_host(gs_ctx* ctx, const float* h_points_xyzi, const int64_t* h_offsets,
                          int n_frames, uint16_t* h_no_ground_out, float* h_xyz4_out,
                          uint64_t xyz4_capacity, int64_t* h_out_offsets);
/* The same for records of stride_bytes (a multiple of 4, >= 12) with x, y, z floats first: 12-byte XYZ records move a quarter
 * less over PCIe than XYZI (the path never reads the intensity, alfa_gs.cpp:259-267), 32-byte pcl::PointXYZRGB records are
 * accepted as they are; anything but 16 bytes is repacked on the device.  All offsets are validated before the first copy is
 * issued (GS_EINVAL), and on every return — success or error — no copy from or into the caller's buffers is still in flight. */
int gs_process_batch_host_strided(gs_ctx* ctx, const void* h_points, size_t stride_bytes, const int64_t* h_offsets,
                                  int n_frames, uint16_t* h_no_ground_out, float* h_xyz4_out,
                                  uint64_t xyz4_capacity, int64_t* h_out_offsets);
int gs_synchronize(gs_ctx* ctx);

/* ---- multi-GPU host driver: frames are independent, so a batch is sharded frame-parallel over the GPUs of one
 * box with no collective.  One host thread + one gs_ctx + its streams per GPU; every worker thread is bound to the CPUs next to
 * its GPU before it creates its context.  The reference runs one frame per callback on one ros::spin thread
 * (src/alfa_node.cpp:298-319,387); N such nodes on N GPUs is what this replaces. */
typedef struct gs_multi gs_multi;
/* devices: CUDA device ordinals (a device may be listed more than once: two shards on one GPU); NULL / 0 = every visible device */
int gs_multi_create(const gs_params* p, const int* devices, int n_devices, gs_multi** out);
void gs_multi_destroy(gs_multi* m);
int gs_multi_n_gpus(const gs_multi* m);
int gs_multi_worker_info(const gs_multi* m, int shard, int* device, int* numa_node, int* cpu_bound);
const char* gs_multi_last_error(const gs_multi* m);
/* page-locked host memory for batch buffers whose pages are first touched by the worker that will copy them (block partition) */
int gs_multi_host_alloc(gs_multi* m, size_t bytes, void** out);
int gs_multi_host_free(gs_multi* m, void* p, size_t bytes);
/* gs_process_batch_host_strided on every GPU at once: shard g of G processes the frames [g * F / G, (g + 1) * F / G).  Shard g
 * packs its clouds from record g * (xyz4_capacity / G) of h_xyz4_out on; h_out_first[f] / h_out_count[f] (n_frames entries each,
 * required with h_xyz4_out) locate frame f's cloud.  Outputs are identical to the single-GPU call's, frame by frame. */
int gs_multi_process_batch_host(gs_multi* m, const void* h_points, size_t stride_bytes, const int64_t* h_offsets,
                                int n_frames, uint16_t* h_no_ground_out, float* h_xyz4_out, uint64_t xyz4_capacity,
                                int64_t* h_out_first, uint32_t* h_out_count);

#ifdef __cplusplus
}
#endif
#endif /* GS_B200_H */
