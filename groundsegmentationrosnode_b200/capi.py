"""ctypes binding of libgs_b200.so (include/gs_b200.h) — the only way Python reaches the CUDA path.

The library is built in-tree by `groundsegmentationrosnode_b200/csrc/build.py` (nvcc, sm_100a).  There is
no fallback: a missing library raises ImportError-like RuntimeError at load time, and a missing GPU makes
`Context()` raise `GsError(GS_ECUDA)`.
"""
import ctypes as C
import os

import numpy as np

from .params import (GS_ECUDA, GS_EINVAL, GS_OK, GsDerived, GsEval, GsParams, GsStats)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GS_B200_LIB") or os.path.join(HERE, "libgs_b200.so")  # GS_B200_LIB: another build of the same ABI (A/B runs)

GS_OPT_COLLECT_STATS = 1
GS_OPT_FORCE_STAGED = 2
GS_OPT_PHASE_TIMING = 3
GS_OPT_CAS_UNIONS = 4
GS_OPT_NO_LONE_PLAN = 5

# every symbol include/gs_b200.h declares (tests check that the library exports all of them)
EXPORTS = [
    "gs_abi_version", "gs_default_params", "gs_derive", "gs_create", "gs_destroy", "gs_set_params", "gs_get_params",
    "gs_get_derived", "gs_set_stream", "gs_set_chunk_frames", "gs_set_option", "gs_get_stats", "gs_reset_stats", "gs_get_phase_cycles",
    "gs_last_error", "gs_project", "gs_angle_image", "gs_smooth", "gs_repair_gaps", "gs_erase_ground",
    "gs_label_one_component", "gs_label_components", "gs_backproject", "gs_process_frame", "gs_process_batch_device",
    "gs_process_batch_host", "gs_process_batch_host_strided", "gs_process_batch_device_labels", "gs_label_components_batch_device", "gs_backproject_camera_info", "gs_synchronize", "gs_unpack_points", "gs_process_cloud2", "gs_pack_cloud_xyzrgb",
    "gs_get_camera_info", "gs_create_labeled_range_image", "gs_evaluate",
    "gs_fpga_store_pointcloud", "gs_fpga_read_range_image", "gs_fpga_read_angle_image", "gs_colored_angle_image",
    "gs_multi_create", "gs_multi_destroy", "gs_multi_n_gpus", "gs_multi_worker_info", "gs_multi_last_error", "gs_multi_host_alloc",
    "gs_multi_host_free", "gs_multi_process_batch_host",
]


class GsError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("gs_b200 error %d: %s" % (code, message))
        self.code = code


_lib = None


def load(build_if_missing=False):
    """dlopen the C-ABI library.  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from .csrc import build as _b

            _b.build()
        else:
            raise RuntimeError("%s is missing: run `python groundsegmentationrosnode_b200/csrc/build.py` "
                               "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u16p, u8p, f32p, i32p, i64p = (C.c_void_p,) * 6
    L.gs_abi_version.restype = C.c_int
    L.gs_default_params.argtypes = [C.c_int, C.POINTER(GsParams)]
    L.gs_derive.argtypes = [C.POINTER(GsParams), C.POINTER(GsDerived)]
    L.gs_create.argtypes = [C.POINTER(GsParams), C.c_int, C.POINTER(C.c_void_p)]
    L.gs_destroy.argtypes = [C.c_void_p]
    L.gs_destroy.restype = None
    L.gs_set_params.argtypes = [C.c_void_p, C.POINTER(GsParams)]
    L.gs_get_params.argtypes = [C.c_void_p, C.POINTER(GsParams)]
    L.gs_get_derived.argtypes = [C.c_void_p, C.POINTER(GsDerived)]
    L.gs_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.gs_set_chunk_frames.argtypes = [C.c_void_p, C.c_int]
    L.gs_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.gs_get_stats.argtypes = [C.c_void_p, C.POINTER(GsStats)]
    L.gs_reset_stats.argtypes = [C.c_void_p]
    L.gs_get_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.gs_last_error.argtypes = [C.c_void_p]
    L.gs_last_error.restype = C.c_char_p
    L.gs_project.argtypes = [C.c_void_p, vp, C.c_size_t, C.c_size_t, u16p]
    L.gs_angle_image.argtypes = [C.c_void_p, u16p, u16p]
    L.gs_smooth.argtypes = [C.c_void_p, u16p, C.c_int, u16p]
    L.gs_repair_gaps.argtypes = [C.c_void_p, u16p, C.c_int, C.c_float, u16p]
    L.gs_erase_ground.argtypes = [C.c_void_p, u16p, u16p, C.c_double, C.c_double, C.c_int, u16p, u8p]
    L.gs_label_one_component.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, u16p, u16p, u8p]
    L.gs_label_components.argtypes = [C.c_void_p, u16p, u16p, C.c_double, i32p, C.POINTER(C.c_int32)]
    L.gs_backproject.argtypes = [C.c_void_p, u16p, f32p, C.POINTER(C.c_uint32)]
    L.gs_process_frame.argtypes = [C.c_void_p, vp, C.c_size_t, C.c_size_t, u16p, u16p, u16p, f32p, C.POINTER(C.c_uint32)]
    L.gs_process_batch_device.argtypes = [C.c_void_p, vp, i64p, C.c_int, vp, vp, vp, vp, vp]
    L.gs_process_batch_host.argtypes = [C.c_void_p, vp, i64p, C.c_int, vp, vp, C.c_uint64, i64p]
    L.gs_process_batch_device_labels.argtypes = [C.c_void_p, vp, i64p, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_double]
    L.gs_label_components_batch_device.argtypes = [C.c_void_p, vp, vp, C.c_int, C.c_double, vp, vp]
    L.gs_process_batch_host_strided.argtypes = [C.c_void_p, vp, C.c_size_t, i64p, C.c_int, vp, vp, C.c_uint64, i64p]
    L.gs_backproject_camera_info.argtypes = [C.c_void_p, u16p, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_int, f32p, C.POINTER(C.c_uint32)]
    L.gs_synchronize.argtypes = [C.c_void_p]
    L.gs_unpack_points.argtypes = [C.c_void_p, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int32, C.c_int, f32p]
    L.gs_process_cloud2.argtypes = [C.c_void_p, vp, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, u16p, u16p, u16p, f32p,
                                    C.POINTER(C.c_uint32)]
    L.gs_pack_cloud_xyzrgb.argtypes = [C.c_void_p, f32p, C.c_uint32, vp]
    L.gs_get_camera_info.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.gs_create_labeled_range_image.argtypes = [C.c_void_p, vp, C.c_size_t, u16p]
    L.gs_evaluate.argtypes = [C.c_void_p, u16p, u16p, u16p, C.POINTER(GsEval)]
    L.gs_fpga_store_pointcloud.argtypes = [C.c_void_p, vp, C.c_size_t, C.c_size_t, vp]
    L.gs_fpga_read_range_image.argtypes = [C.c_void_p, vp, C.c_uint32, C.c_uint32, u16p]
    L.gs_fpga_read_angle_image.argtypes = [C.c_void_p, vp, C.c_uint32, C.c_uint32, u16p]
    L.gs_colored_angle_image.argtypes = [C.c_void_p, u16p, C.c_int32, C.c_int32, vp]
    L.gs_multi_create.argtypes = [C.POINTER(GsParams), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    L.gs_multi_destroy.argtypes = [C.c_void_p]
    L.gs_multi_destroy.restype = None
    L.gs_multi_n_gpus.argtypes = [C.c_void_p]
    L.gs_multi_worker_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.gs_multi_last_error.argtypes = [C.c_void_p]
    L.gs_multi_last_error.restype = C.c_char_p
    L.gs_multi_host_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
    L.gs_multi_host_free.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.gs_multi_process_batch_host.argtypes = [C.c_void_p, vp, C.c_size_t, i64p, C.c_int, vp, vp, C.c_uint64, i64p, vp]
    _lib = L
    return L


def default_params(preset):
    p = GsParams()
    rc = load().gs_default_params(preset, C.byref(p))
    if rc != GS_OK:
        raise GsError(rc, "unknown preset %r" % (preset,))
    return p


def derive(p):
    d = GsDerived()
    rc = load().gs_derive(C.byref(p), C.byref(d))
    if rc != GS_OK:
        raise GsError(rc, "invalid parameters")
    return d


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u16(img, shape):
    a = np.ascontiguousarray(img, dtype=np.uint16)
    if a.shape != tuple(shape):
        raise GsError(GS_EINVAL, "image shape %r != %r" % (a.shape, tuple(shape)))
    return a


class Context:
    """One gs_ctx: parameters, LUTs, device scratch and a stream on one GPU."""

    def __init__(self, params, device=0):
        self._L = load()
        h = C.c_void_p()
        rc = self._L.gs_create(C.byref(params), device, C.byref(h))
        if rc != GS_OK:
            raise GsError(rc, "gs_create failed (%s)" % ("no usable CUDA device" if rc == GS_ECUDA else "invalid parameters"))
        self._h = h
        self.params = self.get_params()
        self.derived = self.get_derived()
        self.shape = (self.derived.n_rows, self.derived.n_cols)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gs_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != GS_OK:
            raise GsError(rc, self._L.gs_last_error(self._h).decode())

    # ---- configuration
    def set_params(self, params):
        self._ck(self._L.gs_set_params(self._h, C.byref(params)))
        self.params = self.get_params()
        self.derived = self.get_derived()
        self.shape = (self.derived.n_rows, self.derived.n_cols)

    def get_params(self):
        p = GsParams()
        self._ck(self._L.gs_get_params(self._h, C.byref(p)))
        return p

    def get_derived(self):
        d = GsDerived()
        self._ck(self._L.gs_get_derived(self._h, C.byref(d)))
        return d

    def set_stream(self, cuda_stream_handle):
        self._ck(self._L.gs_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def set_chunk_frames(self, n):
        self._ck(self._L.gs_set_chunk_frames(self._h, n))

    def set_option(self, opt, value):
        self._ck(self._L.gs_set_option(self._h, opt, int(value)))

    def stats(self):
        s = GsStats()
        self._ck(self._L.gs_get_stats(self._h, C.byref(s)))
        return s

    PHASES = ["clear+deal", "project rounds", "wait: rounds done", "remote records", "angle walk", "wait: angles done",
              "edges+seeds", "wait: edges done", "link scan", "boundary rounds", "forest readout",
              "label+dilate+transpose", "wait: row counts", "images+cloud out", "union pass", "seed flags"]

    def phase_cycles(self):
        out = (C.c_uint64 * 16)()
        self._ck(self._L.gs_get_phase_cycles(self._h, out))
        return dict(zip(self.PHASES, [int(v) for v in out]))

    def reset_stats(self):
        self._ck(self._L.gs_reset_stats(self._h))

    def synchronize(self):
        self._ck(self._L.gs_synchronize(self._h))

    # ---- stage API (host numpy in / out)
    def project(self, points):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.empty(self.shape, np.uint16)
        self._ck(self._L.gs_project(self._h, _ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 16, _ptr(out)))
        return out

    def angle_image(self, range_image):
        out = np.empty(self.shape, np.uint16)
        self._ck(self._L.gs_angle_image(self._h, _ptr(_u16(range_image, self.shape)), _ptr(out)))
        return out

    def smooth(self, angle_image, window):
        out = np.empty(self.shape, np.uint16)
        self._ck(self._L.gs_smooth(self._h, _ptr(_u16(angle_image, self.shape)), window, _ptr(out)))
        return out

    def repair_gaps(self, range_image, step=5, depth_threshold=1.0):
        out = np.empty(self.shape, np.uint16)
        self._ck(self._L.gs_repair_gaps(self._h, _ptr(_u16(range_image, self.shape)), step, depth_threshold, _ptr(out)))
        return out

    def erase_ground(self, range_image, angle_image, ground_thr=None, start_thr=None, kernel_size=None):
        p = self.params
        res = np.empty(self.shape, np.uint16)
        lab = np.empty(self.shape, np.uint8)
        self._ck(self._L.gs_erase_ground(self._h, _ptr(_u16(range_image, self.shape)), _ptr(_u16(angle_image, self.shape)),
                                         p.ground_angle_threshold if ground_thr is None else ground_thr,
                                         p.start_angle_threshold if start_thr is None else start_thr,
                                         p.window_size if kernel_size is None else kernel_size, _ptr(res), _ptr(lab)))
        return res, lab

    def label_one_component(self, label, row, col, thr, range_image, angle_image, label_image):
        lab = np.ascontiguousarray(label_image, dtype=np.uint8).copy()
        self._ck(self._L.gs_label_one_component(self._h, label, row, col, thr, _ptr(_u16(range_image, self.shape)),
                                                _ptr(_u16(angle_image, self.shape)), _ptr(lab)))
        return lab

    def label_components(self, image, angle_image, beta):
        out = np.empty(self.shape, np.int32)
        n = C.c_int32(0)
        self._ck(self._L.gs_label_components(self._h, _ptr(_u16(image, self.shape)), _ptr(_u16(angle_image, self.shape)), beta,
                                             _ptr(out), C.byref(n)))
        return out, n.value

    def backproject(self, image):
        H, W = self.shape
        xyz = np.empty((H * W, 4), np.float32)
        n = C.c_uint32(0)
        self._ck(self._L.gs_backproject(self._h, _ptr(_u16(image, self.shape)), _ptr(xyz), C.byref(n)))
        return xyz[: n.value].copy()

    def process_frame(self, points):
        """-> dict(range, angle, no_ground, xyz)"""
        H, W = self.shape
        pts = np.ascontiguousarray(points, dtype=np.float32)
        rng, ang, res = (np.empty((H, W), np.uint16) for _ in range(3))
        xyz = np.empty((H * W, 4), np.float32)
        n = C.c_uint32(0)
        self._ck(self._L.gs_process_frame(self._h, _ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 16, _ptr(rng),
                                          _ptr(ang), _ptr(res), _ptr(xyz), C.byref(n)))
        return dict(range=rng, angle=ang, no_ground=res, xyz=xyz[: n.value].copy())

    def process_frame_into(self, points, no_ground_out, xyz4_out):
        """One frame with caller-owned (ideally pinned) buffers: no-ground image and cloud only.  Returns the number of points."""
        n = C.c_uint32(0)
        self._ck(self._L.gs_process_frame(self._h, _ptr(points), points.shape[0], points.strides[0] if points.shape[0] else 16, None,
                                          None, _ptr(no_ground_out), _ptr(xyz4_out), C.byref(n)))
        return n.value

    # ---- wire formats (N1)
    def unpack_points(self, data, point_step, off_x, off_y, off_z, off_intensity=-1, big_endian=False):
        """PointCloud2-style byte buffer -> (n, 4) float32 {x, y, z, intensity}"""
        raw = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        n = raw.shape[0] // point_step
        out = np.empty((n, 4), np.float32)
        self._ck(self._L.gs_unpack_points(self._h, _ptr(raw), n, point_step, off_x, off_y, off_z, off_intensity, int(big_endian), _ptr(out)))
        return out

    def process_cloud2(self, data, point_step, off_x, off_y, off_z, big_endian=False):
        H, W = self.shape
        raw = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        n_pts = raw.shape[0] // point_step
        rng, ang, res = (np.empty((H, W), np.uint16) for _ in range(3))
        xyz = np.empty((H * W, 4), np.float32)
        n = C.c_uint32(0)
        self._ck(self._L.gs_process_cloud2(self._h, _ptr(raw), n_pts, point_step, off_x, off_y, off_z, int(big_endian), _ptr(rng), _ptr(ang),
                                           _ptr(res), _ptr(xyz), C.byref(n)))
        return dict(range=rng, angle=ang, no_ground=res, xyz=xyz[: n.value].copy())

    def pack_cloud_xyzrgb(self, xyz4):
        a = np.ascontiguousarray(xyz4, dtype=np.float32)
        out = np.zeros((a.shape[0], 32), np.uint8)
        self._ck(self._L.gs_pack_cloud_xyzrgb(self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def camera_info(self):
        h, w = C.c_int32(0), C.c_int32(0)
        K = (C.c_double * 9)()
        self._ck(self._L.gs_get_camera_info(self._h, C.byref(h), C.byref(w), K))
        return h.value, w.value, [float(x) for x in K]

    # ---- FPGA wire formats + coloured angle image
    def fpga_store_pointcloud(self, points):
        pts = np.ascontiguousarray(points, dtype=np.float32)
        out = np.zeros(pts.shape[0], np.uint64)
        self._ck(self._L.gs_fpga_store_pointcloud(self._h, _ptr(pts), pts.shape[0], pts.strides[0] if pts.shape[0] else 16, _ptr(out)))
        return out

    def fpga_read_range_image(self, words, rows, cols):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.empty((64, 1800), np.uint16)
        self._ck(self._L.gs_fpga_read_range_image(self._h, _ptr(w), rows, cols, _ptr(out)))
        return out

    def fpga_read_angle_image(self, words, rows, cols):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        out = np.empty((rows, cols), np.uint16)
        self._ck(self._L.gs_fpga_read_angle_image(self._h, _ptr(w), rows, cols, _ptr(out)))
        return out

    def colored_angle_image(self, angle):
        a = np.ascontiguousarray(angle, dtype=np.uint16)
        out = np.empty(a.shape + (3,), np.uint8)
        self._ck(self._L.gs_colored_angle_image(self._h, _ptr(a), a.shape[0], a.shape[1], _ptr(out)))
        return out

    # ---- evaluation
    def create_labeled_range_image(self, points32):
        """Alfa_GS::CreateLabeledRangeImage on 32-byte pcl::PointXYZRGB records (uint8 array [n, 32])."""
        H, W = self.shape
        a = np.ascontiguousarray(points32, dtype=np.uint8).reshape(-1, 32)
        out = np.empty((H, W), np.uint16)
        self._ck(self._L.gs_create_labeled_range_image(self._h, _ptr(a), a.shape[0], _ptr(out)))
        return out

    def evaluate(self, og_range, seg_range, labeled_range):
        """Per-frame counters and IoU_g / Recall_g / Recall_mo of Alfa_GS::CheckNumberOfDetectedRIdGnd."""
        out = GsEval()
        imgs = [np.ascontiguousarray(x, dtype=np.uint16) for x in (og_range, seg_range, labeled_range)]
        self._ck(self._L.gs_evaluate(self._h, _ptr(imgs[0]), _ptr(imgs[1]), _ptr(imgs[2]), C.byref(out)))
        return out

    # ---- batch API
    def process_batch_device(self, d_points, h_offsets, n_frames, d_range=0, d_angle=0, d_no_ground=0, d_xyz4=0, d_counts=0):
        """Raw device pointers (ints).  Asynchronous on the context's stream."""
        offs = np.ascontiguousarray(h_offsets, dtype=np.int64)
        self._ck(self._L.gs_process_batch_device(self._h, C.c_void_p(d_points), _ptr(offs), n_frames, C.c_void_p(d_range or None),
                                                 C.c_void_p(d_angle or None), C.c_void_p(d_no_ground or None),
                                                 C.c_void_p(d_xyz4 or None), C.c_void_p(d_counts or None)))

    def process_batch_device_labels(self, d_points, h_offsets, n_frames, d_no_ground=0, d_xyz4=0, d_counts=0, d_labels=0, beta=0.0,
                                    d_range=0, d_angle=0, d_n_components=0):
        """process_batch_device plus the Labeler CCL labels (int32 [n_frames, H, W]).  Raw device pointers.  Asynchronous."""
        offs = np.ascontiguousarray(h_offsets, dtype=np.int64)
        self._ck(self._L.gs_process_batch_device_labels(self._h, C.c_void_p(d_points), _ptr(offs), n_frames, C.c_void_p(d_range or None),
                                                        C.c_void_p(d_angle or None), C.c_void_p(d_no_ground or None), C.c_void_p(d_xyz4 or None),
                                                        C.c_void_p(d_counts or None), C.c_void_p(d_labels or None),
                                                        C.c_void_p(d_n_components or None), float(beta)))

    def label_components_batch_device(self, d_images, d_angles, n_frames, beta, d_labels, d_n_components=0):
        self._ck(self._L.gs_label_components_batch_device(self._h, C.c_void_p(d_images), C.c_void_p(d_angles), n_frames, float(beta),
                                                          C.c_void_p(d_labels), C.c_void_p(d_n_components or None)))

    def process_batch_host(self, h_points, h_offsets, no_ground_out=None, xyz4_out=None, out_offsets=None):
        """numpy (ideally pinned) buffers; h_points is [n, 4] XYZI or [n, 3] XYZ float32 (or any [n, k >= 3] float32 records);
        returns (no_ground_out, xyz4_out, out_offsets)."""
        offs = np.ascontiguousarray(h_offsets, dtype=np.int64)
        n_frames = len(offs) - 1
        if out_offsets is None:
            out_offsets = np.zeros(n_frames + 1, np.int64)
        cap = 0 if xyz4_out is None else xyz4_out.shape[0]
        stride = 16 if h_points.ndim < 2 else int(h_points.strides[0])
        self._ck(self._L.gs_process_batch_host_strided(self._h, _ptr(h_points), stride, _ptr(offs), n_frames, _ptr(no_ground_out), _ptr(xyz4_out),
                                                       cap, _ptr(out_offsets)))
        return no_ground_out, xyz4_out, out_offsets

    def backproject_camera_info(self, image, K, range_as_double):
        """Back-projection with the decoding parameters of a CameraInfo.K (range2cloud_node.cpp:51-58 / alfa_gs.cpp:1010-1015)."""
        img = np.ascontiguousarray(image, dtype=np.uint16)
        H, W = img.shape
        Kc = (C.c_double * 9)(*(list(K) + [0.0] * 9)[:9])
        xyz = np.empty((H * W, 4), np.float32)
        n = C.c_uint32(0)
        self._ck(self._L.gs_backproject_camera_info(self._h, _ptr(img), H, W, Kc, int(range_as_double), _ptr(xyz), C.byref(n)))
        return xyz[: n.value].copy()


class MultiGpu:
    """gs_multi: one worker thread + gs_ctx per GPU, frames sharded in contiguous blocks, no collective."""

    def __init__(self, params, devices=None):
        self._L = load()
        h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices) if devices else None
        rc = self._L.gs_multi_create(C.byref(params), arr, len(devices) if devices else 0, C.byref(h))
        if rc != GS_OK:
            raise GsError(rc, "gs_multi_create failed (%s)" % ("no usable CUDA device" if rc == GS_ECUDA else "invalid parameters / devices"))
        self._h = h
        self.n_gpus = self._L.gs_multi_n_gpus(h)
        self.shape = (params.n_beams, params.n_cols)
        self._bufs = []

    def close(self):
        if getattr(self, "_h", None):
            for ptr, n in self._bufs:
                self._L.gs_multi_host_free(self._h, ptr, n)
            self._bufs = []
            self._L.gs_multi_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def workers(self):
        out = []
        for g in range(self.n_gpus):
            d, n, b = C.c_int(0), C.c_int(0), C.c_int(0)
            self._L.gs_multi_worker_info(self._h, g, C.byref(d), C.byref(n), C.byref(b))
            out.append({"device": d.value, "numa_node": n.value, "cpu_bound": bool(b.value)})
        return out

    def host_array(self, shape, dtype):
        """A pinned numpy array whose pages were first touched shard by shard (NUMA-local to the GPU that will copy them)."""
        dt = np.dtype(dtype)
        n = int(np.prod(shape)) * dt.itemsize
        p = C.c_void_p()
        rc = self._L.gs_multi_host_alloc(self._h, max(n, 1), C.byref(p))
        if rc != GS_OK:
            raise GsError(rc, "gs_multi_host_alloc(%d bytes) failed" % n)
        self._bufs.append((p, max(n, 1)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)

    def process_batch_host(self, h_points, h_offsets, no_ground_out=None, xyz4_out=None):
        """-> (no_ground_out, xyz4_out, first[n_frames], count[n_frames])"""
        offs = np.ascontiguousarray(h_offsets, dtype=np.int64)
        n_frames = len(offs) - 1
        first = np.zeros(n_frames, np.int64)
        count = np.zeros(n_frames, np.uint32)
        cap = 0 if xyz4_out is None else xyz4_out.shape[0]
        stride = 16 if h_points.ndim < 2 else int(h_points.strides[0])
        rc = self._L.gs_multi_process_batch_host(self._h, _ptr(h_points), stride, _ptr(offs), n_frames, _ptr(no_ground_out), _ptr(xyz4_out), cap,
                                                 _ptr(first), _ptr(count))
        if rc != GS_OK:
            raise GsError(rc, self._L.gs_multi_last_error(self._h).decode())
        return no_ground_out, xyz4_out, first, count
