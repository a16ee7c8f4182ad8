"""numpy/ctypes front end of the CPU oracle — TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg.
The product package (groundsegmentationrosnode_b200) never imports this module.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from groundsegmentationrosnode_b200.params import GsDerived, GsEval, GsParams  # noqa: E402  (POD layouts only)

from . import build as _build  # noqa: E402

_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")

_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build_oracle())
        L = _lib
        L.gso_derive.argtypes = [C.POINTER(GsParams), C.POINTER(GsDerived)]
        L.gso_k_of_threshold.argtypes = [C.c_double]
        L.gso_s_of_threshold.argtypes = [C.c_double]
        L.gso_uniform_kernel.argtypes = [C.c_int, _u8p]
        L.gso_sg_kernel.argtypes = [C.c_int, _f32p]
        L.gso_dilate.argtypes = [C.c_int, C.c_int, C.c_int, _u8p, _u8p]
        L.gso_create_range_image.argtypes = [C.POINTER(GsParams), C.c_void_p, C.c_size_t, C.c_size_t, _u16p,
                                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gso_create_angle_image.argtypes = [C.POINTER(GsParams), _u16p, _u16p, C.POINTER(C.c_uint64)]
        L.gso_sg_smooth.argtypes = [C.POINTER(GsParams), _u16p, C.c_int, _u16p]
        L.gso_repair_gaps.argtypes = [C.POINTER(GsParams), _u16p, C.c_int, C.c_float, _u16p]
        L.gso_erase_bfs.argtypes = [C.POINTER(GsParams), _u16p, _u16p, C.c_double, C.c_double, C.c_int, _u16p,
                                    C.c_void_p, C.c_void_p]
        L.gso_label_one_component.argtypes = [C.POINTER(GsParams), C.c_int, C.c_int, C.c_int, C.c_float, _u16p, _u16p,
                                              _u8p]
        L.gso_label_components.argtypes = [C.POINTER(GsParams), _u16p, _u16p, C.c_double, _i32p, C.POINTER(C.c_int32)]
        L.gso_camera_cb.argtypes = [C.POINTER(GsParams), _u16p, _f32p, C.POINTER(C.c_uint32)]
        L.gso_process_frame.argtypes = [C.POINTER(GsParams), C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p]
        L.gso_process_batch.argtypes = [C.POINTER(GsParams), _f32p, _i64p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.POINTER(C.c_double), C.c_void_p]
        L.gso_create_labeled_range_image.argtypes = [C.POINTER(GsParams), C.c_void_p, C.c_size_t, _u16p]
        L.gso_evaluate.argtypes = [C.POINTER(GsParams), _u16p, _u16p, _u16p, C.POINTER(GsEval)]
        for name, n in (("gso_point_range", 3), ("gso_point_azimuth", 2), ("gso_point_altitude", 3), ("gso_rad_deg", 1)):
            fn = getattr(L, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_double] * n
        for fn in (L.gso_atan2f, L.gso_sinf, L.gso_cosf):
            fn.restype = C.c_float
        L.gso_atan2f.argtypes = [C.c_float, C.c_float]
        L.gso_sinf.argtypes = [C.c_float]
        L.gso_cosf.argtypes = [C.c_float]
    return _lib


def ref():
    """The reference's own Labeler/label.cpp + lib/utils.h (oracle/_ref), or None if it could not be built
    and did not travel."""
    global _ref
    if _ref is None:
        path = _build.build_ref()
        if path is None:
            return None
        _ref = C.CDLL(path)
        R = _ref
        for name, n in (("gsref_point_range", 3), ("gsref_point_azimuth", 2), ("gsref_point_altitude", 3),
                        ("gsref_deg_rad", 1), ("gsref_rad_deg", 1)):
            fn = getattr(R, name)
            fn.restype = C.c_double
            fn.argtypes = [C.c_double] * n
        R.gsref_wrap_cols.argtypes = [C.c_int, C.c_int, C.c_int]
        R.gsref_label_one_component.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _u16p, _u16p,
                                                _u8p]
        R.gsref_label_one_component.restype = None
        R.gsref_ground_labels.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _u16p, _u16p, _u8p]
        R.gsref_ground_labels.restype = None
    return _ref


def _check(rc, what):
    if rc != 0:
        raise ValueError("%s failed with status %d" % (what, rc))


def derive(p):
    d = GsDerived()
    _check(lib().gso_derive(C.byref(p), C.byref(d)), "gso_derive")
    return d


def uniform_kernel(ks):
    out = np.zeros(ks, np.uint8)
    _check(lib().gso_uniform_kernel(ks, out), "uniform kernel")
    return out


def sg_kernel(ws):
    out = np.zeros(ws, np.float32)
    _check(lib().gso_sg_kernel(ws, out), "sg kernel")
    return out


def dilate(label_image, ks):
    lab = np.ascontiguousarray(label_image, dtype=np.uint8)
    out = np.zeros_like(lab)
    _check(lib().gso_dilate(lab.shape[0], lab.shape[1], ks, lab, out), "dilate")
    return out


def sg_smooth_image(image, ws):
    """SG smoothing of an arbitrary H x W u16 image (parameters only carry the shape)."""
    img = np.ascontiguousarray(image, dtype=np.uint16)
    p = GsParams(n_beams=img.shape[0], rpm=60, sample_freq=img.shape[1], min_angle=-0.1, max_angle=0.1,
                 min_range=0.5, max_range=100.0, window_size=ws)
    out = np.zeros_like(img)
    _check(lib().gso_sg_smooth(C.byref(p), img, ws, out), "sg smooth")
    return out


def _points_arg(points):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    assert pts.ndim == 2 and pts.shape[1] >= 3
    return pts, pts.ctypes.data_as(C.c_void_p), pts.shape[0], pts.strides[0]


def create_range_image(p, points, with_counts=False):
    H, W = p.shape
    pts, ptr, n, stride = _points_arg(points)
    img = np.zeros((H, W), np.uint16)
    ne, dr = C.c_uint64(0), C.c_uint64(0)
    _check(lib().gso_create_range_image(C.byref(p), ptr, n, stride, img, C.byref(ne), C.byref(dr)), "range image")
    return (img, ne.value, dr.value) if with_counts else img


def create_angle_image(p, range_image, with_counts=False):
    out = np.zeros(p.shape, np.uint16)
    ne = C.c_uint64(0)
    _check(lib().gso_create_angle_image(C.byref(p), np.ascontiguousarray(range_image), out, C.byref(ne)), "angle image")
    return (out, ne.value) if with_counts else out


def sg_smooth(p, angle_image, ws):
    out = np.zeros(p.shape, np.uint16)
    _check(lib().gso_sg_smooth(C.byref(p), np.ascontiguousarray(angle_image), ws, out), "sg smooth")
    return out


def repair_gaps(p, range_image, step=5, depth_threshold=1.0):
    out = np.zeros(p.shape, np.uint16)
    _check(lib().gso_repair_gaps(C.byref(p), np.ascontiguousarray(range_image), step, depth_threshold, out), "repair")
    return out


def erase_bfs(p, range_image, angle_image, ground_thr=None, start_thr=None, kernel_size=None):
    """-> (res u16, labels u8, dilated u8)"""
    ground_thr = p.ground_angle_threshold if ground_thr is None else ground_thr
    start_thr = p.start_angle_threshold if start_thr is None else start_thr
    kernel_size = p.window_size if kernel_size is None else kernel_size
    res = np.zeros(p.shape, np.uint16)
    lab = np.zeros(p.shape, np.uint8)
    dil = np.zeros(p.shape, np.uint8)
    _check(lib().gso_erase_bfs(C.byref(p), np.ascontiguousarray(range_image), np.ascontiguousarray(angle_image),
                               ground_thr, start_thr, kernel_size, res, lab.ctypes.data_as(C.c_void_p),
                               dil.ctypes.data_as(C.c_void_p)), "erase bfs")
    return res, lab, dil


def label_one_component(p, label, row, col, thr, range_image, angle_image, label_image):
    out = np.ascontiguousarray(label_image, dtype=np.uint8).copy()
    _check(lib().gso_label_one_component(C.byref(p), label, row, col, thr, np.ascontiguousarray(range_image),
                                         np.ascontiguousarray(angle_image), out), "label one")
    return out


def label_components(p, image, angle_image, beta):
    out = np.zeros(p.shape, np.int32)
    n = C.c_int32(0)
    _check(lib().gso_label_components(C.byref(p), np.ascontiguousarray(image), np.ascontiguousarray(angle_image), beta,
                                      out, C.byref(n)), "label components")
    return out, n.value


def camera_cb(p, image):
    H, W = p.shape
    xyz = np.zeros((H * W, 4), np.float32)
    n = C.c_uint32(0)
    _check(lib().gso_camera_cb(C.byref(p), np.ascontiguousarray(image), xyz, C.byref(n)), "camera cb")
    return xyz[: n.value].copy()


def process_frame(p, points):
    """-> dict(range, angle, no_ground, xyz, stage_ms)"""
    H, W = p.shape
    pts, ptr, n, stride = _points_arg(points)
    rng, ang, res = (np.zeros((H, W), np.uint16) for _ in range(3))
    xyz = np.zeros((H * W, 4), np.float32)
    cnt = C.c_uint32(0)
    ms = np.zeros(4, np.float64)
    _check(lib().gso_process_frame(C.byref(p), ptr, n, stride, rng.ctypes.data_as(C.c_void_p),
                                   ang.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p),
                                   xyz.ctypes.data_as(C.c_void_p), C.byref(cnt), ms.ctypes.data_as(C.c_void_p)),
           "process frame")
    return dict(range=rng, angle=ang, no_ground=res, xyz=xyz[: cnt.value].copy(), stage_ms=ms)


def process_batch(p, points_xyzi, offsets, n_threads=1, want_outputs=True):
    """CPU baseline driver. -> dict(no_ground, xyz, counts, seconds, stage_ms)"""
    H, W = p.shape
    P = H * W
    pts = np.ascontiguousarray(points_xyzi, dtype=np.float32).reshape(-1)
    offs = np.ascontiguousarray(offsets, dtype=np.int64)
    F = len(offs) - 1
    res = np.zeros((F, H, W), np.uint16) if want_outputs else None
    xyz = np.zeros((F, P, 4), np.float32) if want_outputs else None
    cnt = np.zeros(F, np.uint32)
    sec = C.c_double(0)
    ms = np.zeros(4, np.float64)
    _check(lib().gso_process_batch(C.byref(p), pts, offs, F, n_threads,
                                   res.ctypes.data_as(C.c_void_p) if want_outputs else None,
                                   xyz.ctypes.data_as(C.c_void_p) if want_outputs else None,
                                   cnt.ctypes.data_as(C.c_void_p), C.byref(sec), ms.ctypes.data_as(C.c_void_p)),
           "process batch")
    return dict(no_ground=res, xyz=xyz, counts=cnt, seconds=sec.value, stage_ms=ms)


def create_labeled_range_image(p, points32):
    """Alfa_GS::CreateLabeledRangeImage on 32-byte PointXYZRGB records (uint8 array [n, 32])."""
    H, W = p.shape
    pts = np.ascontiguousarray(points32, np.uint8).reshape(-1, 32)
    img = np.zeros((H, W), np.uint16)
    _check(lib().gso_create_labeled_range_image(C.byref(p), pts.ctypes.data_as(C.c_void_p), pts.shape[0], img), "create_labeled_range_image")
    return img


def evaluate(p, og, seg, labeled):
    """Per-frame counters and metrics of Alfa_GS::CheckNumberOfDetectedRIdGnd."""
    out = GsEval()
    _check(lib().gso_evaluate(C.byref(p), np.ascontiguousarray(og, np.uint16), np.ascontiguousarray(seg, np.uint16),
                              np.ascontiguousarray(labeled, np.uint16), C.byref(out)), "evaluate")
    return out


# ---- the reference's own nodes, compiled verbatim into oracle/_ref (oracle/ref_alfa_wrap.cpp, ref_c2r_wrap.cpp,
# ref_r2c_wrap.cpp) — used to pin the restatement above and as the CPU arm of bench.py ("kind": "reference") ----------
_node_bound = False


def _bind_nodes(R):
    global _node_bound
    if _node_bound:
        return
    vp, sz = C.c_void_p, C.c_size_t
    R.gsref_alfa_ctor_values.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_int)]
    R.gsref_alfa_create.argtypes = [C.POINTER(GsParams)]
    R.gsref_alfa_create.restype = vp
    R.gsref_alfa_destroy.argtypes = [vp]
    R.gsref_alfa_shape.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    R.gsref_assert_failures.restype = C.c_long
    R.gsref_log.argtypes = [C.c_char_p, C.c_int]
    R.gsref_alfa_create_range_image.argtypes = [vp, vp, sz, sz, _u16p]
    R.gsref_alfa_create_angle_img.argtypes = [vp, _u16p, _u16p]
    R.gsref_alfa_repair_gaps.argtypes = [vp, _u16p, C.c_int, C.c_float, _u16p]
    R.gsref_alfa_sg_smooth.argtypes = [vp, _u16p, C.c_int, C.c_int, C.c_int, _u16p]
    R.gsref_alfa_sg_kernel.argtypes = [vp, C.c_int, _f32p]
    R.gsref_alfa_uniform_kernel.argtypes = [vp, C.c_int, _u8p]
    R.gsref_alfa_dilate.argtypes = [vp, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
    R.gsref_alfa_erase_bfs.argtypes = [vp, _u16p, _u16p, C.c_double, C.c_double, C.c_int, _u16p]
    R.gsref_alfa_camera_cb.argtypes = [vp, _u16p, vp, sz]
    R.gsref_alfa_camera_cb.restype = sz
    R.gsref_alfa_labeled_range_image.argtypes = [vp, vp, sz, sz, vp, _u16p]
    R.gsref_alfa_check_detected.argtypes = [vp, _u16p, _u16p, _u16p]
    R.gsref_alfa_colored_angle_image.argtypes = [vp, _u16p, C.c_int, C.c_int, _u8p]
    R.gsref_alfa_process_frame.argtypes = [vp, vp, sz, sz, vp, vp, sz]
    R.gsref_alfa_process_frame.restype = C.c_long
    R.gsref_node_store_pointcloud_hardware.argtypes = [vp, vp, sz, sz, vp]
    R.gsref_node_read_hardware_pointcloud.argtypes = [vp, vp, C.c_uint, C.c_uint, _u16p]
    R.gsref_node_read_hardware_filtered_angle_image.argtypes = [vp, vp, C.c_uint, C.c_uint, _u16p]
    R.gsref_c2r_create.argtypes = [C.POINTER(GsParams)]
    R.gsref_c2r_create.restype = vp
    R.gsref_c2r_destroy.argtypes = [vp]
    R.gsref_c2r_derived.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                    C.POINTER(C.c_double), C.POINTER(C.c_double)]
    R.gsref_c2r_cloud_cb.argtypes = [vp, vp, sz, sz, _u16p, C.POINTER(C.c_double)]
    R.gsref_c2r_set_thresholds.argtypes = [vp, C.c_double, C.c_double]
    R.gsref_c2r_repair_gaps.argtypes = [vp, _u16p, C.c_int, C.c_float, _u16p]
    R.gsref_c2r_create_angle_img.argtypes = [vp, _u16p, _u16p]
    R.gsref_c2r_sg_smooth.argtypes = [vp, _u16p, C.c_int, _u16p]
    R.gsref_c2r_erase_ground_bfs.argtypes = [vp, _u16p, _u16p, C.c_double, C.c_double, C.c_int, _u16p]
    R.gsref_r2c_create.restype = vp
    R.gsref_r2c_destroy.argtypes = [vp]
    R.gsref_r2c_camera_cb.argtypes = [vp, _u16p, C.c_int, C.c_int, C.POINTER(C.c_double), vp, sz]
    R.gsref_r2c_camera_cb.restype = C.c_long
    _node_bound = True


def ref_log():
    R = ref()
    buf = C.create_string_buffer(1 << 16)
    R.gsref_log(buf, len(buf))
    return buf.value.decode("utf-8", "replace").splitlines()


class RefAlfaGS:
    """The reference's Alfa_GS node (src/alfa_gs.cpp + src/alfa_node.cpp + Labeler/label.cpp, verbatim).  One per thread."""

    def __init__(self, p=None):
        self.R = ref()
        if self.R is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _bind_nodes(self.R)
        self.h = self.R.gsref_alfa_create(C.byref(p) if p is not None else None)
        if not self.h:
            raise RuntimeError("the reference node came up on its FPGA path or its derivations disagree")
        r, c = C.c_int(0), C.c_int(0)
        self.R.gsref_alfa_shape(self.h, C.byref(r), C.byref(c))
        self.shape = (r.value, c.value)

    def close(self):
        if self.h:
            self.R.gsref_alfa_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _img(self):
        return np.zeros(self.shape, np.uint16)

    def create_range_image(self, points):
        pts, ptr, n, stride = _points_arg(points)
        out = self._img()
        self.R.gsref_alfa_create_range_image(self.h, ptr, n, stride, out)
        return out

    def create_angle_img(self, range_image):
        out = self._img()
        self.R.gsref_alfa_create_angle_img(self.h, np.ascontiguousarray(range_image, np.uint16), out)
        return out

    def repair_gaps(self, range_image, step=5, depth_threshold=1.0):
        out = self._img()
        self.R.gsref_alfa_repair_gaps(self.h, np.ascontiguousarray(range_image, np.uint16), step, depth_threshold, out)
        return out

    def sg_smooth(self, image, ws):
        img = np.ascontiguousarray(image, np.uint16)
        out = np.zeros_like(img)
        if self.R.gsref_alfa_sg_smooth(self.h, img, img.shape[0], img.shape[1], ws, out):
            raise ValueError("std::logic_error")
        return out

    def sg_kernel(self, ws):
        out = np.zeros(ws, np.float32)
        if self.R.gsref_alfa_sg_kernel(self.h, ws, out):
            raise ValueError("std::logic_error")
        return out

    def uniform_kernel(self, ks):
        out = np.zeros(ks, np.uint8)
        if self.R.gsref_alfa_uniform_kernel(self.h, ks, out):
            raise ValueError("std::logic_error")
        return out

    def dilate(self, label_image, ks):
        lab = np.ascontiguousarray(label_image, np.uint8)
        out = np.zeros_like(lab)
        if self.R.gsref_alfa_dilate(self.h, lab, lab.shape[0], lab.shape[1], ks, out):
            raise ValueError("std::logic_error")
        return out

    def erase_bfs(self, range_image, angle_image, ground_thr, start_thr, kernel_size):
        out = self._img()
        if self.R.gsref_alfa_erase_bfs(self.h, np.ascontiguousarray(range_image, np.uint16),
                                       np.ascontiguousarray(angle_image, np.uint16), ground_thr, start_thr, kernel_size, out):
            raise ValueError("std::logic_error")
        return out

    def camera_cb(self, image):
        """-> float32 [n, 4] {x, y, z, 1.0} (the first 16 bytes of the 32-byte pcl::PointXYZRGB records)"""
        H, W = self.shape
        buf = np.zeros((H * W, 8), np.float32)
        n = self.R.gsref_alfa_camera_cb(self.h, np.ascontiguousarray(image, np.uint16), buf.ctypes.data_as(C.c_void_p), H * W)
        return buf[:n, :4].copy()

    def labeled_range_image(self, points, rgb):
        pts, ptr, n, stride = _points_arg(points)
        rgb = np.ascontiguousarray(rgb, np.uint8)
        out = self._img()
        self.R.gsref_alfa_labeled_range_image(self.h, ptr, n, stride, rgb.ctypes.data_as(C.c_void_p), out)
        return out

    def check_detected(self, og, seg, labeled):
        """CheckNumberOfDetectedRIdGnd reports through ROS_INFO only: returns the logged lines."""
        self.R.gsref_alfa_check_detected(self.h, np.ascontiguousarray(og, np.uint16), np.ascontiguousarray(seg, np.uint16),
                                         np.ascontiguousarray(labeled, np.uint16))
        return ref_log()

    def colored_angle_image(self, angle_image):
        img = np.ascontiguousarray(angle_image, np.uint16)
        out = np.zeros(img.shape + (3,), np.uint8)
        self.R.gsref_alfa_colored_angle_image(self.h, img, img.shape[0], img.shape[1], out)
        return out

    def process_frame(self, points, want_outputs=True):
        """A PointCloud2 through AlfaNode::cloud_cb -> Alfa_GS::process_pointcloud; what the node published."""
        pts, ptr, n, stride = _points_arg(points)
        H, W = self.shape
        if not want_outputs:
            return self.R.gsref_alfa_process_frame(self.h, ptr, n, stride, None, None, 0)
        img = self._img()
        buf = np.zeros((H * W, 8), np.float32)
        np_out = self.R.gsref_alfa_process_frame(self.h, ptr, n, stride, img.ctypes.data_as(C.c_void_p),
                                                 buf.ctypes.data_as(C.c_void_p), H * W)
        if np_out < 0:
            raise RuntimeError("the reference node published nothing")
        return dict(no_ground=img, xyz=buf[:np_out, :4].copy(), records=buf[:np_out].copy())

    def store_pointcloud_hardware(self, points):
        pts, ptr, n, stride = _points_arg(points)
        out = np.zeros(n, np.uint64)
        self.R.gsref_node_store_pointcloud_hardware(self.h, ptr, n, stride, out.ctypes.data_as(C.c_void_p))
        return out

    def read_hardware_pointcloud(self, words, rows, cols):
        w = np.ascontiguousarray(words, np.uint64)
        out = np.zeros((64, 1800), np.uint16)  # the reference hard-codes this shape (alfa_node.cpp:162)
        self.R.gsref_node_read_hardware_pointcloud(self.h, w.ctypes.data_as(C.c_void_p), rows, cols, out)
        return out

    def read_hardware_filtered_angle_image(self, words, rows, cols):
        w = np.ascontiguousarray(words, np.uint64).copy()  # the reference shifts the words it reads
        out = np.zeros((rows, cols), np.uint16)
        self.R.gsref_node_read_hardware_filtered_angle_image(self.h, w.ctypes.data_as(C.c_void_p), rows, cols, out)
        return out


def ref_ctor_values():
    """What the UNTOUCHED Alfa_GS constructor computed: (n_cols, d_azimuth, d_altitude, K[0..5], hw)."""
    R = ref()
    _bind_nodes(R)
    nc, hw = C.c_int(0), C.c_int(0)
    da, dl = C.c_double(0), C.c_double(0)
    K = (C.c_double * 6)()
    R.gsref_alfa_ctor_values(C.byref(nc), C.byref(da), C.byref(dl), K, C.byref(hw))
    return nc.value, da.value, dl.value, list(K), hw.value


class RefCloud2Range:
    """The reference's Cloud2RangeNode (src/cloud2range_node.cpp, verbatim).  One per thread."""

    def __init__(self, p=None):
        self.R = ref()
        if self.R is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _bind_nodes(self.R)
        self.h = self.R.gsref_c2r_create(C.byref(p) if p is not None else None)
        r, c = C.c_int(0), C.c_int(0)
        da, dl = C.c_double(0), C.c_double(0)
        K = (C.c_double * 6)()
        self.R.gsref_c2r_derived(self.h, C.byref(r), C.byref(c), C.byref(da), C.byref(dl), K)
        self.shape = (r.value, c.value)
        self.d_azimuth, self.d_altitude, self.K = da.value, dl.value, list(K)

    def close(self):
        if self.h:
            self.R.gsref_c2r_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def set_thresholds(self, ground_thr, start_thr):
        self.R.gsref_c2r_set_thresholds(self.h, ground_thr, start_thr)

    def cloud_cb(self, points):
        """-> (published mono16 image, published CameraInfo.K[0..8])"""
        pts, ptr, n, stride = _points_arg(points)
        img = np.zeros(self.shape, np.uint16)
        K = (C.c_double * 9)()
        if self.R.gsref_c2r_cloud_cb(self.h, ptr, n, stride, img, K):
            raise RuntimeError("the reference node published nothing")
        return img, list(K)

    def repair_gaps(self, range_image, step=5, depth_threshold=1.0):
        out = np.zeros(self.shape, np.uint16)
        self.R.gsref_c2r_repair_gaps(self.h, np.ascontiguousarray(range_image, np.uint16), step, depth_threshold, out)
        return out

    def create_angle_img(self, range_image):
        out = np.zeros(self.shape, np.uint16)
        self.R.gsref_c2r_create_angle_img(self.h, np.ascontiguousarray(range_image, np.uint16), out)
        return out

    def sg_smooth(self, image, ws):
        out = np.zeros(self.shape, np.uint16)
        if self.R.gsref_c2r_sg_smooth(self.h, np.ascontiguousarray(image, np.uint16), ws, out):
            raise ValueError("std::logic_error")
        return out

    def erase_ground_bfs(self, range_image, angle_image, ground_thr, start_thr, kernel_size):
        out = np.zeros(self.shape, np.uint16)
        if self.R.gsref_c2r_erase_ground_bfs(self.h, np.ascontiguousarray(range_image, np.uint16),
                                             np.ascontiguousarray(angle_image, np.uint16), ground_thr, start_thr, kernel_size, out):
            raise ValueError("std::logic_error")
        return out


class RefRange2Cloud:
    """The reference's Range2CloudNode (src/range2cloud_node.cpp, verbatim).  One per thread."""

    def __init__(self):
        self.R = ref()
        if self.R is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _bind_nodes(self.R)
        self.h = self.R.gsref_r2c_create()

    def close(self):
        if self.h:
            self.R.gsref_r2c_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def camera_cb(self, image, K):
        img = np.ascontiguousarray(image, np.uint16)
        K9 = (C.c_double * 9)(*(list(K) + [0.0] * 9)[:9])
        buf = np.zeros((img.size, 4), np.float32)
        n = self.R.gsref_r2c_camera_cb(self.h, img, img.shape[0], img.shape[1], K9, buf.ctypes.data_as(C.c_void_p), img.size)
        if n < 0:
            raise RuntimeError("the reference node published nothing")
        return buf[:n].copy()


def ref_assert_failures():
    R = ref()
    _bind_nodes(R)
    return R.gsref_assert_failures()
