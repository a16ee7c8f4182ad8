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
