"""ctypes mirror of the POD structs in include/gs_b200.h plus the reference's sensor presets.

Presets restate the constant blocks of Alfa_GS::Alfa_GS (reference src/alfa_gs.cpp:16-38) and the
(partly commented) parameter blocks of Cloud2RangeNode (src/cloud2range_node.cpp:66-86, 122-135,
281-300); `gs_default_params()` in the C library holds the same table and a test checks they agree.
"""
import ctypes as C

GS_OK, GS_EINVAL, GS_ECUDA, GS_ENOMEM, GS_ESTATE = 0, 1, 2, 3, 4
GS_VARIANT_ALFA_GS, GS_VARIANT_CLOUD2RANGE = 0, 1
GS_PIXEL_LAST, GS_PIXEL_NEAREST = 0, 1
GS_COL_FLAT_ALIAS, GS_COL_WRAP = 0, 1
GS_SMOOTH_NONE, GS_SMOOTH_SG = 0, 1

GS_PRESET_ALFA_GS_DEFAULT = 0
GS_PRESET_VLP16 = 1
GS_PRESET_HDL64E_2048 = 2
GS_PRESET_OS1_128 = 3
GS_PRESET_CLOUD2RANGE_DEFAULT = 4
GS_PRESET_HDL32 = 5


class GsParams(C.Structure):
    _fields_ = [
        ("n_beams", C.c_int32),
        ("rpm", C.c_int32),
        ("sample_freq", C.c_int32),
        ("min_angle", C.c_double),
        ("max_angle", C.c_double),
        ("min_range", C.c_double),
        ("max_range", C.c_double),
        ("ground_angle_threshold", C.c_double),
        ("start_angle_threshold", C.c_double),
        ("window_size", C.c_int32),
        ("erase_ground", C.c_int32),
        ("variant", C.c_int32),
        ("pixel_policy", C.c_int32),
        ("col_policy", C.c_int32),
        ("smoothing", C.c_int32),
        ("repair_gaps", C.c_int32),
        ("reserved", C.c_int32),
    ]

    def copy(self, **kw):
        q = GsParams.from_buffer_copy(self)
        for k, v in kw.items():
            setattr(q, k, v)
        return q

    @property
    def n_cols(self):
        return self.sample_freq * 60 // self.rpm

    @property
    def shape(self):
        return (self.n_beams, self.n_cols)


class GsEval(C.Structure):
    """gs_eval of include/gs_b200.h (CheckNumberOfDetectedRIdGnd, alfa_gs.cpp:831-889)."""
    _fields_ = [(n, C.c_uint64) for n in ("points", "labeled_gnd", "labeled_obst", "seg_gnd", "true_positive", "false_positive",
                                          "true_negative", "false_negative", "true_detected_obst")] + \
               [(n, C.c_float) for n in ("iou_g", "recall_g", "recall_mo")]


class GsDerived(C.Structure):
    _fields_ = [
        ("n_rows", C.c_int32),
        ("n_cols", C.c_int32),
        ("d_azimuth", C.c_double),
        ("d_altitude", C.c_double),
        ("k_ground", C.c_int32),
        ("s_start", C.c_int32),
        ("n_dilate", C.c_int32),
        ("dilate_off", C.c_int32 * 2),
    ]


class GsStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("near_edge_project", C.c_uint64),
        ("near_edge_angle", C.c_uint64),
        ("dropped_points", C.c_uint64),
        ("slow_path_points", C.c_uint64),
        ("fast_path_mismatches", C.c_uint64),
        ("remote_points", C.c_uint64),
    ]


def _mk(n_beams, sample_freq, min_angle, max_angle, min_range, max_range, ground, window, variant=GS_VARIANT_ALFA_GS,
        smoothing=GS_SMOOTH_NONE, rpm=600, start=0.523598, repair_gaps=0):
    return GsParams(n_beams=n_beams, rpm=rpm, sample_freq=sample_freq, min_angle=min_angle, max_angle=max_angle,
                    min_range=min_range, max_range=max_range, ground_angle_threshold=ground,
                    start_angle_threshold=start, window_size=window, erase_ground=1, variant=variant,
                    pixel_policy=GS_PIXEL_LAST, col_policy=GS_COL_FLAT_ALIAS, smoothing=smoothing, repair_gaps=repair_gaps,
                    reserved=0)


def preset(which):
    """Python twin of gs_default_params()."""
    if which == GS_PRESET_ALFA_GS_DEFAULT:  # alfa_gs.cpp:16-38 -> 64 x 1800
        return _mk(64, 18000, -0.453785, 0.09599, 0.5, 100.0, 0.115, 11)
    if which == GS_PRESET_VLP16:  # cloud2range_node.cpp:66-86 -> 16 x 1800 (BASELINE config C1)
        return _mk(16, 18000, -0.2617993878, 0.2617993878, 0.5, 100.0, 0.087266, 5)
    if which == GS_PRESET_HDL64E_2048:  # alfa_gs.cpp:16-38 with sample_freq 20480 -> 64 x 2048 (C2/C4)
        return _mk(64, 20480, -0.453785, 0.09599, 0.5, 100.0, 0.115, 11)
    if which == GS_PRESET_OS1_128:  # cloud2range_node.cpp:281-300, beams doubled, 20480 -> 128 x 2048 (C3/C5)
        return _mk(128, 20480, -0.392699, 0.392699, 0.3, 120.0, 0.09, 7)
    if which == GS_PRESET_CLOUD2RANGE_DEFAULT:  # cloud2range_node.cpp:281-300 as shipped: OS1-64, 64 x 2000, RepairGaps + SG-7 (:468-470)
        return _mk(64, 20000, -0.392699, 0.392699, 0.3, 120.0, 0.09, 7, variant=GS_VARIANT_CLOUD2RANGE,
                   smoothing=GS_SMOOTH_SG, repair_gaps=1)
    if which == GS_PRESET_HDL32:  # cloud2range_node.cpp:122-135 -> 32 x 1800
        return _mk(32, 18000, -0.535293, 0.186227, 0.5, 100.0, 0.087266, 5)
    raise ValueError("unknown preset %r" % (which,))
