"""N3 — evaluation helpers (Alfa_GS::CreateLabeledRangeImage alfa_gs.cpp:945-995, CheckNumberOfDetectedRIdGnd :831-889).

CPU: the C++ oracle against a plain-Python restatement written from the same reference lines (small inputs).
GPU: the CUDA path (C ABI) against the oracle, bit-exact counters and images, identical float metrics.
"""
import math

import numpy as np
import pytest

from groundsegmentationrosnode_b200 import params as P, synth
from oracle import oracle as O

GROUND_COLOURS = [(75, 0, 75), (255, 0, 255), (255, 150, 255)]


def make_labelled_cloud(p, seed, n_skip_runs=True):
    """32-byte PointXYZRGB records from a synthetic frame: ground-plane returns get a ground colour, some points are pushed
    outside the gates (so that the sticky gnd_flag matters), a few are non-finite."""
    pts = synth.make_frame(p, synth.SEED0 + seed)
    rng = np.random.default_rng(seed)
    n = pts.shape[0]
    rec = np.zeros((n, 32), np.uint8)
    xyz = pts[:, :3].copy()
    ground = xyz[:, 2] < -1.5
    if n_skip_runs:
        # runs of skipped points: beyond max_range, below min_range, or non-finite
        for _ in range(40):
            a = int(rng.integers(0, n - 20))
            ln = int(rng.integers(1, 18))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                xyz[a:a + ln] *= 1000.0
            elif kind == 1:
                xyz[a:a + ln] *= 1e-4
            else:
                xyz[a:a + ln, 0] = np.nan
    rec[:, 0:12] = xyz.astype(np.float32).view(np.uint8).reshape(n, 12)
    rec[:, 12:16] = np.float32(1.0).tobytes() and np.frombuffer(np.float32(1.0).tobytes(), np.uint8)
    cols = rng.integers(0, 3, n)
    other = rng.integers(0, 256, (n, 3)).astype(np.uint8)
    other[:, 1] = 7  # never a ground colour
    for k, (r, g, b) in enumerate(GROUND_COLOURS):
        m = ground & (cols == k)
        other[m] = (r, g, b)
    rec[:, 16] = other[:, 2]  # b
    rec[:, 17] = other[:, 1]  # g
    rec[:, 18] = other[:, 0]  # r
    rec[:, 19] = 255
    return rec


def py_labeled_image(p, rec):
    """alfa_gs.cpp:945-995 line by line (alfa_gs variant, contract Q2/Q3 as in CreateRangeImage)."""
    d = O.derive(p)
    H, W = p.shape
    img = np.ones(H * W, np.uint16)
    xyz = rec[:, 0:12].copy().view(np.float32).reshape(-1, 3)
    flag = 0
    for i in range(rec.shape[0]):
        r, g, b = int(rec[i, 18]), int(rec[i, 17]), int(rec[i, 16])
        if (r, g, b) in GROUND_COLOURS:
            flag = 1
        x, y, z = (float(v) for v in xyz[i])
        if not (math.isfinite(x) and math.isfinite(y) and math.isfinite(z)):
            continue
        rng = math.sqrt(x * x + y * y + z * z)
        if rng < p.min_range or rng > p.max_range:
            continue
        a = math.atan2(y, x)
        az = a if y >= 0 else a + 2 * math.pi
        alt = math.atan2(z, math.hypot(x, y))
        if alt < p.min_angle or alt > p.max_angle:
            continue
        rowf = (H - 1) - ((alt - d.d_altitude / 2 - p.min_angle) / d.d_altitude)
        colf = (az + d.d_azimuth / 2) / d.d_azimuth
        if not colf >= 0.0:
            continue
        row, col = int(rowf), int(colf)
        if row < 0 or row >= H:
            continue
        flat = row * W + col
        if flat >= H * W:
            continue
        if flag == 1:
            img[flat] = 0
        flag = 0
    return img.reshape(H, W)


def py_evaluate(og, seg, lab):
    og, seg, lab = (x.astype(np.int64).ravel() for x in (og, seg, lab))
    c = dict(points=og.size, labeled_gnd=int((lab == 0).sum()), labeled_obst=int(((lab != 0) & (og != 0)).sum()),
             seg_gnd=int(((seg == 0) & (og != 0)).sum()), true_positive=int(((seg == lab) & (og != 0)).sum()),
             false_positive=int(((seg == 0) & (lab != 0) & (og != 0)).sum()), true_negative=int(((seg != 0) & (lab != 0)).sum()),
             false_negative=int(((seg != 0) & (lab == 0)).sum()), true_detected_obst=int(((seg != 0) & (lab != 0)).sum()))
    f = np.float32
    tp, fp, tn, fn = f(c["true_positive"]), f(c["false_positive"]), f(c["true_negative"]), f(c["false_negative"])
    with np.errstate(invalid="ignore", divide="ignore"):
        c["iou_g"] = f(f(tp / f(f(tp + fp) + fn)) * f(100))
        c["recall_g"] = f(f(tp / f(tp + fn)) * f(100))
        c["recall_mo"] = f(f(tn / f(tn + fp)) * f(100))
    return c


def eval_dict(e):
    return {n: getattr(e, n) for n, _ in type(e)._fields_}


def same_metric(a, b):
    return (np.isnan(a) and np.isnan(b)) or np.float32(a) == np.float32(b)


def test_oracle_labeled_image_matches_python_restatement():
    p = P.preset(P.GS_PRESET_VLP16)
    rec = make_labelled_cloud(p, 3)[:6000]
    got = O.create_labeled_range_image(p, rec)
    want = py_labeled_image(p, rec)
    assert np.array_equal(got, want)
    assert (got == 0).sum() > 100 and set(np.unique(got)) <= {0, 1}


def test_oracle_sticky_ground_flag():
    """A ground-coloured point that is skipped by the range gate marks the NEXT kept point (alfa_gs.cpp:961-992)."""
    p = P.preset(P.GS_PRESET_VLP16)
    rec = np.zeros((3, 32), np.uint8)
    xyz = np.array([[1000.0, 0.5, 0.0], [5.0, 0.5, 0.0], [5.0, 3.0, 0.0]], np.float32)  # first: beyond max_range
    rec[:, 0:12] = xyz.view(np.uint8).reshape(3, 12)
    rec[0, 16:19] = (75, 0, 75)  # b g r of (75, 0, 75): ground colour on the skipped point only
    img = O.create_labeled_range_image(p, rec)
    assert (img == 0).sum() == 1  # the second point's pixel; the third is kept with the flag down
    assert np.array_equal(img, py_labeled_image(p, rec))


def test_oracle_evaluate_matches_numpy_restatement():
    p = P.preset(P.GS_PRESET_VLP16)
    rec = make_labelled_cloud(p, 5)
    xyzi = np.zeros((rec.shape[0], 4), np.float32)
    xyzi[:, :3] = rec[:, 0:12].copy().view(np.float32).reshape(-1, 3)
    out = O.process_frame(p, xyzi)
    lab = O.create_labeled_range_image(p, rec)
    e = eval_dict(O.evaluate(p, out["range"], out["no_ground"], lab))
    w = py_evaluate(out["range"], out["no_ground"], lab)
    for k, v in w.items():
        if k in ("iou_g", "recall_g", "recall_mo"):
            assert same_metric(e[k], v), k
        else:
            assert e[k] == v, k
    assert e["true_positive"] > 0 and e["true_negative"] > 0
    # degenerate: everything empty -> 0/0 = NaN, as in the reference's float arithmetic
    z = np.zeros_like(lab)
    e0 = eval_dict(O.evaluate(p, z, z, z))
    assert np.isnan(e0["iou_g"]) and e0["points"] == lab.size and e0["labeled_gnd"] == lab.size


@pytest.mark.gpu
@pytest.mark.parametrize("preset", [P.GS_PRESET_VLP16, P.GS_PRESET_HDL64E_2048])
def test_gpu_evaluation_matches_oracle(preset):
    from groundsegmentationrosnode_b200 import capi

    p = P.preset(preset)
    with capi.Context(p) as ctx:
        for seed in (1, 2):
            rec = make_labelled_cloud(p, 10 + seed)
            lab = ctx.create_labeled_range_image(rec)
            want_lab = O.create_labeled_range_image(p, rec)
            assert np.array_equal(lab, want_lab)
            xyzi = np.zeros((rec.shape[0], 4), np.float32)
            xyzi[:, :3] = rec[:, 0:12].copy().view(np.float32).reshape(-1, 3)
            got = ctx.process_frame(xyzi)
            e = eval_dict(ctx.evaluate(got["range"], got["no_ground"], lab))
            w = eval_dict(O.evaluate(p, got["range"], got["no_ground"], want_lab))
            for k in w:
                if k in ("iou_g", "recall_g", "recall_mo"):
                    assert same_metric(e[k], w[k]), k
                else:
                    assert e[k] == w[k], k
        # edge cases: an empty cloud, a cloud of skipped ground-coloured points only
        assert np.array_equal(ctx.create_labeled_range_image(np.zeros((0, 32), np.uint8)), np.ones(p.shape, np.uint16))
        rec = make_labelled_cloud(p, 3)[:500].copy()
        rec[:, 0:12] = (np.full((500, 3), 1e6, np.float32)).view(np.uint8).reshape(500, 12)
        assert np.array_equal(ctx.create_labeled_range_image(rec), O.create_labeled_range_image(p, rec))
