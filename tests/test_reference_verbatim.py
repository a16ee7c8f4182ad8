"""The oracle's restatement of the hot path against the reference's OWN translation units — src/alfa_gs.cpp,
src/alfa_node.cpp, Labeler/label.cpp, src/cloud2range_node.cpp, src/range2cloud_node.cpp — compiled verbatim from
/root/reference into oracle/_ref/libgs_ref.so against stand-in ROS / PCL / OpenCV headers (oracle/shim, recipe oracle/build.py).

This is the pin of SURVEY.md §8(c): CreateRangeImage (R1), the cloud2range projection loop (R1'), CreateAngleImg (A1),
SavitskyGolaySmoothing + GetSavitskyGolayKernel (S2), GetUniformKernel + cv::dilate (G2), EraseBFS / EraseGroundBFS (G1),
CameraCb x 2 (B1, B1'), the constructors' derived values (P0) and the whole frame as the node runs it under ROS
(PointCloud2 -> AlfaNode::cloud_cb -> Alfa_GS::process_pointcloud -> publishers; Cloud2RangeNode::CloudCb) are compared
with oracle/gs_oracle.cpp bit for bit.  The stand-in cv::dilate / cv::filter2D are themselves checked against the committed
OpenCV 4.13 vectors (tests/golden).  The library travels prebuilt to the GPU box; these tests need no GPU.
"""
import os
import re

import numpy as np
import pytest

from groundsegmentationrosnode_b200 import params as P, synth

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_golden.npz"))


@pytest.fixture(scope="module")
def O(oracle):
    if oracle.ref() is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return oracle


def _small(H, W, **kw):
    base = dict(n_beams=H, rpm=600, sample_freq=W * 10, min_angle=-0.4, max_angle=0.15, min_range=0.5, max_range=100.0,
                ground_angle_threshold=0.115, start_angle_threshold=0.523598, window_size=11, erase_ground=1)
    base.update(kw)
    return P.GsParams(**base)


ALFA_SHAPES = {
    "alfa_default_64x1800": P.GS_PRESET_ALFA_GS_DEFAULT,
    "C1_vlp16_16x1800": P.GS_PRESET_VLP16,
    "hdl32_32x1800": P.GS_PRESET_HDL32,
    "C2_hdl64_64x2048": P.GS_PRESET_HDL64E_2048,
    "C3_os1_128x2048": P.GS_PRESET_OS1_128,
}


# ---- P0: what the constructors derive ------------------------------------------------------------------------------
def test_alfa_gs_constructor_values(O):
    n_cols, d_az, d_alt, K, hw = O.ref_ctor_values()  # the untouched constructor, src/alfa_gs.cpp:12-93
    assert hw == 0, "the reference node must stay on its software path (no /dev/mem)"
    p = P.preset(P.GS_PRESET_ALFA_GS_DEFAULT)
    d = O.derive(p)
    assert (d.n_rows, d.n_cols) == (64, n_cols) == (64, 1800)
    assert d.d_azimuth == d_az and d.d_altitude == d_alt
    assert K == [p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth, d.d_altitude]


@pytest.mark.parametrize("preset", [P.GS_PRESET_CLOUD2RANGE_DEFAULT, P.GS_PRESET_VLP16, P.GS_PRESET_OS1_128])
def test_cloud2range_constructor_values(O, preset):
    # every parameter arrives through NodeHandle::param, the derivations are the constructor's own (cloud2range_node.cpp:305-328)
    p = P.preset(preset)
    node = O.RefCloud2Range(None if preset == P.GS_PRESET_CLOUD2RANGE_DEFAULT else p)
    d = O.derive(p)
    assert node.shape == (d.n_rows, d.n_cols)
    assert node.d_azimuth == d.d_azimuth and node.d_altitude == d.d_altitude
    assert node.K == [p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth, d.d_altitude]


# ---- the stand-in OpenCV calls and the reference's kernels against real OpenCV answers ----------------------------------
@pytest.mark.parametrize("ks", [int(k) for k in G["dil_ks"]])
def test_reference_uniform_kernel_and_dilate_against_opencv(O, ks):
    node = O.RefAlfaGS()
    assert node.uniform_kernel(ks).tolist() == G["dil_kernel_%d" % ks].tolist() == O.uniform_kernel(ks).tolist()
    assert np.array_equal(node.dilate(G["dil_in_%d" % ks], ks), G["dil_out_%d" % ks])


@pytest.mark.parametrize("ws", [int(w) for w in G["sg_ws"]])
def test_reference_savitzky_golay_against_opencv(O, ws):
    node = O.RefAlfaGS()
    assert np.array_equal(node.sg_kernel(ws), G["sg_kernel_%d" % ws])
    assert np.array_equal(node.sg_kernel(ws), O.sg_kernel(ws))
    for tag in "abc":
        assert np.array_equal(node.sg_smooth(G["sg_in_%d_%s" % (ws, tag)], ws), G["sg_out_%d_%s" % (ws, tag)])


def test_reference_window_errors(O):
    node = O.RefAlfaGS()
    for bad in (4, 6, 13):  # std::logic_error in GetSavitskyGolayKernel (alfa_gs.cpp:642-650)
        with pytest.raises(ValueError):
            node.sg_kernel(bad)
        with pytest.raises(ValueError):
            O.sg_kernel(bad)
    with pytest.raises(ValueError):
        node.uniform_kernel(6)
    with pytest.raises(ValueError):
        O.uniform_kernel(6)


# ---- the stages of Alfa_GS, one member function each -----------------------------------------------------------------
@pytest.mark.parametrize("name", list(ALFA_SHAPES))
def test_alfa_gs_stages_match_oracle(O, name):
    p = P.preset(ALFA_SHAPES[name])
    node = O.RefAlfaGS(p)
    assert node.shape == p.shape
    pts = synth.make_frame(p, synth.SEED0 + 21)
    rng_ref = node.create_range_image(pts)                       # Alfa_GS::CreateRangeImage, alfa_gs.cpp:247-298
    assert np.array_equal(rng_ref, O.create_range_image(p, pts))
    ang_ref = node.create_angle_img(rng_ref)                     # CreateAngleImg, :335-386
    assert np.array_equal(ang_ref, O.create_angle_image(p, rng_ref))
    res_ref = node.erase_bfs(rng_ref, ang_ref, p.ground_angle_threshold, p.start_angle_threshold, p.window_size)  # EraseBFS :715-766
    res_or, _, _ = O.erase_bfs(p, rng_ref, ang_ref)
    assert np.array_equal(res_ref, res_or)
    assert np.array_equal(node.camera_cb(res_ref), O.camera_cb(p, res_ref))  # CameraCb :997-1074, bit-exact floats
    assert O.ref_assert_failures() == 0


@pytest.mark.parametrize("kind", ["edges", "collisions", "sparse"])
@pytest.mark.parametrize("name", ["C1_vlp16_16x1800", "C2_hdl64_64x2048"])
def test_alfa_gs_adversarial_frames(O, name, kind):
    # cell-edge points (incl. the flat alias column W, Q2), several returns per pixel in random order (Q1), empty columns with
    # empty seeds and the column-1 quirk (Q5, Q7).  Non-finite inputs are undefined behaviour in the reference (Q3): not run.
    p = P.preset(ALFA_SHAPES[name])
    node = O.RefAlfaGS(p)
    pts = synth.make_adversarial_frame(p, synth.SEED0 + 5, kind)
    got = node.process_frame(pts)
    want = O.process_frame(p, pts)
    assert np.array_equal(node.create_range_image(pts), want["range"])
    assert np.array_equal(node.create_angle_img(want["range"]), want["angle"])
    assert np.array_equal(got["no_ground"], want["no_ground"])
    assert np.array_equal(got["xyz"], want["xyz"])


@pytest.mark.parametrize("window", [5, 7, 9, 11, 13])
@pytest.mark.parametrize("erase_ground", [1, 0])
def test_erase_bfs_kernel_sizes_and_complement(O, window, erase_ground):
    p = P.preset(P.GS_PRESET_HDL32).copy(window_size=window, erase_ground=erase_ground)
    node = O.RefAlfaGS(p)
    pts = synth.make_frame(p, synth.SEED0 + 33)
    r = O.create_range_image(p, pts)
    a = O.create_angle_image(p, r)
    want, _, _ = O.erase_bfs(p, r, a)
    assert np.array_equal(node.erase_bfs(r, a, p.ground_angle_threshold, p.start_angle_threshold, window), want)


@pytest.mark.parametrize("ground_deg", [2, 5, 10, 15, 20])
def test_erase_bfs_threshold_sweep_with_smoothed_angles(O, ground_deg):
    # BASELINE config C5's thresholds; a smoothed angle image makes the start gate (alfa_gs.cpp:733) bite
    p = P.preset(P.GS_PRESET_VLP16)
    node = O.RefAlfaGS(p)
    pts = synth.make_frame(p, synth.SEED0 + 40 + ground_deg)
    r = O.create_range_image(p, pts)
    a = O.sg_smooth(p, O.create_angle_image(p, r), 7)
    assert np.array_equal(node.sg_smooth(O.create_angle_image(p, r), 7), a)
    g = float(np.radians(ground_deg))
    for start in (0.523598, 0.05):
        want, _, _ = O.erase_bfs(p, r, a, ground_thr=g, start_thr=start)
        assert np.array_equal(node.erase_bfs(r, a, g, start, p.window_size), want)


# ---- whole frames, the way ROS drives the node ----------------------------------------------------------------------
@pytest.mark.parametrize("name", list(ALFA_SHAPES))
def test_alfa_gs_whole_frame_matches_oracle(O, name):
    p = P.preset(ALFA_SHAPES[name])
    node = O.RefAlfaGS(p)
    for k, kw in enumerate((dict(), dict(shuffle=True))):
        pts = synth.make_frame(p, synth.SEED0 + 50 + k, **kw)
        got = node.process_frame(pts)  # PointCloud2 -> AlfaNode::cloud_cb -> Alfa_GS::process_pointcloud -> publish_*
        want = O.process_frame(p, pts)
        assert np.array_equal(got["no_ground"], want["no_ground"])
        assert np.array_equal(got["xyz"], want["xyz"])
        # the published cloud is pcl::PointXYZRGB: pad 1.0, rgba 0, 0, 0, 255
        rec = got["records"].view(np.uint8).reshape(-1, 32)
        assert np.all(rec[:, 16:19] == 0) and np.all(rec[:, 19] == 255)
    assert O.ref_assert_failures() == 0


@pytest.mark.parametrize("seed", range(6))
def test_random_small_sensors_whole_frame(O, seed):
    rng = np.random.default_rng(900 + seed)
    H, W = int(rng.choice([8, 16, 24, 40])), int(rng.choice([60, 128, 250, 512]))
    lo = -float(rng.uniform(0.2, 0.6))
    p = _small(H, W, min_angle=lo, max_angle=float(rng.uniform(0.05, 0.4)), max_range=float(rng.choice([50.0, 100.0, 120.0])),
               ground_angle_threshold=float(rng.uniform(0.03, 0.3)), window_size=int(rng.choice([5, 7, 9, 11])))
    assert p.shape == (H, W)
    node = O.RefAlfaGS(p)
    pts = synth.make_frame(p, 77 + seed, dropout=float(rng.uniform(0, 0.4)), jitter=0.45, shuffle=bool(seed & 1))
    got = node.process_frame(pts)
    want = O.process_frame(p, pts)
    assert np.array_equal(got["no_ground"], want["no_ground"])
    assert np.array_equal(got["xyz"], want["xyz"])


# ---- the two-node packaging: cloud2range + range2cloud --------------------------------------------------------------
def _c2r_params(preset):
    base = P.preset(preset)
    return base.copy(variant=P.GS_VARIANT_CLOUD2RANGE, smoothing=P.GS_SMOOTH_SG, repair_gaps=1)


@pytest.mark.parametrize("preset", [P.GS_PRESET_CLOUD2RANGE_DEFAULT, P.GS_PRESET_VLP16, P.GS_PRESET_HDL64E_2048])
def test_cloud2range_projection_loop(O, preset):
    # a start threshold of -1 rejects every seed: nothing is labelled and the node publishes the plain range image of
    # its projection loop (cloud2range_node.cpp:396-455)
    p = _c2r_params(preset)
    node = O.RefCloud2Range(p)
    node.set_thresholds(p.ground_angle_threshold, -1.0)
    for kw in (dict(), dict(shuffle=True, jitter=0.45)):
        pts = synth.make_frame(p, synth.SEED0 + 60, **kw)
        img, K = node.cloud_cb(pts)
        if O.ref_assert_failures():  # ROS_ASSERT(row < n_beams_ / col < n_cols_) aborts the reference (Q4): contract = drop
            pytest.skip("frame trips the reference's own assertions")
        assert np.array_equal(img, O.create_range_image(p, pts))


@pytest.mark.parametrize("preset", [P.GS_PRESET_CLOUD2RANGE_DEFAULT, P.GS_PRESET_VLP16, P.GS_PRESET_HDL64E_2048])
def test_cloud2range_whole_callback_and_range2cloud(O, preset):
    # Cloud2RangeNode::CloudCb (:381-480): projection -> RepairGaps (typed reading, Q22) -> CreateAngleImg -> SG(window) ->
    # EraseGroundBFS -> publish; then Range2CloudNode::CameraCb (range2cloud_node.cpp:36-99) on what was published
    p = _c2r_params(preset)
    node = O.RefCloud2Range(p)
    back = O.RefRange2Cloud()
    pts = synth.make_frame(p, synth.SEED0 + 61)
    img, K = node.cloud_cb(pts)
    if O.ref_assert_failures():
        pytest.skip("frame trips the reference's own assertions")
    want = O.process_frame(p, pts)
    assert np.array_equal(img, want["no_ground"])
    d = O.derive(p)
    assert K[:6] == [p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth, d.d_altitude]
    cloud = back.camera_cb(img, K)
    assert np.array_equal(cloud, want["xyz"])
    # stage by stage
    r = O.create_range_image(p, pts)
    rep = node.repair_gaps(r)
    assert np.array_equal(rep, O.repair_gaps(p, r))
    assert (rep != r).any(), "the frame should exercise the in-painting"
    ang = node.create_angle_img(rep)
    assert np.array_equal(ang, O.create_angle_image(p, rep))
    sm = node.sg_smooth(ang, p.window_size)
    assert np.array_equal(sm, O.sg_smooth(p, ang, p.window_size))
    res, _, _ = O.erase_bfs(p, r, sm)
    assert np.array_equal(node.erase_ground_bfs(r, sm, p.ground_angle_threshold, p.start_angle_threshold, p.window_size), res)


def test_range2cloud_with_foreign_camera_info(O):
    # K[4] / K[5] are taken as published, not re-derived (range2cloud_node.cpp:51-58)
    p = P.preset(P.GS_PRESET_VLP16).copy(variant=P.GS_VARIANT_CLOUD2RANGE)
    pts = synth.make_frame(p, synth.SEED0 + 62)
    img = O.process_frame(p, pts)["no_ground"]
    d = O.derive(p)
    K = [p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth * (1 + 3e-16), d.d_altitude * (1 - 2e-16)]
    cloud = O.RefRange2Cloud().camera_cb(img, K)
    rows, cols = np.nonzero(img)
    rng = (img[rows, cols].astype(np.float64) - 1) / 65534 * (K[3] - K[2]) + K[2]
    alt = (img.shape[0] - rows) * K[5] + K[0]
    az = cols * K[4]
    want = np.stack([np.cos(alt) * np.cos(az) * rng, np.cos(alt) * np.sin(az) * rng, np.sin(alt) * rng], 1).astype(np.float32)
    assert np.array_equal(cloud[:, :3], want)


# ---- N3: the evaluation members ------------------------------------------------------------------------------------------
def test_labeled_range_image_and_metrics(O):
    p = P.preset(P.GS_PRESET_VLP16)
    node = O.RefAlfaGS(p)
    pts = synth.make_frame(p, synth.SEED0 + 70)
    rng = np.random.default_rng(5)
    ground = pts[:, 2] < -1.5
    palette = np.array([[75, 0, 75], [255, 0, 255], [255, 150, 255]], np.uint8)  # alfa_gs.cpp:960
    rgb = np.where(ground[:, None], palette[rng.integers(0, 3, size=len(pts))], np.array([[10, 200, 30]], np.uint8)).astype(np.uint8)
    labeled_ref = node.labeled_range_image(pts, rgb)  # CreateLabeledRangeImage, alfa_gs.cpp:943-995
    rec = np.zeros((len(pts), 32), np.uint8)
    rec[:, :12] = pts[:, :3].copy().view(np.uint8).reshape(-1, 12)
    rec[:, 16], rec[:, 17], rec[:, 18], rec[:, 19] = rgb[:, 2], rgb[:, 1], rgb[:, 0], 255  # pcl: b, g, r, a
    labeled = O.create_labeled_range_image(p, rec)
    assert np.array_equal(labeled_ref, labeled)
    fr = O.process_frame(p, pts)
    ev = O.evaluate(p, fr["range"], fr["no_ground"], labeled)
    lines = node.check_detected(fr["range"], fr["no_ground"], labeled)  # CheckNumberOfDetectedRIdGnd, :831-941
    first = [l for l in lines if l.startswith("IoU ->")][-1]
    iou, rg, rmo = (float(x) for x in re.findall(r"-> ([-0-9.naif]+)", first))
    for got, want in ((iou, ev.iou_g), (rg, ev.recall_g), (rmo, ev.recall_mo)):
        assert abs(got - want) <= 5e-6 * max(1.0, abs(want))  # "%f" prints six decimals
    text = "\n".join(lines)
    assert int(re.search(r"true gnd points present in the range image -> (\d+)", text).group(1)) == ev.labeled_gnd
    assert int(re.search(r"number of gnd points detected in the range image -> (\d+)", text).group(1)) == ev.seg_gnd
