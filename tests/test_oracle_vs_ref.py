"""The oracle's restatement of Labeler/label.{h,cpp} and lib/utils.h against the reference's OWN sources,
compiled verbatim from /root/reference into oracle/_ref/libgs_ref.so (oracle/build.py, logical cv::Mat
shim).  Also the union-find reformulation (tests/uf_model.py) against both."""
import math

import numpy as np
import pytest

from groundsegmentationrosnode_b200 import params as P, synth

from . import uf_model


@pytest.fixture(scope="module")
def ref(oracle):
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return r


def _small_params(H, W):
    return P.GsParams(n_beams=H, rpm=60, sample_freq=W, min_angle=-0.4, max_angle=0.1, min_range=0.5, max_range=100.0,
                      ground_angle_threshold=0.115, start_angle_threshold=0.523598, window_size=11, erase_ground=1)


def _random_images(rng, H, W, hole=0.2, amax=40):
    rng_img = rng.integers(1, 65535, size=(H, W)).astype(np.uint16)
    rng_img[rng.random((H, W)) < hole] = 0
    # piecewise-smooth angles so that components are neither trivial nor everything
    base = np.cumsum(rng.integers(-3, 4, size=(H, W)), axis=1) + rng.integers(0, amax, size=(H, 1))
    ang = np.clip(base, 0, 157).astype(np.uint16)
    return rng_img, ang


def test_utils_match_reference(oracle, ref):
    rng = np.random.default_rng(1)
    L = oracle.lib()
    pts = rng.normal(0, 20, size=(2000, 3))
    pts[:10] = [[1, 0, 0], [-1, 0, 0], [-1, -0.0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1], [1e-30, 1e-30, 0],
                [3, -4, 0], [-3, -4, 12]]
    for x, y, z in pts:
        assert L.gso_point_range(x, y, z) == ref.gsref_point_range(x, y, z)
        assert L.gso_point_azimuth(x, y) == ref.gsref_point_azimuth(x, y)
        assert L.gso_point_altitude(x, y, z) == ref.gsref_point_altitude(x, y, z)
    for deg in (360.0 / 1800, 360.0 / 2048, 0.2, 1.0):
        assert L.gso_rad_deg(deg) == ref.gsref_rad_deg(deg)


def test_wrap_cols(ref):
    for W in (8, 1800, 2048):
        assert ref.gsref_wrap_cols(4, W, -1) == W - 1
        assert ref.gsref_wrap_cols(4, W, W) == 0
        assert ref.gsref_wrap_cols(4, W, 5) == 5


@pytest.mark.parametrize("seed", range(12))
def test_label_one_component_matches_reference(oracle, ref, seed):
    rng = np.random.default_rng(100 + seed)
    H, W = int(rng.integers(3, 14)), int(rng.integers(4, 30))
    p = _small_params(H, W)
    rimg, ang = _random_images(rng, H, W, hole=float(rng.uniform(0, 0.5)))
    lab0 = (rng.random((H, W)) < 0.1).astype(np.uint8) * 7  # pre-existing labels act as barriers
    for _ in range(6):
        r, c = int(rng.integers(0, H)), int(rng.integers(0, W))
        thr = float(rng.choice([0.02, 0.05, 0.09, 0.115, 0.2]))
        label = int(rng.integers(1, 255))
        want = lab0.copy()
        ref.gsref_label_one_component(H, W, label, r, c, thr, rimg, ang, want)
        got = oracle.label_one_component(p, label, r, c, thr, rimg, ang, lab0)
        assert np.array_equal(got, want)
        lab0 = got  # accumulate, like successive seeds in EraseBFS


@pytest.mark.parametrize("preset,seed", [(P.GS_PRESET_VLP16, 0), (P.GS_PRESET_VLP16, 1), (P.GS_PRESET_HDL64E_2048, 2)])
def test_ground_labels_match_reference_on_scans(oracle, ref, preset, seed):
    p = P.preset(preset)
    H, W = p.shape
    pts = synth.make_frame(p, synth.SEED0 + seed)
    rimg = oracle.create_range_image(p, pts)
    ang = oracle.create_angle_image(p, rimg)
    _, lab, _ = oracle.erase_bfs(p, rimg, ang)
    want = np.zeros((H, W), np.uint8)
    ref.gsref_ground_labels(H, W, p.ground_angle_threshold, p.start_angle_threshold, rimg, ang, want)
    assert np.array_equal(lab, (want != 0).astype(np.uint8))
    d = oracle.derive(p)
    assert np.array_equal(uf_model.ground_labels(rimg, ang, d.k_ground, d.s_start), lab)


@pytest.mark.parametrize("seed", range(20))
def test_union_find_model_equals_bfs(oracle, seed):
    """BFS == union-find reformulation on random images, including non-zero angles at empty pixels (the
    smoothed-image case), empty columns (empty seeds, Q7) and wrap-around columns."""
    rng = np.random.default_rng(500 + seed)
    H, W = int(rng.integers(2, 16)), int(rng.integers(3, 40))
    p = _small_params(H, W)
    rimg, ang = _random_images(rng, H, W, hole=float(rng.uniform(0, 0.7)), amax=70)
    if seed % 3 == 0:
        rimg[:, rng.integers(0, W)] = 0  # a fully empty column
    if seed % 2 == 0:
        ang[rimg == 0] = 0  # un-smoothed angle images hold 0 at empty pixels
    thr = float(rng.choice([0.03, 0.09, 0.115, 0.3]))
    start = float(rng.choice([0.1, 0.523598]))
    _, lab, _ = oracle.erase_bfs(p, rimg, ang, ground_thr=thr, start_thr=start)
    K = oracle.lib().gso_k_of_threshold(thr)
    S = oracle.lib().gso_s_of_threshold(start)
    assert np.array_equal(uf_model.ground_labels(rimg, ang, K, S), lab)


@pytest.mark.parametrize("seed", range(8))
def test_all_components_model(oracle, seed):
    rng = np.random.default_rng(900 + seed)
    H, W = int(rng.integers(2, 16)), int(rng.integers(3, 40))
    p = _small_params(H, W)
    rimg, ang = _random_images(rng, H, W, hole=float(rng.uniform(0, 0.6)))
    beta = math.radians(float(rng.choice([5, 10, 15, 20])))
    got, n = oracle.label_components(p, rimg, ang, beta)
    K = oracle.lib().gso_k_of_threshold(beta)
    want = uf_model.components(rimg != 0, ang, K)
    assert np.array_equal(got, want)
    assert n == len(set(want[want >= 0].tolist()))
