"""GPU parity: the CUDA path, called through the C ABI (libgs_b200.so), against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star):
  * range-image pixel assignment and ground / no-ground images: bit-exact, except pixels that trace back to a point
    within 1e-5 rad of a cell edge (projection) or an inclination angle within 1e-5 rad of a centi-radian edge —
    those are counted (gs_stats / oracle counters) and bounded here;
  * component labels: equal after canonical min-index relabelling (both sides emit the canonical form);
  * back-projected XYZ: relative error <= 1e-6 (TOL_XYZ); the implementation is in fact bit-exact because cos/sin
    come from host-computed double tables.
"""
import math

import numpy as np
import pytest

from groundsegmentationrosnode_b200 import capi, params as P, synth

pytestmark = pytest.mark.gpu

TOL_XYZ = 1e-6

CONFIGS = {
    "C1_vlp16": P.GS_PRESET_VLP16,
    "C2_hdl64_2048": P.GS_PRESET_HDL64E_2048,
    "C3_os1_128": P.GS_PRESET_OS1_128,
    "alfa_default_1800": P.GS_PRESET_ALFA_GS_DEFAULT,
    "cloud2range_sg7": P.GS_PRESET_CLOUD2RANGE_DEFAULT,
    "hdl32": P.GS_PRESET_HDL32,
}


@pytest.fixture(scope="module")
def ctxs():
    cache = {}

    def get(p, staged=False, stats=True):
        # stats=True: GS_OPT_COLLECT_STATS selects the CHECKED instantiations of the fused kernel (the exact arithmetic is evaluated
        # beside every fast result and disagreements are counted); stats=False is the PRODUCTION instantiation bench.py times.
        # A call that brings one frame spreads it over the widest cluster that fits (the generic instantiation, lowest latency):
        # mode "lone".  The other fused modes pin single frames to the batch plan's cluster, i.e. to the instantiation batches run.
        lone = staged == "lone"
        if staged in ("production", "lone"):
            staged, stats = False, False
        key = (bytes(p), staged, stats, lone)
        if key not in cache:
            c = capi.Context(p)
            c.set_option(capi.GS_OPT_COLLECT_STATS, int(stats))
            c.set_option(capi.GS_OPT_FORCE_STAGED, int(staged))
            c.set_option(capi.GS_OPT_NO_LONE_PLAN, int(not lone))
            cache[key] = c
        return cache[key]

    yield get
    for c in cache.values():
        c.close()


def _mismatch_budget(near_edge):
    # every near-edge point / pixel may flip one pixel and, through the labelling, its neighbourhood; on the
    # synthetic scans the count of actual flips is 0 — the budget only documents the contract
    return near_edge


def _assert_images_equal(got, want, what, budget=0):
    bad = int(np.count_nonzero(got != want))
    assert bad <= budget, "%s: %d pixels differ (budget %d)" % (what, bad, budget)
    return bad


def _xyz_close(got, want):
    assert got.shape == want.shape
    if got.size == 0:
        return
    assert np.array_equal(got[:, 3], want[:, 3])
    num = np.abs(got[:, :3].astype(np.float64) - want[:, :3].astype(np.float64))
    den = np.maximum(np.abs(want[:, :3].astype(np.float64)), 1e-3)  # absolute floor near sin ~ 0
    assert float((num / den).max()) <= TOL_XYZ


MODES = [False, True, "production", "lone"]
MODE_IDS = ["fused_checked", "staged", "fused_production", "fused_lone_frame_plan"]


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("staged", MODES, ids=MODE_IDS)
def test_full_pipeline_matches_oracle(ctxs, oracle, name, staged):
    p = P.preset(CONFIGS[name])
    ctx = ctxs(p, staged)
    production = staged in ("production", "lone")
    for seed in range(2):
        pts = synth.make_frame(p, synth.SEED0 + seed, shuffle=(seed == 1))
        want = oracle.process_frame(p, pts)
        ctx.reset_stats()
        launches0 = ctx.stats().kernel_launches
        got = ctx.process_frame(pts)
        st = ctx.stats()
        # every preset here is one the fused kernel covers: the whole frame is ONE launch; the staged path is 9-10
        assert (st.kernel_launches - launches0 == 1) == (staged is not True)
        _, ne_p, _ = oracle.create_range_image(p, pts, with_counts=True)
        if not production:
            assert st.near_edge_project == ne_p
        # with statistics on, the fused kernel evaluates the exact arithmetic beside every fast result it accepts
        assert st.fast_path_mismatches == 0
        if staged is False:
            assert st.slow_path_points < 0.02 * (pts.shape[0] + p.n_beams * p.n_cols)
        bad = _assert_images_equal(got["range"], want["range"], "range image", _mismatch_budget(ne_p))
        if bad == 0:
            _assert_images_equal(got["angle"], want["angle"], "angle image")
            _assert_images_equal(got["no_ground"], want["no_ground"], "no-ground image")
            _xyz_close(got["xyz"], want["xyz"])
            assert np.array_equal(got["xyz"], want["xyz"])  # stronger than the bar: bit-exact


@pytest.mark.parametrize("name", ["C1_vlp16", "C2_hdl64_2048", "cloud2range_sg7"])
def test_stage_api_matches_oracle(ctxs, oracle, name):
    p = P.preset(CONFIGS[name])
    ctx = ctxs(p)
    pts = synth.make_frame(p, synth.SEED0 + 7)
    rng = ctx.project(pts)
    assert np.array_equal(rng, oracle.create_range_image(p, pts))
    # 32-byte pcl::PointXYZRGB records (the layout Alfa_GS::CreateRangeImage reads, alfa_gs.h:25)
    wide = np.zeros((pts.shape[0], 8), np.float32)
    wide[:, :3] = pts[:, :3]
    wide[:, 3] = 1.0
    assert np.array_equal(ctx.project(wide), rng)
    ang = ctx.angle_image(rng)
    assert np.array_equal(ang, oracle.create_angle_image(p, rng))
    for ws in (5, 7, 9, 11):
        assert np.array_equal(ctx.smooth(ang, ws), oracle.sg_smooth(p, ang, ws))
    with pytest.raises(capi.GsError):
        ctx.smooth(ang, 13)  # alfa_gs.cpp:645-649 throws std::logic_error("bad window size")
    rep = ctx.repair_gaps(rng)
    assert np.array_equal(rep, oracle.repair_gaps(p, rng))
    for ground, start, ks in ((None, None, None), (0.05, 0.523598, 7), (0.2, 0.1, 5), (0.09, 0.3, 9), (0.115, 0.523598, 13)):
        res, lab = ctx.erase_ground(rng, ang, ground, start, ks)
        wres, wlab, _ = oracle.erase_bfs(p, rng, ang, ground, start, ks)
        assert np.array_equal(lab, wlab)
        assert np.array_equal(res, wres)
    with pytest.raises(capi.GsError):
        ctx.erase_ground(rng, ang, 0.1, 0.5, 8)  # even kernel: std::logic_error in the reference
    res, _ = ctx.erase_ground(rng, ang)
    xyz = ctx.backproject(res)
    assert np.array_equal(xyz, oracle.camera_cb(p, res))
    xyz_full = ctx.backproject(rng)
    assert np.array_equal(xyz_full, oracle.camera_cb(p, rng))


def test_golden_opencv_vectors_on_gpu(ctxs):
    """The committed OpenCV vectors (tests/golden) straight against the CUDA kernels: SG filter2D and, through
    gs_erase_ground's label + result pair, the effective dilation taps."""
    import os

    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_golden.npz"))
    for ws in (int(w) for w in G["sg_ws"]):
        for tag in "abc":
            img = G["sg_in_%d_%s" % (ws, tag)]
            H, W = img.shape
            p = P.GsParams(n_beams=H, rpm=60, sample_freq=W, min_angle=-0.1, max_angle=0.1, min_range=0.5, max_range=100.0,
                           ground_angle_threshold=0.1, start_angle_threshold=0.5, window_size=ws, erase_ground=1)
            with capi.Context(p) as c:
                assert np.array_equal(c.smooth(img, ws), G["sg_out_%d_%s" % (ws, tag)])
    for ks in (int(k) for k in G["dil_ks"]):
        lab = G["dil_in_%d" % ks]
        H, W = lab.shape
        p = P.GsParams(n_beams=H, rpm=60, sample_freq=W, min_angle=-0.1, max_angle=0.1, min_range=0.5, max_range=100.0,
                       ground_angle_threshold=0.1, start_angle_threshold=0.5, window_size=ks + 2, erase_ground=1)
        d = capi.derive(p)
        offs = list(d.dilate_off)[: d.n_dilate]
        dil = np.zeros_like(lab)
        for o in offs:
            src = np.zeros_like(lab)
            if o < 0:
                src[-o:, :] = lab[:o, :]
            elif o > 0:
                src[:-o, :] = lab[o:, :]
            else:
                src = lab
            dil |= src
        assert np.array_equal(dil, G["dil_out_%d" % ks]), ks


@pytest.mark.parametrize("kind", ["edges", "collisions", "nonfinite", "sparse"])
@pytest.mark.parametrize("name", ["C1_vlp16", "C2_hdl64_2048"])
@pytest.mark.parametrize("staged", MODES, ids=MODE_IDS)
def test_adversarial_frames(ctxs, oracle, name, kind, staged):
    p = P.preset(CONFIGS[name])
    ctx = ctxs(p, staged)
    pts = synth.make_adversarial_frame(p, synth.SEED0 + 3, kind)
    want = oracle.process_frame(p, pts)
    wrng, ne_p, dropped = oracle.create_range_image(p, pts, with_counts=True)
    ctx.reset_stats()
    got = ctx.process_frame(pts)
    st = ctx.stats()
    if staged not in ("production", "lone"):
        assert st.dropped_points == dropped
    assert st.fast_path_mismatches == 0
    budget = ne_p if kind == "edges" else 0
    bad = _assert_images_equal(got["range"], want["range"], "range image", budget)
    if bad:
        # points exactly on cell edges: double atan2/hypot of the device and of glibc may differ in the last ulp and
        # move such a point to the adjacent cell (the north-star tolerance set).  Re-anchor the later stages on the
        # oracle's range image.
        ang = ctx.angle_image(want["range"])
        assert np.array_equal(ang, want["angle"])
        res, _ = ctx.erase_ground(want["range"], ang)
        assert np.array_equal(res, want["no_ground"])
        assert np.array_equal(ctx.backproject(res), want["xyz"])
    else:
        assert np.array_equal(got["angle"], want["angle"])
        assert np.array_equal(got["no_ground"], want["no_ground"])
        assert np.array_equal(got["xyz"], want["xyz"])


@pytest.mark.parametrize("policy", ["nearest", "wrap", "keep_ground"])
def test_policy_variants(ctxs, oracle, policy):
    base = P.preset(P.GS_PRESET_HDL64E_2048)
    p = {"nearest": base.copy(pixel_policy=P.GS_PIXEL_NEAREST), "wrap": base.copy(col_policy=P.GS_COL_WRAP),
         "keep_ground": base.copy(erase_ground=0)}[policy]
    for staged in MODES:
        ctx = ctxs(p, staged)
        pts = synth.make_adversarial_frame(p, synth.SEED0 + 11, "collisions")
        want = oracle.process_frame(p, pts)
        got = ctx.process_frame(pts)
        for k in ("range", "angle", "no_ground"):
            assert np.array_equal(got[k], want[k]), k
        assert np.array_equal(got["xyz"], want["xyz"])


def test_empty_and_tiny_inputs(ctxs, oracle):
    p = P.preset(P.GS_PRESET_VLP16)
    for staged in MODES:
        ctx = ctxs(p, staged)
        for pts in (np.zeros((0, 4), np.float32), np.array([[5.0, 0.1, -0.3, 0.0]], np.float32),
                    np.array([[5.0, 0.0, 0.0, 0.0], [5.0, 0.0, 0.0, 0.0]], np.float32)):
            want = oracle.process_frame(p, pts)
            got = ctx.process_frame(pts)
            for k in ("range", "angle", "no_ground"):
                assert np.array_equal(got[k], want[k]), k
            assert np.array_equal(got["xyz"], want["xyz"])


@pytest.mark.parametrize("seed", range(6))
def test_label_one_component_matches_oracle(ctxs, oracle, seed):
    """Label::LabelOneComponent through the C ABI, including pre-labelled barrier pixels and empty seeds."""
    rng = np.random.default_rng(4000 + seed)
    H, W = int(rng.integers(3, 20)), int(rng.integers(4, 60))
    p = P.GsParams(n_beams=H, rpm=60, sample_freq=W, min_angle=-0.4, max_angle=0.1, min_range=0.5, max_range=100.0,
                   ground_angle_threshold=0.115, start_angle_threshold=0.523598, window_size=11, erase_ground=1)
    ctx = ctxs(p)
    rimg = rng.integers(1, 65535, size=(H, W)).astype(np.uint16)
    rimg[rng.random((H, W)) < 0.3] = 0
    ang = np.clip(np.cumsum(rng.integers(-3, 4, size=(H, W)), axis=1) + rng.integers(0, 40, size=(H, 1)), 0, 157).astype(np.uint16)
    lab = (rng.random((H, W)) < 0.1).astype(np.uint8) * 9
    for _ in range(5):
        r, c = int(rng.integers(0, H)), int(rng.integers(0, W))
        thr = float(rng.choice([0.02, 0.05, 0.09, 0.115, 0.2]))
        label = int(rng.integers(1, 255))
        want = oracle.label_one_component(p, label, r, c, thr, rimg, ang, lab)
        got = ctx.label_one_component(label, r, c, thr, rimg, ang, lab)
        assert np.array_equal(got, want)
        lab = got


@pytest.mark.parametrize("beta_deg", [5, 10, 15, 20])
def test_label_components_sweep(ctxs, oracle, beta_deg):
    """BASELINE config C5: the Labeler as an all-components CCL on 128 x 2048 no-ground images."""
    p = P.preset(P.GS_PRESET_OS1_128)
    ctx = ctxs(p)
    pts = synth.make_frame(p, synth.SEED0 + 21)
    o = oracle.process_frame(p, pts)
    beta = math.radians(beta_deg)
    want, n_want = oracle.label_components(p, o["no_ground"], o["angle"], beta)
    got, n_got = ctx.label_components(o["no_ground"], o["angle"], beta)
    assert n_got == n_want
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ground_deg", [2, 4, 6, 8, 10])
def test_ground_threshold_sweep(ctxs, oracle, ground_deg):
    p = P.preset(P.GS_PRESET_OS1_128).copy(ground_angle_threshold=math.radians(ground_deg))
    for staged in MODES:
        ctx = ctxs(p, staged)
        pts = synth.make_frame(p, synth.SEED0 + 22)
        want = oracle.process_frame(p, pts)
        got = ctx.process_frame(pts)
        assert np.array_equal(got["no_ground"], want["no_ground"])
        assert np.array_equal(got["xyz"], want["xyz"])


@pytest.mark.parametrize("staged", MODES, ids=MODE_IDS)
def test_batch_device_and_host_match_per_frame_oracle(ctxs, oracle, staged):
    import torch

    p = P.preset(P.GS_PRESET_HDL64E_2048)
    ctx = ctxs(p, staged)
    H, W = ctx.shape
    Pn = H * W
    pool, offs = synth.make_pool(p, 5, seed0=synth.SEED0 + 40)
    pts, offsets = synth.tile_pool(pool, offs, 13)  # ragged frame sizes, more frames than distinct
    F = len(offsets) - 1
    want = [oracle.process_frame(p, pool[offs[i]:offs[i + 1]]) for i in range(5)]
    d_pts = torch.from_numpy(pts).cuda()
    d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_rng = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_ang = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_xyz = torch.zeros((F, Pn, 4), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.set_chunk_frames(4)  # several waves, a ragged last one
    ctx.process_batch_device(d_pts.data_ptr(), offsets, F, d_rng.data_ptr(), d_ang.data_ptr(), d_res.data_ptr(), d_xyz.data_ptr(),
                             d_cnt.data_ptr())
    ctx.synchronize()
    ctx.set_chunk_frames(0)
    res = d_res.cpu().numpy().view(np.uint16)
    rngs = d_rng.cpu().numpy().view(np.uint16)
    angs = d_ang.cpu().numpy().view(np.uint16)
    cnt = d_cnt.cpu().numpy()
    xyz = d_xyz.cpu().numpy()
    for f in range(F):
        w = want[f % 5]
        assert np.array_equal(rngs[f], w["range"])
        assert np.array_equal(angs[f], w["angle"])
        assert np.array_equal(res[f], w["no_ground"])
        assert cnt[f] == w["xyz"].shape[0]
        assert np.array_equal(xyz[f, : cnt[f]], w["xyz"])
    # host batch: packed clouds
    h_res = np.zeros((F, H, W), np.uint16)
    h_xyz = np.zeros((F * Pn, 4), np.float32)
    ctx.set_chunk_frames(5)
    _, _, out_offs = ctx.process_batch_host(pts, offsets, h_res, h_xyz)
    ctx.set_chunk_frames(0)
    for f in range(F):
        w = want[f % 5]
        assert np.array_equal(h_res[f], w["no_ground"])
        assert np.array_equal(h_xyz[out_offs[f]:out_offs[f + 1]], w["xyz"])


@pytest.mark.parametrize("preset", [P.GS_PRESET_VLP16, P.GS_PRESET_HDL64E_2048, P.GS_PRESET_OS1_128])
def test_fused_union_paths_agree_with_oracle(oracle, preset):
    """The fused kernel has two union paths: warp-synchronous without atomics (the link list fits in shared memory) and
    compare-and-swap on chunks of the list (it does not).  Both must give the oracle's no-ground image on a batch."""
    import torch

    p = P.preset(preset)
    H, W = p.shape
    pool, offs = synth.make_pool(p, 6, seed0=synth.SEED0 + 900)
    pts, offsets = synth.tile_pool(pool, offs, 40)
    F = len(offsets) - 1
    want = [oracle.process_frame(p, pool[offs[i]:offs[i + 1]]) for i in range(6)]
    d_pts = torch.from_numpy(pts).cuda()
    for cas in (0, 1):
        with capi.Context(p) as ctx:
            ctx.set_option(capi.GS_OPT_CAS_UNIONS, cas)
            d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
            d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            l0 = ctx.stats().kernel_launches
            ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), 0, d_cnt.data_ptr())
            ctx.synchronize()
            # one launch of the cluster kernel, plus one of its software-cluster twin when SMs are left over (hybrid mode)
            assert 1 <= ctx.stats().kernel_launches - l0 <= 2, "the fused kernel should have taken this batch"
            res = d_res.cpu().numpy().view(np.uint16)
            for f in range(F):
                assert np.array_equal(res[f], want[f % 6]["no_ground"]), "frame %d, cas=%d" % (f, cas)
                assert int(d_cnt[f]) == want[f % 6]["xyz"].shape[0]


@pytest.mark.parametrize("noise_sigma,dropout", [(0.0, 0.0), (0.0, 0.3), (0.002, 0.0), (0.08, 0.05), (0.25, 0.0)])
def test_union_stage_stress_frames(oracle, noise_sigma, dropout):
    """Frames that stress the labelling stage of the fused kernel: without range noise the angle image is smooth, the
    ground is a handful of giant components and almost every link of a warp's batch hangs the same root (maximal contention
    in the warp-synchronous unions); with heavy noise nearly every pixel is a run of its own (maximal link and run counts,
    the list spills into the compare-and-swap path for some frames).  Both union paths, bit-exact against the oracle."""
    import torch

    p = P.preset(P.GS_PRESET_HDL64E_2048)
    H, W = p.shape
    frames = [synth.make_frame(p, synth.SEED0 + 1200 + k, dropout=dropout, noise_sigma=noise_sigma) for k in range(3)]
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
    pool = np.concatenate(frames)
    pts, offsets = synth.tile_pool(pool, offs, 36)  # more frames than co-resident clusters
    F = len(offsets) - 1
    want = [oracle.process_frame(p, f) for f in frames]
    d_pts = torch.from_numpy(pts).cuda()
    for cas in (0, 1):
        with capi.Context(p) as ctx:
            ctx.set_option(capi.GS_OPT_CAS_UNIONS, cas)
            d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
            d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            for rep in range(2):  # races are timing dependent: run the batch twice
                ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), 0, d_cnt.data_ptr())
                ctx.synchronize()
                res = d_res.cpu().numpy().view(np.uint16)
                for f in range(F):
                    assert np.array_equal(res[f], want[f % 3]["no_ground"]), "frame %d, cas=%d, rep=%d" % (f, cas, rep)
                    assert int(d_cnt[f]) == want[f % 3]["xyz"].shape[0]


@pytest.mark.parametrize("preset", [P.GS_PRESET_CLOUD2RANGE_DEFAULT, P.GS_PRESET_HDL64E_2048, P.GS_PRESET_VLP16, P.GS_PRESET_OS1_128])
def test_repair_gaps_in_the_pipeline(oracle, preset):
    """N2 / O1': with repair_gaps on (the cloud2range default, as Cloud2RangeNode::CloudCb runs it), RepairGaps(range, 5, 1.0f)
    feeds the angle image only (cloud2range_node.cpp:468-473, typed version).  The fused kernel in-paints its range slab in shared
    memory; the per-stage kernels do the same in global memory.  Frame and batch entry points against the oracle."""
    import torch

    p = P.preset(preset).copy(repair_gaps=1)
    H, W = p.shape
    frames = [synth.make_frame(p, synth.SEED0 + 700 + k, dropout=0.15) for k in range(3)]
    want = [oracle.process_frame(p, f) for f in frames]
    assert not np.array_equal(want[0]["angle"], oracle.process_frame(p.copy(repair_gaps=0), frames[0])["angle"]), "RepairGaps had no effect"
    with capi.Context(p) as ctx:
        for stats, staged in ((1, 0), (0, 0), (0, 1)):
            ctx.set_option(capi.GS_OPT_COLLECT_STATS, stats)
            ctx.set_option(capi.GS_OPT_FORCE_STAGED, staged)
            for f, w in zip(frames, want):
                l0 = ctx.stats().kernel_launches
                got = ctx.process_frame(f)
                assert (ctx.stats().kernel_launches - l0 == 1) == (not staged), "RepairGaps runs inside the fused kernel"
                for k in ("range", "angle", "no_ground"):
                    assert np.array_equal(got[k], w[k]), (stats, staged, k)
                assert np.array_equal(got["xyz"], w["xyz"])
        ctx.set_option(capi.GS_OPT_FORCE_STAGED, 0)
        offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
        d_pts = torch.from_numpy(np.concatenate(frames)).cuda()
        d_res = torch.zeros((3, H, W), dtype=torch.int16, device="cuda")
        d_cnt = torch.zeros(3, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.process_batch_device(d_pts.data_ptr(), offs, 3, 0, 0, d_res.data_ptr(), 0, d_cnt.data_ptr())
        ctx.synchronize()
        for k in range(3):
            assert np.array_equal(d_res[k].cpu().numpy().view(np.uint16), want[k]["no_ground"])
            assert int(d_cnt[k]) == want[k]["xyz"].shape[0]


def test_frames_beyond_the_fused_kernels_point_limit(oracle):
    """A frame with more than 2^20 points (many returns per pixel: the last one wins) cannot take the fused kernel; the wave
    it is in runs on the per-stage kernels.  Same results, frame and batch entry points."""
    import torch

    p = P.preset(P.GS_PRESET_HDL64E_2048)
    H, W = p.shape
    small = synth.make_frame(p, synth.SEED0 + 800)
    big = np.concatenate([synth.make_frame(p, synth.SEED0 + 801 + k) for k in range(9)])
    assert big.shape[0] > (1 << 20)
    want_big, want_small = oracle.process_frame(p, big), oracle.process_frame(p, small)
    with capi.Context(p) as ctx:
        l0 = ctx.stats().kernel_launches
        got = ctx.process_frame(big)
        assert ctx.stats().kernel_launches - l0 > 1, "expected the per-stage kernels"
        for k in ("range", "angle", "no_ground"):
            assert np.array_equal(got[k], want_big[k]), k
        assert np.array_equal(got["xyz"], want_big["xyz"])
        offs = np.array([0, small.shape[0], small.shape[0] + big.shape[0]], np.int64)
        d_pts = torch.from_numpy(np.concatenate([small, big])).cuda()
        d_res = torch.zeros((2, H, W), dtype=torch.int16, device="cuda")
        d_xyz = torch.zeros((2, H * W, 4), dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.process_batch_device(d_pts.data_ptr(), offs, 2, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
        ctx.synchronize()
        for k, w in enumerate((want_small, want_big)):
            assert np.array_equal(d_res[k].cpu().numpy().view(np.uint16), w["no_ground"])
            n = int(d_cnt[k])
            assert n == w["xyz"].shape[0]
            assert np.array_equal(d_xyz[k, :n].cpu().numpy(), w["xyz"])


def test_set_params_recomputes_derived(ctxs, oracle):
    """gs_set_params replaces Alfa_GS::UpdateSegmentationSettings (alfa_gs.cpp:1122-1136) and, unlike it, refreshes
    the derived values."""
    p = P.preset(P.GS_PRESET_VLP16)
    with capi.Context(p) as ctx:
        q = P.preset(P.GS_PRESET_HDL32)
        ctx.set_params(q)
        assert ctx.shape == q.shape
        pts = synth.make_frame(q, synth.SEED0 + 5)
        want = oracle.process_frame(q, pts)
        got = ctx.process_frame(pts)
        assert np.array_equal(got["no_ground"], want["no_ground"])
        assert np.array_equal(got["xyz"], want["xyz"])


def test_full_size_properties(ctxs, oracle):
    """Size-independent properties on a C4-style batch (no oracle at this size): idempotence of the batch call,
    frame independence (a frame's result does not depend on its neighbours in the batch), counts == non-zero pixels,
    no-ground image is a subset of the range image, back-projection of the range image has one point per pixel."""
    import torch

    p = P.preset(P.GS_PRESET_HDL64E_2048)
    ctx = ctxs(p)
    H, W = ctx.shape
    Pn = H * W
    pool, offs = synth.make_pool(p, 4, seed0=synth.SEED0 + 60)
    F = 600  # several frames per resident cluster: the persistent kernel re-uses its shared-memory state across frames
    pts, offsets = synth.tile_pool(pool, offs, F)
    d_pts = torch.from_numpy(pts).cuda()
    outs = []
    for _ in range(2):
        d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
        d_rng = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
        d_xyz = torch.zeros((F, Pn, 4), dtype=torch.float32, device="cuda")
        d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.process_batch_device(d_pts.data_ptr(), offsets, F, d_rng.data_ptr(), 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
        ctx.synchronize()
        outs.append((d_res, d_rng, d_xyz, d_cnt))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)  # idempotent / deterministic
    d_res, d_rng, d_xyz, d_cnt = outs[0]
    for f in range(4, F):
        assert torch.equal(d_res[f], d_res[f % 4])  # tiled frames give identical results wherever they sit
        assert d_cnt[f] == d_cnt[f % 4]
    assert torch.equal((d_res != 0).sum(dim=(1, 2)).to(torch.int32), d_cnt)
    assert bool(((d_res == 0) | (d_res == d_rng)).all())
    n0 = int(d_cnt[0])
    assert bool((d_xyz[0, :n0, 3] == 1.0).all())
    for i in range(4):  # and the four distinct frames equal the oracle
        w = oracle.process_frame(p, pool[offs[i]:offs[i + 1]])
        assert np.array_equal(d_res[i].cpu().numpy().view(np.uint16), w["no_ground"])
        assert np.array_equal(d_xyz[i, : int(d_cnt[i])].cpu().numpy(), w["xyz"])


def test_wire_formats(ctxs, oracle):
    """N1: PointCloud2-style records with fields at arbitrary offsets in, PointXYZRGB records and
    CameraInfo out — what pcl::fromROSMsg / pcl::toROSMsg / the CameraInfo block do around the path
    (alfa_node.cpp:311, :35-41; alfa_gs.cpp:55-63)."""
    p = P.preset(P.GS_PRESET_HDL64E_2048)
    ctx = ctxs(p)
    pts = synth.make_frame(p, synth.SEED0 + 90)
    n = pts.shape[0]
    # a velodyne-driver style layout: x@0 y@4 z@8 intensity@16 ring(u16)@20, point_step 22 -> unaligned records
    for step, ox, oy, oz, oi, big in ((22, 0, 4, 8, 16, False), (32, 4, 12, 20, 0, False), (16, 8, 0, 4, 12, True)):
        raw = np.zeros((n, step), np.uint8)
        dt = ">f4" if big else "<f4"
        for col, off in ((0, ox), (1, oy), (2, oz), (3, oi)):
            raw[:, off:off + 4] = pts[:, col].astype(dt).view(np.uint8).reshape(n, 4)
        got = ctx.unpack_points(raw, step, ox, oy, oz, oi, big_endian=big)
        assert np.array_equal(got.view(np.uint32), pts.view(np.uint32))
        out = ctx.process_cloud2(raw, step, ox, oy, oz, big_endian=big)
        want = oracle.process_frame(p, pts)
        for k in ("range", "angle", "no_ground"):
            assert np.array_equal(out[k], want[k]), k
        assert np.array_equal(out["xyz"], want["xyz"])
    # 32-byte pcl::PointXYZRGB input through gs_process_frame takes the fused kernel as well (one unpack + one fused launch)
    wide = np.zeros((n, 8), np.float32)
    wide[:, :3] = pts[:, :3]
    l0 = ctx.stats().kernel_launches
    out = ctx.process_frame(wide)
    assert ctx.stats().kernel_launches - l0 == 2
    assert np.array_equal(out["no_ground"], want["no_ground"]) and np.array_equal(out["xyz"], want["xyz"])
    # cloud out: PointXYZRGB records as Alfa_GS::CameraCb builds them
    rec = ctx.pack_cloud_xyzrgb(want["xyz"])
    assert rec.shape == (want["xyz"].shape[0], 32)
    assert np.array_equal(rec[:, :12].copy().view(np.float32).reshape(-1, 3), want["xyz"][:, :3])
    assert np.array_equal(rec[:, 12:16].copy().view(np.float32).reshape(-1), np.ones(rec.shape[0], np.float32))
    assert np.all(rec[:, 16:19] == 0) and np.all(rec[:, 19] == 255) and np.all(rec[:, 20:] == 0)
    h, w, K = ctx.camera_info()
    d = oracle.derive(p)
    assert (h, w) == (64, 2048) and K[:6] == [p.min_angle, p.max_angle, p.min_range, p.max_range, d.d_azimuth, d.d_altitude]


@pytest.mark.parametrize("seed", range(16))
def test_random_sensor_configurations(oracle, seed):
    """Random sensor shapes and thresholds through the fused kernel (uneven last slab, odd widths, every cluster size the
    plan search picks, both variants, smoothing on/off, keep-ground mode, shuffled and sparse clouds) against the oracle."""
    rng = np.random.default_rng(7000 + seed)
    H = int(rng.choice([16, 32, 64, 128]))
    W = int(rng.choice([517, 900, 1000, 1024, 1500, 1800, 2000, 2048, 2500]))
    if H == 128:
        W = min(W, 2048)
    lo = -float(rng.uniform(0.2, 0.5))
    hi = float(rng.uniform(0.05, 0.4))
    variant = int(rng.integers(0, 2))
    smoothing = int(rng.integers(0, 2))
    ws = int(rng.choice([5, 7, 9, 11]))
    p = P.GsParams(n_beams=H, rpm=600, sample_freq=W * 10, min_angle=lo, max_angle=hi, min_range=float(rng.uniform(0.3, 1.0)),
                   max_range=float(rng.uniform(60, 150)), ground_angle_threshold=float(rng.uniform(0.03, 0.2)),
                   start_angle_threshold=float(rng.uniform(0.2, 0.8)), window_size=ws, erase_ground=int(rng.integers(0, 2)),
                   variant=variant, pixel_policy=P.GS_PIXEL_LAST, col_policy=int(rng.integers(0, 2)), smoothing=smoothing, repair_gaps=0)
    assert p.shape == (H, W)
    pts = synth.make_frame(p, synth.SEED0 + 500 + seed, dropout=float(rng.choice([0.0, 0.05, 0.3, 0.8])), shuffle=bool(rng.integers(0, 2)),
                           jitter=float(rng.choice([0.3, 0.45])))
    want = oracle.process_frame(p, pts)
    with capi.Context(p) as ctx:
        for stats in (1, 0):  # the checked instantiation on the batch plan's cluster, the production one on the lone-frame plan
            ctx.set_option(capi.GS_OPT_COLLECT_STATS, stats)
            ctx.set_option(capi.GS_OPT_NO_LONE_PLAN, stats)
            l0 = ctx.stats().kernel_launches
            got = ctx.process_frame(pts)
            st = ctx.stats()
            assert st.kernel_launches - l0 == 1, "the fused kernel should cover this configuration"
            assert st.fast_path_mismatches == 0
            for k in ("range", "angle", "no_ground"):
                assert np.array_equal(got[k], want[k]), (stats, k)
            assert np.array_equal(got["xyz"], want["xyz"])
        ctx.set_option(capi.GS_OPT_FORCE_STAGED, 1)
        got2 = ctx.process_frame(pts)
        for k in ("range", "angle", "no_ground"):
            assert np.array_equal(got2[k], want[k]), "staged " + k
        assert np.array_equal(got2["xyz"], want["xyz"])


def test_large_batch_uses_every_launch_of_the_fused_path(oracle):
    """More frames than co-resident clusters: the cluster kernel and (hybrid mode) the software clusters beside it draw frames
    from one queue; every frame must equal the oracle's, whatever cluster processed it."""
    import torch

    p = P.preset(P.GS_PRESET_HDL64E_2048)
    H, W = p.shape
    pool, offs = synth.make_pool(p, 3, seed0=synth.SEED0 + 90)
    want = [oracle.process_frame(p, pool[offs[i]:offs[i + 1]]) for i in range(3)]
    F = 150
    pts, offsets = synth.tile_pool(pool, offs, F)
    d_pts = torch.from_numpy(pts).cuda()
    d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_xyz = torch.zeros((F, H * W, 4), dtype=torch.float32, device="cuda")
    d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with capi.Context(p) as ctx:
        for rep in range(2):  # the second batch reuses the queue and the mailboxes
            d_res.zero_()
            d_cnt.zero_()
            torch.cuda.synchronize()
            l0 = ctx.stats().kernel_launches
            ctx.process_batch_device(d_pts.data_ptr(), offsets, F, 0, 0, d_res.data_ptr(), d_xyz.data_ptr(), d_cnt.data_ptr())
            ctx.synchronize()
            assert 1 <= ctx.stats().kernel_launches - l0 <= 2
            res = d_res.cpu().numpy().view(np.uint16)
            cnt = d_cnt.cpu().numpy()
            for f in range(F):
                w = want[f % 3]
                assert np.array_equal(res[f], w["no_ground"]), (rep, f)
                assert cnt[f] == w["xyz"].shape[0]
            xyz = d_xyz[F - 1, : cnt[F - 1]].cpu().numpy()
            assert np.array_equal(xyz, want[(F - 1) % 3]["xyz"])


def test_backproject_with_foreign_camera_info(ctxs, oracle):
    """Range2CloudNode::CameraCb decodes with K[0], K[2..5] as published (range2cloud_node.cpp:51-58): a CameraInfo whose
    d_azimuth / d_altitude are not what this library would derive must be honoured bit for bit."""
    p = P.preset(P.GS_PRESET_VLP16).copy(variant=P.GS_VARIANT_CLOUD2RANGE)
    ctx = ctxs(p)
    pts = synth.make_frame(p, synth.SEED0 + 62)
    img = oracle.process_frame(p, pts)["no_ground"]
    d = oracle.derive(p)
    K = [p.min_angle, p.max_angle, p.min_range * 1.5, p.max_range * 0.9, d.d_azimuth * (1 + 3e-16), d.d_altitude * (1 - 2e-16)]
    rows, cols = np.nonzero(img)
    rng = (img[rows, cols].astype(np.float64) - 1) / 65534 * (K[3] - K[2]) + K[2]
    alt = (img.shape[0] - rows) * K[5] + K[0]
    az = cols * K[4]
    got = ctx.backproject_camera_info(img, K, True)
    want = np.stack([np.cos(alt) * np.cos(az) * rng, np.cos(alt) * np.sin(az) * rng, np.sin(alt) * rng], 1).astype(np.float32)
    assert np.array_equal(got[:, :3], want) and np.all(got[:, 3] == 1.0)
    ref = oracle.ref()
    if ref is not None:  # the reference's own node on the same message
        assert np.array_equal(got, oracle.RefRange2Cloud().camera_cb(img, K))
    rngf = rng.astype(np.float32).astype(np.float64)  # Alfa_GS::CameraCb rounds the range to float (alfa_gs.cpp:1027,1045)
    got_f = ctx.backproject_camera_info(img, K, False)
    want_f = np.stack([np.cos(alt) * np.cos(az) * rngf, np.cos(alt) * np.sin(az) * rngf, np.sin(alt) * rngf], 1).astype(np.float32)
    assert np.array_equal(got_f[:, :3], want_f)
    with pytest.raises(capi.GsError):
        ctx.backproject_camera_info(img[:-1], K, True)  # not the context's shape


def test_host_batch_record_strides_and_error_exits(ctxs, oracle):
    """gs_process_batch_host_strided: 12-byte XYZ and 32-byte PointXYZRGB records give the XYZI result; a bad offset table is
    refused before anything is copied and a too small cloud buffer is reported — the context stays usable after both."""
    p = P.preset(P.GS_PRESET_VLP16)
    ctx = ctxs(p, "production")
    H, W = p.shape
    pool, offs = synth.make_pool(p, 4, seed0=synth.SEED0 + 95)
    pts, offsets = synth.tile_pool(pool, offs, 9)
    F = len(offsets) - 1
    want = [oracle.process_frame(p, pool[offs[i]:offs[i + 1]]) for i in range(4)]

    def run(points, cap=None, offsets_=offsets):
        h_res = np.zeros((F, H, W), np.uint16)
        h_xyz = np.zeros((cap if cap is not None else F * H * W, 4), np.float32)
        _, _, out_offs = ctx.process_batch_host(points, offsets_, h_res, h_xyz)
        return h_res, h_xyz, out_offs

    ctx.set_chunk_frames(4)
    for points in (pts, np.ascontiguousarray(pts[:, :3]), np.concatenate([pts, np.zeros_like(pts)], 1)):
        assert points.strides[0] in (16, 12, 32)
        h_res, h_xyz, out_offs = run(points)
        for f in range(F):
            w = want[f % 4]
            assert np.array_equal(h_res[f], w["no_ground"])
            assert np.array_equal(h_xyz[out_offs[f]:out_offs[f + 1]], w["xyz"])
    bad = offsets.copy()
    bad[5] = bad[4] - 7
    with pytest.raises(capi.GsError) as ei:
        run(pts, offsets_=bad)
    assert ei.value.code == P.GS_EINVAL
    with pytest.raises(capi.GsError) as ei:
        run(pts, cap=int(want[0]["xyz"].shape[0]) + 10)  # room for the first frame only
    assert ei.value.code == P.GS_EINVAL
    h_res, h_xyz, out_offs = run(pts)  # and the context still works
    assert np.array_equal(h_res[F - 1], want[(F - 1) % 4]["no_ground"])
    ctx.set_chunk_frames(0)


@pytest.mark.parametrize("beta_deg", [5, 10, 15, 20])
@pytest.mark.parametrize("name", ["C3_os1_128", "cloud2range_sg7", "C1_vlp16"])
def test_batch_labels_match_oracle(oracle, name, beta_deg):
    """gs_process_batch_device_labels: the Labeler as an all-components CCL on every frame of a batch (BASELINE configs C3 / C5),
    labels in canonical min-index form against the oracle's serial raster-scan driver of the BFS."""
    import torch

    if name != "C3_os1_128" and beta_deg not in (5, 20):
        pytest.skip("the full sweep runs on the C3 / C5 shape")
    p = P.preset(CONFIGS[name])
    H, W = p.shape
    beta = math.radians(beta_deg)
    pool, offs = synth.make_pool(p, 3, seed0=synth.SEED0 + 300 + beta_deg)
    F = 7 if name == "C3_os1_128" else 45
    pts, offsets = synth.tile_pool(pool, offs, F)
    want = []
    for i in range(3):
        w = oracle.process_frame(p, pool[offs[i]:offs[i + 1]])
        lab, n = oracle.label_components(p, w["no_ground"], w["angle"], beta)
        want.append((w, lab, n))
    d_pts = torch.from_numpy(pts).cuda()
    d_res = torch.zeros((F, H, W), dtype=torch.int16, device="cuda")
    d_lab = torch.full((F, H, W), -7, dtype=torch.int32, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with capi.Context(p) as ctx:
        # with and without caller-provided image buffers (wave scratch)
        for give_res in (True, False):
            d_lab.fill_(-7)
            torch.cuda.synchronize()
            ctx.process_batch_device_labels(d_pts.data_ptr(), offsets, F, d_res.data_ptr() if give_res else 0, 0, d_cnt.data_ptr(), d_lab.data_ptr(), beta,
                                            d_n_components=d_n.data_ptr())
            ctx.synchronize()
            lab = d_lab.cpu().numpy()
            nn = d_n.cpu().numpy()
            for f in range(F):
                w, wl, wn = want[f % 3]
                assert np.array_equal(lab[f], wl), (give_res, f)
                assert nn[f] == wn
                if give_res:
                    assert np.array_equal(d_res[f].cpu().numpy().view(np.uint16), w["no_ground"])


def test_native_multi_gpu_driver_shards_equal_one_gpu(oracle):
    """gs_multi (one host thread + context per GPU, contiguous frame blocks, no collective): N shards == 1 shard, bit for bit.
    With one GPU visible the two shards share it (two contexts); with more, every GPU takes a shard as well."""
    import torch

    p = P.preset(P.GS_PRESET_VLP16)
    H, W = p.shape
    pool, offs = synth.make_pool(p, 5, seed0=synth.SEED0 + 1500)
    F = 23
    pts, offsets = synth.tile_pool(pool, offs, F)
    want = [oracle.process_frame(p, pool[offs[i]:offs[i + 1]]) for i in range(5)]
    n_dev = torch.cuda.device_count()
    layouts = [[0, 0], [0, 0, 0]] + ([list(range(n_dev))] if n_dev > 1 else [])
    for devices in layouts:
        with capi.MultiGpu(p, devices) as mg:
            assert mg.n_gpus == len(devices)
            assert [w["device"] for w in mg.workers()] == devices
            h_pts = mg.host_array(pts.shape, np.float32)
            h_pts[:] = pts
            h_res = mg.host_array((F, H, W), np.uint16)
            h_xyz = mg.host_array((F * H * W, 4), np.float32)
            for points in (h_pts, np.ascontiguousarray(pts[:, :3])):  # pinned XYZI, pageable 12-byte XYZ
                h_res[:] = 0
                _, _, first, count = mg.process_batch_host(points, offsets, h_res, h_xyz)
                for f in range(F):
                    w = want[f % 5]
                    assert np.array_equal(h_res[f], w["no_ground"]), (devices, f)
                    assert count[f] == w["xyz"].shape[0]
                    assert np.array_equal(h_xyz[first[f]:first[f] + count[f]], w["xyz"])
            bad = offsets.copy()
            bad[3] = bad[2] - 1
            with pytest.raises(capi.GsError):
                mg.process_batch_host(h_pts, bad, h_res, h_xyz)
            _, _, first, count = mg.process_batch_host(h_pts, offsets, h_res, h_xyz)  # still usable
            assert count[F - 1] == want[(F - 1) % 5]["xyz"].shape[0]


def test_fpga_wire_formats_and_colored_angle_image(ctxs, oracle):
    """N4: AlfaNode's FPGA codecs (src/alfa_node.cpp:105-271) and Alfa_GS::CreateColoredAngleImage (src/alfa_gs.cpp:768-829) against
    the reference's OWN members (oracle/_ref, compiled verbatim) and against numpy restatements of the integer formats."""
    p = P.preset(P.GS_PRESET_ALFA_GS_DEFAULT)
    ctx = ctxs(p)
    rng = np.random.default_rng(41)
    pts = synth.make_frame(p, synth.SEED0 + 77)
    extra = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 1, 0], [0, 0, -1, 0], [3, -4, 12, 0], [-2.5, -0.0, 0.1, 0],
                      [1e-3, 1e-3, 1e-3, 0], [100, -250, 30, 0]], np.float32)
    pts = np.concatenate([pts, extra])
    words = ctx.fpga_store_pointcloud(pts)
    # numpy restatement of the packing (the float arithmetic is the reference's own, below)
    assert np.all((words >> np.uint64(48)) == 0)
    rng_cm = ((words >> np.uint64(32)) & np.uint64(0xffff)).astype(np.int64)
    r = np.sqrt((pts[:, 0] * pts[:, 0] + pts[:, 1] * pts[:, 1] + pts[:, 2] * pts[:, 2]).astype(np.float32))
    assert np.array_equal(rng_cm, (r * np.float32(100)).astype(np.int64) & 0xffff)
    # read_hardware_pointcloud: 4 pixels per word, column-major, 64 rows per column, hard-coded 64 x 1800 image
    for rows, cols in ((64, 1800), (64, 900), (16, 1800)):
        px = rng.integers(0, 65536, size=rows * cols, dtype=np.uint64)
        w = (px[0::4] | (px[1::4] << np.uint64(16)) | (px[2::4] << np.uint64(32)) | (px[3::4] << np.uint64(48))).astype(np.uint64)
        got = ctx.fpga_read_range_image(w, rows, cols)
        want = np.zeros((64, 1800), np.uint16)
        k = np.arange(rows * cols)
        want[k % 64, k // 64] = px.astype(np.uint16)
        assert np.array_equal(got, want)
    with pytest.raises(capi.GsError):
        ctx.fpga_read_range_image(np.zeros(64 * 2048 // 4, np.uint64), 64, 2048)
    # read_hardware_filtered_angle_image: 10-bit fields, 6 per word, every 43rd word 4
    for cols in (1800, 50, 4):
        n = 64 * cols
        vals = rng.integers(0, 1024, size=n, dtype=np.uint64)
        n_words = (n // 256) * 43 + (43 if n % 256 >= 252 else -(-(n % 256) // 6))
        w = np.zeros(n_words + 1, np.uint64)
        for i in range(n):
            b, wi = divmod(i, 256)
            word, field = (b * 43 + wi // 6, wi % 6) if wi < 252 else (b * 43 + 42, wi - 252)
            w[word] |= vals[i] << np.uint64(10 * field)
        got = ctx.fpga_read_angle_image(w, 64, cols)
        want = np.zeros((64, cols), np.uint16)
        k = np.arange(n)
        want[k % 64, k // 64] = vals.astype(np.uint16)
        assert np.array_equal(got, want), cols
    with pytest.raises(capi.GsError):
        ctx.fpga_read_angle_image(np.zeros(100, np.uint64), 32, 10)
    # coloured angle image: running maximum in raster order
    ang = np.clip(rng.integers(-30, 158, size=(64, 300)), 0, 157).astype(np.uint16)
    ang[0, :5] = 0
    got = ctx.colored_angle_image(ang)
    flat = ang.reshape(-1).astype(np.float64)
    mx = np.maximum.accumulate(flat)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.where(mx > 0, np.floor(255.0 / np.where(mx > 0, mx, 1) * flat), 0).astype(np.int64)
    want = np.zeros((flat.size, 3), np.uint8)
    hi, lo = s >= 128, (s > 0) & (s < 128)
    want[hi, 0], want[hi, 1] = ((s[hi] - 128) * 2).astype(np.uint8), (255 - (s[hi] - 128) * 2).astype(np.uint8)
    want[lo, 1], want[lo, 2] = (s[lo] * 2).astype(np.uint8), (255 - 2 * s[lo]).astype(np.uint8)
    assert np.array_equal(got.reshape(-1, 3), want)
    if oracle.ref() is not None:  # the reference's own members
        node = oracle.RefAlfaGS(p)
        assert np.array_equal(words, node.store_pointcloud_hardware(pts))
        px = rng.integers(0, 65536, size=64 * 1800, dtype=np.uint64)
        w = (px[0::4] | (px[1::4] << np.uint64(16)) | (px[2::4] << np.uint64(32)) | (px[3::4] << np.uint64(48))).astype(np.uint64)
        assert np.array_equal(ctx.fpga_read_range_image(w, 64, 1800), node.read_hardware_pointcloud(w, 64, 1800))
        wa = rng.integers(0, 1 << 62, size=(64 * 1800 // 256 + 1) * 43, dtype=np.uint64)
        assert np.array_equal(ctx.fpga_read_angle_image(wa, 64, 1800), node.read_hardware_filtered_angle_image(wa, 64, 1800))
        assert np.array_equal(got, node.colored_angle_image(ang))


@pytest.mark.parametrize("name", ["C1_vlp16", "C2_hdl64_2048", "C3_os1_128", "cloud2range_sg7"])
def test_lone_frame_into_pinned_buffers(oracle, name):
    """The latency path of gs_process_frame: with page-locked output buffers the fused kernel stores image, cloud and count straight
    into host memory (no device-to-host copies, one synchronisation).  Same bits as the copying path and the oracle."""
    import torch

    p = P.preset(CONFIGS[name])
    H, W = p.shape
    frames = [synth.make_frame(p, synth.SEED0 + 1700 + k) for k in range(3)]
    frames.append(np.zeros((0, 4), np.float32))
    h_res = torch.zeros((H, W), dtype=torch.int16).pin_memory()
    h_xyz = torch.zeros((H * W, 4), dtype=torch.float32).pin_memory()
    with capi.Context(p) as ctx:
        for pts in frames:
            want = oracle.process_frame(p, pts)
            h_in = torch.from_numpy(pts.copy()).pin_memory() if pts.shape[0] else torch.zeros((0, 4), dtype=torch.float32)
            h_res.fill_(-1)
            l0 = ctx.stats().kernel_launches
            n = ctx.process_frame_into(h_in.numpy(), h_res.numpy().view(np.uint16), h_xyz.numpy())
            assert ctx.stats().kernel_launches - l0 == 1
            assert n == want["xyz"].shape[0]
            assert np.array_equal(h_res.numpy().view(np.uint16), want["no_ground"])
            assert np.array_equal(h_xyz.numpy()[:n], want["xyz"])
            # image only, and pageable buffers (the copying path)
            h_res.fill_(-1)
            assert ctx.process_frame_into(h_in.numpy(), h_res.numpy().view(np.uint16), None) == n
            assert np.array_equal(h_res.numpy().view(np.uint16), want["no_ground"])
            res2 = np.zeros((H, W), np.uint16)
            xyz2 = np.zeros((H * W, 4), np.float32)
            assert ctx.process_frame_into(pts, res2, xyz2) == n
            assert np.array_equal(res2, want["no_ground"]) and np.array_equal(xyz2[:n], want["xyz"])


@pytest.mark.parametrize("name", ["C1_vlp16", "C2_hdl64_2048", "C3_os1_128", "cloud2range_sg7"])
def test_nearest_return_policy_in_the_fused_kernel(ctxs, oracle, name):
    """pixel_policy = nearest (the north-star variant: smallest range wins a pixel, ties to the lowest index) runs in the fused kernel
    like the reference's last-writer rule: one launch, every instantiation, several returns per pixel in random order, ranges
    on the gates, and a batch."""
    import torch

    p = P.preset(CONFIGS[name]).copy(pixel_policy=P.GS_PIXEL_NEAREST)
    H, W = p.shape
    frames = [synth.make_adversarial_frame(p, synth.SEED0 + 21, "collisions"), synth.make_adversarial_frame(p, synth.SEED0 + 22, "edges"),
              synth.make_frame(p, synth.SEED0 + 23, shuffle=True)]
    want = [oracle.process_frame(p, f) for f in frames]
    assert not np.array_equal(want[0]["range"], oracle.process_frame(p.copy(pixel_policy=P.GS_PIXEL_LAST), frames[0])["range"])
    for mode in MODES:
        ctx = ctxs(p, mode)
        for f, w in zip(frames, want):
            l0 = ctx.stats().kernel_launches
            got = ctx.process_frame(f)
            assert (ctx.stats().kernel_launches - l0 == 1) == (mode is not True)
            _, ne_p, _ = oracle.create_range_image(p, f, with_counts=True)
            bad = _assert_images_equal(got["range"], w["range"], "range image", ne_p)
            if bad == 0:
                assert np.array_equal(got["angle"], w["angle"])
                assert np.array_equal(got["no_ground"], w["no_ground"])
                assert np.array_equal(got["xyz"], w["xyz"])
    offs = np.concatenate([[0], np.cumsum([f.shape[0] for f in frames])]).astype(np.int64)
    pts, offsets = synth.tile_pool(np.concatenate(frames), offs, 40)
    d_pts = torch.from_numpy(pts).cuda()
    d_rng = torch.zeros((40, H, W), dtype=torch.int16, device="cuda")
    d_cnt = torch.zeros(40, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx = ctxs(p, "production")
    ctx.process_batch_device(d_pts.data_ptr(), offsets, 40, d_rng.data_ptr(), 0, 0, 0, d_cnt.data_ptr())
    ctx.synchronize()
    rng_img = d_rng.cpu().numpy().view(np.uint16)
    for f in (0, 2, 3, 39):  # the collision frame and the shuffled one (no cell-edge points), early and late in the queue
        assert np.array_equal(rng_img[f], want[f % 3]["range"])
        assert int(d_cnt[f]) == want[f % 3]["xyz"].shape[0]
