"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors
produced by the real reference.  Bit-exact (`np.array_equal`) everywhere except the point cloud,
whose reference goes through a BLAS sgemm (tolerance stated in the test)."""
import hashlib

import numpy as np
import pytest

from oracle import xmaps_oracle as orc
from xm_helpers import golden_frame, load_golden_tables

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def make_engine(tables, z=None):
    from xmaps_b200.engine import DepthEngine, TableSet

    kw = {}
    if z is not None and "lut_x_f32" in z:
        kw = dict(lut_x_f32=z["lut_x_f32"], lut_y_f32=z["lut_y_f32"])
    return DepthEngine(
        TableSet(
            lut_x=tables.lut_x,
            lut_y=tables.lut_y,
            x_map=tables.x_map,
            remap_xy=tables.remap_xy,
            rect_w=tables.rect_w,
            rect_h=tables.rect_h,
            t_px_scale=tables.t_px_scale,
            x_offset=tables.x_offset,
            depth_scale=tables.depth_scale,
            **kw,
        ),
        device="cuda:0",
    )


@pytest.fixture(scope="module")
def small():
    tables, z = load_golden_tables("small")
    eng = make_engine(tables, z)
    yield tables, z, eng
    eng.close()


@pytest.fixture(scope="module")
def default():
    tables, z = load_golden_tables("default")
    eng = make_engine(tables)
    yield tables, z, eng
    eng.close()


def E():
    import xmaps_b200.engine as e

    return e


VIEWS = [(0, "proj"), (1, "cam")]


# ------------------------------------------------------------------------------------------------
# golden vectors of the real reference
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("view,tag", VIEWS)
@pytest.mark.parametrize("bounds", ["reduce", "sorted", "given"])
@pytest.mark.parametrize("stage_xmap", [1, 0])
def test_small_frame_matches_reference(small, view, tag, bounds, stage_xmap):
    tables, _, eng = small
    e = E()
    eng.set_option("stage_xmap", stage_xmap)
    try:
        ev = orc.synth_events(2, 20_000, 160, 120)
        g = golden_frame(f"small_20k_{tag}")
        kw = {"time_bounds": {"reduce": e.TBOUNDS_REDUCE, "sorted": e.TBOUNDS_SORTED, "given": e.TBOUNDS_GIVEN}[bounds]}
        if bounds == "given":
            pos = ev[ev["p"] == 1]
            kw.update(t_min=int(pos["t"].min()), t_max=int(pos["t"].max()))
        dev = eng.events(ev)
        depth = eng.frame(dev, view=view, output=e.OUT_DEPTH, **kw).cpu().numpy()
        assert np.array_equal(depth, g["depth"])
        st = eng.status()
        assert st["n_valid"] == int(g["n_pos"][0])
        assert st["n_inliers"] == len(g["disp"])
        assert st["flags"] == 0 and not st["fixup_ran"]
        disp = eng.frame(dev, view=view, output=e.OUT_DISPARITY, **kw).cpu().numpy()
        assert np.array_equal(disp, g["disp_map"])
        bgr = eng.frame(dev, view=view, output=e.OUT_BGR, z_near=0.1, z_far=1.0, **kw).cpu().numpy()
        assert np.array_equal(bgr, g["bgr"])
    finally:
        eng.set_option("stage_xmap", 1)


@pytest.mark.parametrize("view,tag", VIEWS)
def test_default_100k_matches_reference(default, manifest, view, tag):
    tables, _, eng = default
    e = E()
    ev = orc.synth_events(0, 100_000, 640, 480)
    g = golden_frame(f"default_100k_{tag}")
    depth = eng.frame(ev, view=view).cpu().numpy()
    assert np.array_equal(depth, g["depth"])
    assert sha(depth) == manifest["configs"]["default"]["hash"][f"depth_{tag}"]
    bgr = eng.frame(ev, view=view, output=e.OUT_BGR).cpu().numpy()
    assert np.array_equal(bgr, g["bgr"])
    st = eng.status()
    assert st["n_valid"] == manifest["configs"]["default"]["n_pos"]
    assert st["n_inliers"] == manifest["configs"]["default"]["n_inliers"]


def test_checked_scatter_variant(default):
    """The tables of a valid calibration select the check-free scatter; the checked variant must
    give the same frames, and corrupt X-map cells must be reported, not written out of bounds."""
    tables, _, eng = default
    ev = orc.synth_events(0, 100_000, 640, 480)
    assert eng.get_option("safe_tables") == 1
    want = golden_frame("default_100k_proj")["depth"]
    eng.set_option("safe_tables", 0)
    try:
        assert eng.get_option("safe_tables") == 0
        assert np.array_equal(eng.frame(ev, view=0).cpu().numpy(), want)
    finally:
        eng.set_option("safe_tables", 1)
    bad = tables.x_map.copy()
    bad[200:900:7, 100:600:5] = tables.x_offset + tables.rect_w + 50  # outside the rectified image
    eng2 = make_engine(type(tables)(**{**tables.__dict__, "x_map": bad}))
    try:
        assert eng2.get_option("safe_tables") == 0
        eng2.frame(ev, view=0)
        assert eng2.status()["scatter_oob"]
    finally:
        eng2.close()


def test_default_1m_hashes(default, manifest):
    _, _, eng = default
    h = manifest["configs"]["default"]["hash"]
    ev = orc.synth_events(1, 1_000_000, 640, 480)
    assert sha(eng.frame(ev, view=0).cpu().numpy()) == h["depth_proj_seed1_1m"]
    assert sha(eng.frame(ev, view=1).cpu().numpy()) == h["depth_cam_seed1_1m"]


# ------------------------------------------------------------------------------------------------
# unsorted input: optimistic bounds must be detected and fixed on the device
# ------------------------------------------------------------------------------------------------
def test_unsorted_timestamps_fixup(small):
    tables, _, eng = small
    e = E()
    ev = orc.synth_events(2, 20_000, 160, 120)
    np.random.default_rng(7).shuffle(ev)
    g = golden_frame("small_20k_shuffled_proj")
    depth = eng.frame(ev, view=0, time_bounds=e.TBOUNDS_SORTED).cpu().numpy()
    st = eng.status()
    assert st["fixup_ran"] and st["tbounds_violated"]
    assert np.array_equal(depth, g["depth"])
    assert st["n_inliers"] == len(g["disp"])
    depth = eng.frame(ev, view=0, time_bounds=e.TBOUNDS_REDUCE).cpu().numpy()
    assert np.array_equal(depth, g["depth"])
    assert not eng.status()["fixup_ran"]
    # a sorted frame right after a fixed-up one is unaffected
    ev2 = orc.synth_events(2, 20_000, 160, 120)
    assert np.array_equal(eng.frame(ev2, view=0).cpu().numpy(), golden_frame("small_20k_proj")["depth"])
    # without the fix-up the violation is only reported
    eng.set_option("auto_fixup", 0)
    try:
        eng.frame(ev, view=0, time_bounds=e.TBOUNDS_SORTED)
        st = eng.status()
        assert st["tbounds_violated"] and not st["fixup_ran"]
    finally:
        eng.set_option("auto_fixup", 1)


def test_wrong_given_bounds_are_fixed(small):
    tables, _, eng = small
    e = E()
    ev = orc.synth_events(2, 20_000, 160, 120)
    g = golden_frame("small_20k_cam")
    depth = eng.frame(ev, view=1, time_bounds=e.TBOUNDS_GIVEN, t_min=100, t_max=5000).cpu().numpy()
    assert eng.status()["fixup_ran"]
    assert np.array_equal(depth, g["depth"])


# ------------------------------------------------------------------------------------------------
# edge cases
# ------------------------------------------------------------------------------------------------
def test_degenerate_frames(small):
    tables, _, eng = small
    e = E()
    empty = np.zeros(0, dtype=orc.EVENT_DTYPE)
    for view in (0, 1):
        assert not eng.frame(empty, view=view).cpu().numpy().any()
    ev = orc.synth_events(5, 64, 160, 120)
    neg = ev.copy()
    neg["p"] = 0
    assert not eng.frame(neg, view=1).cpu().numpy().any()
    assert eng.status()["n_valid"] == 0
    same = ev.copy()
    same["t"] = 1234  # 0/0 -> NaN -> column 0, as NumPy does
    for view in (0, 1):
        for tb in (e.TBOUNDS_SORTED, e.TBOUNDS_REDUCE):
            got = eng.frame(same, view=view, time_bounds=tb).cpu().numpy()
            assert np.array_equal(got, orc.frame_depth(tables, same, view))
    one = ev[:1].copy()
    one["p"] = 1
    assert np.array_equal(eng.frame(one, view=1).cpu().numpy(), orc.frame_depth(tables, one, 1))


def test_polarity_flag_equals_prefiltered(small):
    tables, _, eng = small
    ev = orc.synth_events(11, 30_000, 160, 120, p_on=0.5)
    pos = ev[ev["p"] == 1]
    for view in (0, 1):
        a = eng.frame(ev, view=view, polarity=True).cpu().numpy()
        b = eng.frame(pos, view=view, polarity=False).cpu().numpy()
        assert np.array_equal(a, b)
        assert np.array_equal(a, orc.frame_depth(tables, ev, view))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1023, 1024, 1025, 4097, 70_001])
def test_ragged_sizes(small, n):
    tables, _, eng = small
    ev = orc.synth_events(100 + n, n, 160, 120)
    for view in (0, 1):
        assert np.array_equal(eng.frame(ev, view=view).cpu().numpy(), orc.frame_depth(tables, ev, view))


def test_heavy_collisions_last_write_wins(small):
    """Every event on a handful of pixels: the scatter must keep the LAST event per cell."""
    tables, _, eng = small
    rng = np.random.default_rng(3)
    n = 200_000
    ev = orc.synth_events(3, n, 160, 120, p_on=1.0)
    ev["x"] = rng.integers(60, 64, n)
    ev["y"] = rng.integers(50, 54, n)
    for view in (0, 1):
        assert np.array_equal(eng.frame(ev, view=view).cpu().numpy(), orc.frame_depth(tables, ev, view))


def test_float64_timestamps(small):
    """Evaluation path: dict of columns with a float t (compute_depth_x_maps.py:83-96)."""
    tables, _, eng = small
    e = E()
    rng = np.random.default_rng(5)
    n = 40_000
    evd = {"x": rng.integers(0, 160, n), "y": rng.integers(0, 120, n), "t": rng.random(n)}
    xcr, ycr = orc.rectify_i16(tables, evd)
    disp, mask = orc.event_disparity(tables, xcr, ycr, evd["t"])
    want = orc.disparity_to_depth(orc.scatter_camera_view(tables, evd, mask, disp), tables.depth_scale)
    got = eng.frame(evd, view=1, polarity=False, time_bounds=e.TBOUNDS_REDUCE).cpu().numpy()
    assert np.array_equal(got, want)
    got = eng.frame(evd, view=1, polarity=False, time_bounds=e.TBOUNDS_SORTED).cpu().numpy()  # unsorted -> fix-up
    assert np.array_equal(got, want)
    assert eng.status()["fixup_ran"]


def test_repeat_and_epoch_wrap(small):
    tables, _, eng = small
    ev_a = orc.synth_events(21, 5_000, 160, 120)
    ev_b = orc.synth_events(22, 5_000, 160, 120)
    want_a = orc.frame_depth(tables, ev_a, 0)
    want_b = orc.frame_depth(tables, ev_b, 0)
    eng.set_option("epoch", 0xFFFF - 5)
    for _ in range(6):  # crosses the 16-bit wrap of the scatter-map epoch
        assert np.array_equal(eng.frame(ev_a, view=0).cpu().numpy(), want_a)
        assert np.array_equal(eng.frame(ev_b, view=0).cpu().numpy(), want_b)
    assert eng.get_option("epoch") < 100


def test_batch_equals_single(small):
    tables, _, eng = small
    frames = [orc.synth_events(40 + i, 8_000 + 100 * i, 160, 120) for i in range(5)]
    for view in (0, 1):
        out = eng.frame_batch(frames, view=view).cpu().numpy()
        for i, f in enumerate(frames):
            assert np.array_equal(out[i], orc.frame_depth(tables, f, view))


def test_host_buffer_call(small):
    tables, _, eng = small
    e = E()
    ev = orc.synth_events(2, 20_000, 160, 120)
    g = golden_frame("small_20k_proj")
    assert np.array_equal(eng.frame_host(ev, view=0), g["depth"])
    assert np.array_equal(eng.frame_host(ev, view=0, output=e.OUT_BGR), g["bgr"])
    pinned = torch.from_numpy(ev.view(np.int32).reshape(-1, 4)).pin_memory()
    assert np.array_equal(eng.frame_host(pinned, view=1), golden_frame("small_20k_cam")["depth"])


def test_invalid_pixels_are_flagged(small):
    tables, _, eng = small
    ev = orc.synth_events(2, 2_000, 160, 120)
    ev["x"][7] = 160  # outside the camera image: the reference raises IndexError
    eng.frame(ev, view=1)
    assert eng.status()["pixel_oob"]


def test_error_reporting(small):
    from xmaps_b200._native import XmapsError

    tables, _, eng = small
    with pytest.raises(XmapsError):
        eng.frame(orc.synth_events(2, 100, 160, 120), view=7)
    with pytest.raises(XmapsError):
        eng.set_option("no_such_option", 1)


# ------------------------------------------------------------------------------------------------
# stage-by-stage entry points
# ------------------------------------------------------------------------------------------------
def test_staged_chain_matches_reference(small):
    tables, z, eng = small
    e = E()
    g = golden_frame("small_20k_proj")
    ev = orc.polarity_mask(orc.synth_events(2, 20_000, 160, 120))
    dev = eng.events(ev)
    xr, yr = eng.rectify_i16(dev)
    assert np.array_equal(xr.cpu().numpy(), g["xr"]) and np.array_equal(yr.cpu().numpy(), g["yr"])
    xf, yf = eng.rectify_f32(dev)
    assert np.array_equal(xf.cpu().numpy(), z["lut_x_f32"][ev["y"], ev["x"]])
    assert np.array_equal(yf.cpu().numpy(), z["lut_y_f32"][ev["y"], ev["x"]])
    for given in (True, False):
        full, mask = eng.event_disparity(dev, xr if given else None, yr if given else None)
        assert np.array_equal(np.packbits(mask.cpu().numpy().astype(bool)), g["mask_bits"])
        disp = eng.compact_i16(full, mask)
        assert np.array_equal(disp.cpu().numpy(), g["disp"])
    m = mask.bool()
    xpr = (xr[m] + disp).to(torch.int16)
    rect = eng.scatter_last_wins(yr[m].contiguous(), xpr.contiguous(), disp, tables.rect_h, tables.rect_w)
    assert np.array_equal(rect.cpu().numpy(), g["rect_map"])
    proj = eng.dilate_remap(rect)
    assert np.array_equal(proj.cpu().numpy(), g["disp_map"])
    assert np.array_equal(eng.disp_to_depth(proj).cpu().numpy(), g["depth"])
    assert np.array_equal(eng.colorize(proj, 0.1, 1.0).cpu().numpy(), g["bgr"])


def test_compaction_sizes(small):
    _, _, eng = small
    rng = np.random.default_rng(9)
    for n in (0, 1, 1023, 1024, 1025, 5000, 1_300_000):
        vals = rng.integers(-3000, 3000, n).astype(np.int16)
        mask = (rng.random(n) < 0.3).astype(np.uint8)
        got = eng.compact_i16(torch.from_numpy(vals).cuda(), torch.from_numpy(mask).cuda()).cpu().numpy()
        assert np.array_equal(got, vals[mask.astype(bool)])


def test_dilate_remap_generic_values(small):
    """The materialised-map entry point must follow cv2.dilate for any float values."""
    tables, _, eng = small
    rng = np.random.default_rng(4)
    m = np.where(rng.random((tables.rect_h, tables.rect_w)) < 0.05, rng.normal(0, 50, (tables.rect_h, tables.rect_w)), 0).astype(np.float32)
    want = orc.dilate_remap(tables, m, use_cv2=True)
    got = eng.dilate_remap(torch.from_numpy(m).cuda()).cpu().numpy()
    assert np.array_equal(got, want)


def test_point_cloud(small):
    tables, z, eng = small
    rng = np.random.default_rng(8)
    n = 10_000
    x = rng.uniform(0, tables.rect_w, n).astype(np.float32)
    y = rng.uniform(0, tables.rect_h, n).astype(np.float32)
    d = rng.integers(1, 300, n).astype(np.float32)
    Q = z["Q"]
    pts = np.ones((n, 4), np.float32)
    pts[:, 0], pts[:, 1], pts[:, 2] = x + d, y, -d
    pc = (Q.astype(np.float32) @ pts.T).T  # cam_proj_calibration.py:319-331
    pc = (pc / pc[:, 3:])[:, :3]
    pc[:, 1:] = -pc[:, 1:]
    got = eng.point_cloud(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), torch.from_numpy(d).cuda(), Q).cpu().numpy()
    # float32 dot products in a different summation order than BLAS: tolerance 1e-5 relative
    np.testing.assert_allclose(got, pc, rtol=1e-5, atol=1e-6)


def test_x_map_builder(small):
    from xmaps_b200.engine import build_x_map

    tables, z, _ = small
    w = int(z["consts"][2])
    x_map, t_diffs = build_x_map(z["time_map_rect"], w, int(z["consts"][0]), int(z["consts"][1]), num_scanlines=w)
    assert np.array_equal(x_map.cpu().numpy(), tables.x_map)
    _, want_diffs = orc.build_x_map(z["time_map_rect"], w, int(z["consts"][0]), int(z["consts"][1]), num_scanlines=w)
    assert np.array_equal(t_diffs.cpu().numpy(), want_diffs)


def test_bilinear_lookup_matches_its_oracle(small):
    """XM_FLAG_BILINEAR (opt-in, not reference behaviour): bilinear X-map lookup at the un-rounded rectified row / time
    column, undefined cells left out of the blend.  The kernel evaluates the oracle's float64 expression operation by
    operation, so the float32 maps are compared exactly; the stated tolerance of the feature is rtol = 1e-6."""
    tables, z, eng = small
    e = E()
    lx, ly = z["lut_x_f32"], z["lut_y_f32"]
    ev = orc.synth_events(41, 400_000, tables.cam_w, tables.cam_h)
    for view in (0, 1):
        want = orc.frame_disparity_map_bilinear(tables, lx, ly, ev, view)
        got = eng.frame(ev, view=view, output=e.OUT_DISPARITY, time_bounds=e.TBOUNDS_REDUCE, bilinear=True).cpu().numpy()
        assert (want > 0).sum() > 1000
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=0)
        assert np.array_equal(got, want), f"view {view}: {np.count_nonzero(got != want)} pixels differ in the last bit"
        depth = eng.frame(ev, view=view, output=e.OUT_DEPTH, bilinear=True).cpu().numpy()
        assert np.array_equal(depth, orc.disparity_to_depth(want, tables.depth_scale))
        st = eng.status()
        assert st["n_valid"] == int((ev["p"] == 1).sum()) and not st["tbounds_violated"]
    # it is a different estimator from the nearest lookup (sub-pixel disparities), yet close to it
    near = orc.frame_disparity_map(tables, ev, 0)
    bil = orc.frame_disparity_map_bilinear(tables, lx, ly, ev, 0)
    both = (near > 0) & (bil > 0)
    assert both.sum() > 1000 and np.any(bil[both] != np.rint(bil[both])) and np.median(np.abs(near[both] - bil[both])) < 2.0
    # float64 timestamps, no polarity mask, empty frame
    from xmaps_b200.events import pack_events

    evf = pack_events(ev["x"], ev["y"], ev["t"].astype(np.float64) * 1e-6, ev["p"])
    want = orc.frame_disparity_map_bilinear(tables, lx, ly, evf, 0, apply_polarity=False)
    got = eng.frame(evf, view=0, output=e.OUT_DISPARITY, polarity=False, time_bounds=e.TBOUNDS_REDUCE, bilinear=True).cpu().numpy()
    assert np.array_equal(got, want)
    assert not eng.frame(ev[:0], view=0, output=e.OUT_DISPARITY, bilinear=True).cpu().numpy().any()
    # the nearest path is untouched by a bilinear frame before it
    assert np.array_equal(eng.frame(ev, view=0).cpu().numpy(), orc.frame_depth(tables, ev, 0))


def test_inverse_lut_builder_matches_reference_tables(manifest):
    """xm_build_inverse_lut = initUndistortRectifyMapInverse (cam_proj_calibration.py:31-41) on the device: the float32
    maps and the int16 tables of the default and the HD geometry carry the hashes of the REAL reference's tables, and
    equal OpenCV's undistortPoints value for value (a strongly distorted lens included)."""
    import os

    import cv2

    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps, inverse_rectify_map
    from xmaps_b200.engine import build_inverse_lut
    from xm_helpers import ROOT

    calib = os.path.join(ROOT, "data", "esl_calib_hhi.json")
    h = manifest["configs"]["default"]["hash"]
    maps = CamProjMaps(CamProjCalibrationParams.from_yaml(calib, 640, 480, 720, 1280), table_device="cuda:0")
    assert sha(maps.disp_cam_mapx_f32) == h["lut_x_f32"] and sha(maps.disp_cam_mapy_f32) == h["lut_y_f32"]
    assert sha(maps.disp_cam_mapx_i16) == h["lut_x"] and sha(maps.disp_cam_mapy_i16) == h["lut_y"]
    assert sha(maps.disp_proj_mapxy_i16) == h["remap_xy"]
    # HD geometry (BASELINE config 3) against the host builder, int16 table straight from the device
    p = CamProjCalibrationParams.from_yaml(calib, 1280, 720, 1080, 1920)
    k = p.camera_K.copy()
    k[:2, :] *= 2.0
    k[1, 2] += -120.0
    p.camera_K = k
    host = CamProjMaps(p)
    mx, my, xy = build_inverse_lut(p.camera_K, p.camera_D, host.R1, host.P1, (1280, 720), device="cuda:0", with_i16=True)
    assert np.array_equal(mx.cpu().numpy(), host.disp_cam_mapx_f32) and np.array_equal(my.cpu().numpy(), host.disp_cam_mapy_f32)
    assert np.array_equal(xy.cpu().numpy()[..., 0], host.disp_cam_mapx_i16) and np.array_equal(xy.cpu().numpy()[..., 1], host.disp_cam_mapy_i16)
    # all 12 non-tilt coefficients, strong distortion
    d12 = np.array([-0.31, 0.12, 1.5e-3, -2.1e-3, -0.02, 0.01, -0.004, 0.002, 1e-3, -2e-4, 5e-4, 1e-4])
    want_x, want_y = inverse_rectify_map(p.camera_K, d12, host.R1, host.P1, (1280, 720))
    got_x, got_y = build_inverse_lut(p.camera_K, d12, host.R1, host.P1, (1280, 720), device="cuda:0")
    assert np.array_equal(got_x.cpu().numpy(), want_x) and np.array_equal(got_y.cpu().numpy(), want_y)
    # no coefficient vector at all (OpenCV skips the iteration)
    pts = cv2.undistortPoints(np.array([[[3.0, 5.0]], [[100.0, 7.0]]], dtype=np.float32), p.camera_K, None, None, host.R1, host.P1)
    gx, gy = build_inverse_lut(p.camera_K, None, host.R1, host.P1, (101, 8), device="cuda:0")
    assert gx[5, 3].item() == pts[0, 0, 0] and gy[5, 3].item() == pts[0, 0, 1] and gx[7, 100].item() == pts[1, 0, 0]


# ------------------------------------------------------------------------------------------------
# full-size workloads (BASELINE.json configs 2 and 3)
# ------------------------------------------------------------------------------------------------
def test_5m_events_default_config(default):
    tables, _, eng = default
    e = E()
    ev = orc.synth_events(7, 5_000_000, 640, 480)
    dev = eng.events(ev)
    for view in (0, 1):
        want = orc.frame_depth(tables, ev, view)
        got = eng.frame(dev, view=view).cpu().numpy()
        assert np.array_equal(got, want), f"view {view}: {np.count_nonzero(got != want)} pixels differ"
        # idempotence + bounds-mode independence at full size
        assert np.array_equal(eng.frame(dev, view=view, time_bounds=e.TBOUNDS_REDUCE).cpu().numpy(), want)
        eng.set_option("stage_xmap", 0)
        try:
            assert np.array_equal(eng.frame(dev, view=view).cpu().numpy(), want)
        finally:
            eng.set_option("stage_xmap", 1)
    st = eng.status()
    assert st["n_valid"] == int((ev["p"] == 1).sum())


def test_hd_config_tables_and_frames(manifest):
    """BASELINE config 3 geometry: tables from the host builder (same OpenCV calls as the reference),
    X-map from the GPU builder; hashes recorded from the real reference."""
    import os

    from xmaps_b200.calibration import CamProjCalibrationParams, CamProjMaps
    from xmaps_b200.disparity import XMapsDisparity
    from xmaps_b200.time_map import ProjectorTimeMap
    from xm_helpers import ROOT

    cfg = manifest["configs"]["hd"]
    p = CamProjCalibrationParams.from_yaml(os.path.join(ROOT, "data", "esl_calib_hhi.json"), 1280, 720, 1080, 1920)
    k = p.camera_K.copy()
    k[:2, :] *= 2.0
    k[1, 2] += -120.0
    p.camera_K = k
    maps = CamProjMaps(p)
    tm = ProjectorTimeMap.from_calib(p, maps)
    xd = XMapsDisparity(calib_params=p, cam_proj_maps=maps, proj_time_map_rect=tm.projector_time_map_rectified)
    assert sha(np.asarray(xd.proj_x_map)) == cfg["hash"]["x_map"]
    eng = maps.engine("cuda:0")
    ev = orc.synth_events(3, 1_000_000, 1280, 720)
    assert sha(eng.frame(ev, view=0).cpu().numpy()) == cfg["hash"]["depth_proj_seed3_1m"]
    assert sha(eng.frame(ev, view=1).cpu().numpy()) == cfg["hash"]["depth_cam_seed3_1m"]
    assert eng.status()["n_inliers"] == cfg["n_inliers"]
    # the persistent batch kernel on this geometry (wider tile regions than the default one)
    assert eng.get_option("batch") == 1
    ev2 = orc.synth_events(4, 700_000, 1280, 720)
    for view, key in ((0, "depth_proj_seed3_1m"), (1, "depth_cam_seed3_1m")):
        out = eng.frame_batch([ev, ev2, ev], view=view).cpu().numpy()
        assert sha(out[0]) == cfg["hash"][key] and sha(out[2]) == cfg["hash"][key]
        assert np.array_equal(out[1], eng.frame(ev2, view=view).cpu().numpy())
