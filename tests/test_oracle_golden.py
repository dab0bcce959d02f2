"""Pin the oracle (oracle/xmaps_oracle.py) to vectors produced by the real reference
(tests/golden/generate_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import xmaps_oracle as orc
from xm_helpers import golden_frame, load_golden_tables


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.mark.parametrize(
    "cfg,seed,n,cam",
    [("default", 0, 100_000, (640, 480)), ("small", 2, 20_000, (160, 120))],
)
def test_per_event_stage_matches_reference(cfg, seed, n, cam):
    tables, _ = load_golden_tables(cfg)
    g = golden_frame(f"{cfg}_{n // 1000}k_proj")
    ev = orc.polarity_mask(orc.synth_events(seed, n, *cam))
    assert len(ev) == int(g["n_pos"][0])
    xr, yr = orc.rectify_i16(tables, ev)
    assert np.array_equal(xr, g["xr"]) and np.array_equal(yr, g["yr"])
    disp, mask = orc.event_disparity(tables, xr, yr, ev["t"])
    assert disp.dtype == np.int16
    assert np.array_equal(disp, g["disp"])
    assert np.array_equal(np.packbits(mask), g["mask_bits"])
    rect = orc.scatter_projector_view(tables, xr, yr, mask, disp)
    assert np.array_equal(rect, g["rect_map"])


@pytest.mark.parametrize("use_cv2", [True, False])
@pytest.mark.parametrize(
    "cfg,seed,n,cam",
    [("default", 0, 100_000, (640, 480)), ("small", 2, 20_000, (160, 120))],
)
def test_depth_frames_match_reference(cfg, seed, n, cam, use_cv2):
    tables, _ = load_golden_tables(cfg)
    ev = orc.synth_events(seed, n, *cam)
    for view, tag in ((orc.VIEW_PROJECTOR, "proj"), (orc.VIEW_CAMERA, "cam")):
        g = golden_frame(f"{cfg}_{n // 1000}k_{tag}")
        dm = orc.frame_disparity_map(tables, ev, view, use_cv2=use_cv2)
        assert np.array_equal(dm, g["disp_map"])
        depth = orc.disparity_to_depth(dm, tables.depth_scale)
        assert depth.dtype == np.float32
        assert np.array_equal(depth, g["depth"])
        bgr = orc.colorize(dm, tables.depth_scale, 0.1, 1.0)
        assert np.array_equal(bgr, g["bgr"])


def test_manifest_hashes_default(manifest, tables_default):
    h = manifest["configs"]["default"]["hash"]
    assert sha(tables_default.lut_x) == h["lut_x"]
    assert sha(tables_default.x_map) == h["x_map"]
    assert sha(tables_default.remap_xy) == h["remap_xy"]
    ev = orc.synth_events(0, 100_000, 640, 480)
    assert sha(ev) == h["events_seed0_100k"]
    assert sha(orc.frame_depth(tables_default, ev, orc.VIEW_PROJECTOR)) == h["depth_proj"]
    assert sha(orc.frame_depth(tables_default, ev, orc.VIEW_CAMERA)) == h["depth_cam"]


def test_one_million_event_frame_hashes(manifest, tables_default):
    h = manifest["configs"]["default"]["hash"]
    ev = orc.synth_events(1, 1_000_000, 640, 480)
    assert sha(orc.frame_depth(tables_default, ev, orc.VIEW_PROJECTOR)) == h["depth_proj_seed1_1m"]
    assert sha(orc.frame_depth(tables_default, ev, orc.VIEW_CAMERA)) == h["depth_cam_seed1_1m"]


def test_unsorted_timestamps(tables_small):
    ev = orc.synth_events(2, 20_000, 160, 120)
    np.random.default_rng(7).shuffle(ev)
    g = golden_frame("small_20k_shuffled_proj")
    assert np.array_equal(orc.frame_depth(tables_small, ev, orc.VIEW_PROJECTOR), g["depth"])


def test_x_map_builder_small():
    tables, z = load_golden_tables("small")
    w = int(z["consts"][2])
    x_map, _ = orc.build_x_map(z["time_map_rect"], w, int(z["consts"][0]), int(z["consts"][1]), num_scanlines=w)
    assert np.array_equal(x_map, tables.x_map)


def test_degenerate_frames(tables_small):
    empty = np.zeros(0, dtype=orc.EVENT_DTYPE)
    assert not orc.frame_depth(tables_small, empty, orc.VIEW_PROJECTOR).any()
    ev = orc.synth_events(5, 64, 160, 120)
    ev["t"] = 1234  # all timestamps equal -> NaN column -> 0
    assert orc.time_to_xmap_column(ev["t"], tables_small.t_px_scale).tolist() == [0] * 64
    neg = ev.copy()
    neg["p"] = 0
    assert not orc.frame_depth(tables_small, neg, orc.VIEW_CAMERA).any()
